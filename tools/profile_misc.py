"""One call each of cv_cfn_tables (configs[4] super-sequence) and cv_mle (POS shape, 1M sentences) -- the
workload ncu captures cfn_chain_kernel / mle_count_kernel from (see profiles/)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, consistent_viterbi_b200 as cv
w = bench.workload_cp("heavy")
hm = cv.HMM(w["A"], w["B"], w["pi"])
r = cv.cfn_tables(hm, w["obs"], w["start"], w["comp"], w["ncomp"])
print("cfn device ms", r["device_ms"], "boundaries", r["nboundaries"])
w = bench.workload_pos(0, 1000000)
tg = np.random.default_rng(3019).integers(0, w["K"], len(w["obs"])).astype(np.int32)
h2 = cv.HMM.new(w["K"], (w["B"].shape[1],))
print("mle count ms", h2.mle_arrays(w["obs"], tg, w["off"]), "elements", len(w["obs"]))
