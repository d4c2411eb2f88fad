import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, consistent_viterbi_b200 as cv
w = bench.workload_cp("heavy")
hm = cv.HMM(w["A"], w["B"], w["pi"])
for _ in range(3):
    r = cv.cfn_tables(hm, w["obs"], w["start"], w["comp"], w["ncomp"])
    print("cfn device ms", r["device_ms"], "boundaries", r["nboundaries"])
