import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, consistent_viterbi_b200 as cv
w = bench.workload_cp("trucks")
hm = cv.HMM(w["A"], w["B"], w["pi"])
cv.cp_solve_arrays(hm, w["obs"], w["start"], w["comp"], w["ncomp"], max_nodes=2)
t0 = time.perf_counter()
r = cv.cp_solve_arrays(hm, w["obs"], w["start"], w["comp"], w["ncomp"], max_nodes=int(sys.argv[1]))
dt = time.perf_counter() - t0
print("trucks-like full solve: nodes", r["explored"], "steps", r["steps"], "obj", r["obj"], "s", dt, "ms/node", 1e3*dt/r["explored"])
