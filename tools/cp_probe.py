import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, consistent_viterbi_b200 as cv
kind = sys.argv[1]; budget = int(sys.argv[2])
w = bench.workload_cp(kind)
hm = cv.HMM(w["A"], w["B"], w["pi"])
cv.cp_solve_arrays(hm, w["obs"], w["start"], w["comp"], w["ncomp"], max_nodes=2)
cv._lib.lib().cv_set_timing(1)
t0 = time.perf_counter()
r = cv.cp_solve_arrays(hm, w["obs"], w["start"], w["comp"], w["ncomp"], max_nodes=budget)
dt = time.perf_counter() - t0
loop_ms = cv._lib.lib().cv_last_kernel_ms(hm.device_handle(-1))
print(kind, "device loop ms/node", loop_ms / r["explored"], "(setup + copies", 1e3*dt - loop_ms, "ms)")
print(kind, "nodes", r["explored"], "steps", r["steps"], "ms", 1e3*dt, "ms/node", 1e3*dt/r["explored"])
