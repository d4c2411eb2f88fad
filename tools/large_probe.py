import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, consistent_viterbi_b200 as cv
wl = bench.workload_large(0, int(sys.argv[1]), int(sys.argv[2]))
h = cv.HMM(wl["A"], wl["B"], wl["pi"])
L = cv._lib.lib()
cv.decode_batch(h, wl["obs"], wl["off"])
L.cv_set_timing(1)
t0 = time.perf_counter(); cv.decode_batch(h, wl["obs"], wl["off"]); dt = time.perf_counter() - t0
print("large", sys.argv[1:], "e2e_ms %.2f kernel_ms %.2f cells/s %.3e" % (1e3 * dt, L.cv_last_kernel_ms(h.device_handle()), wl["cells"] / (L.cv_last_kernel_ms(h.device_handle()) * 1e-3)))
