import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import consistent_viterbi_b200 as cv
L = cv._lib.lib()
for hname in "ABC":
    z = np.load(f"tests/golden/ar_house_{hname}.npz")
    hm = cv.HMM(z["logA"], z["logB"], z["logPi"])
    off = z["seq_off"]; lens = np.diff(off)
    cv.decode_batch(hm, z["obs"], off)
    L.cv_set_timing(1)
    for _ in range(3):
        t0 = time.perf_counter(); cv.decode_batch(hm, z["obs"], off); dt = time.perf_counter() - t0
    L.cv_set_timing(0)
    print(hname, "K", hm.nstates(), "B", len(lens), "maxlen", lens.max(), "sum", lens.sum(), "e2e_ms %.3f" % (1e3*dt), "kernel_ms %.3f" % L.cv_last_kernel_ms(hm.device_handle()),
          "us/step %.3f" % (1e3 * L.cv_last_kernel_ms(hm.device_handle()) / lens.max()))
