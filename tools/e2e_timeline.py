"""Timeline of cv_decode_batch (host buffers) at several batch sizes -- what a rank of an N-GPU strong-scaling run sees.

    CV_E2E_PROF=1 python tools/e2e_timeline.py [nseq ...]      (default: 1000000 500000 250000 125000)

Prints the library's own timeline (stderr, CV_E2E_PROF) and the wall-clock per call next to the device-resident step."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import consistent_viterbi_b200 as cv  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [1_000_000, 500_000, 250_000, 125_000]
L = cv._lib.lib()
wl = bench.workload_pos(0, max(sizes))
hmm = cv.HMM(wl["A"], wl["B"], wl["pi"])
h = hmm.device_handle(0)
for B in sizes:
    off = wl["off"][: B + 1].copy()
    N = int(off[-1])
    obs = wl["obs"][:N]

    def pinned(arr):
        p = L.cv_host_alloc(arr.nbytes)
        np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(arr.nbytes,))[:] = arr.view(np.uint8).reshape(-1)
        return p
    p_obs, p_off = pinned(np.ascontiguousarray(obs)), pinned(off)
    p_obs16 = pinned(np.ascontiguousarray(obs.astype(np.uint16)))
    p_path, p_score, p_path8 = L.cv_host_alloc(4 * N), L.cv_host_alloc(8 * B), L.cv_host_alloc(N)
    for name, fn in (("u32", lambda: L.cv_decode_batch(h, p_obs, p_off, B, p_path, p_score)),
                     ("u16u8", lambda: L.cv_decode_batch_u16u8(h, p_obs16, p_off, B, p_path8, p_score))):
        for _ in range(3):
            cv._lib.check(fn())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            cv._lib.check(fn())
        dt = (time.perf_counter() - t0) / 5
        print(f"B={B} {name}: e2e {1e3 * dt:.3f} ms/call", flush=True)
    d_obs = torch.from_numpy(obs.view(np.int32)).cuda(); d_off = torch.from_numpy(off).cuda()
    d_path = torch.empty(N, dtype=torch.int32, device="cuda"); d_score = torch.empty(B, dtype=torch.float64, device="cuda")
    ml = int(np.diff(off).max())
    st = torch.cuda.current_stream()
    for _ in range(3):
        L.cv_decode_batch_dev(h, d_obs.data_ptr(), d_off.data_ptr(), B, N, ml, d_path.data_ptr(), d_score.data_ptr(), st.cuda_stream, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        L.cv_decode_batch_dev(h, d_obs.data_ptr(), d_off.data_ptr(), B, N, ml, d_path.data_ptr(), d_score.data_ptr(), st.cuda_stream, 0)
    e1.record(st); torch.cuda.synchronize()
    print(f"B={B} device-resident step {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
    for p in (p_obs, p_off, p_obs16, p_path, p_score, p_path8):
        L.cv_host_free(p)
