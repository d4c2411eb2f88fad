"""Where the time goes in the device-resident step of a SHORT batch (a slice of a sharded batch).

    [CV_BT_PROF=1] python tools/small_batch_profile.py [nseq]     (default 125000)

Prints the forward / backtrace kernels timed alone (timing mode: sequential), the pipelined step, and with CV_BT_PROF=1
the library's own forward vs forward+backtrace times of the concurrent mode."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import consistent_viterbi_b200 as cv  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
wl = bench.workload_pos(0, B)
L = cv._lib.lib()
hmm = cv.HMM(wl["A"], wl["B"], wl["pi"])
h = hmm.device_handle(0)
N = len(wl["obs"])
d_obs = torch.from_numpy(wl["obs"].view(np.int32)).cuda(); d_off = torch.from_numpy(wl["off"]).cuda()
d_path = torch.empty(N, dtype=torch.int32, device="cuda"); d_score = torch.empty(B, dtype=torch.float64, device="cuda")
ml = int(np.diff(wl["off"]).max())
st = torch.cuda.current_stream()


def run(sync=0):
    cv._lib.check(L.cv_decode_batch_dev(h, d_obs.data_ptr(), d_off.data_ptr(), B, N, ml, d_path.data_ptr(), d_score.data_ptr(), st.cuda_stream, sync))


lens = np.diff(wl["off"])
print(f"B={B} N={N} max_len={ml} ideal share of the 1M step: {11.8 * wl['cells'] / 4.8612e10:.3f} ms; sequences > 100 steps: {(lens > 100).sum()}")
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(10):
    run()
e1.record(st); torch.cuda.synchronize()
print(f"pipelined device step {e0.elapsed_time(e1) / 10:.3f} ms")
e0.record(st)
for _ in range(10):
    run(1)
e1.record(st); torch.cuda.synchronize()
print(f"same with a status sync per step {e0.elapsed_time(e1) / 10:.3f} ms")
L.cv_set_timing(1)
for _ in range(3):
    run(1)
    print(f"timing mode: forward {L.cv_last_kernel_ms(h):.3f} ms, backtrace after it {L.cv_last_backtrace_ms(h):.3f} ms")
L.cv_set_timing(0)
