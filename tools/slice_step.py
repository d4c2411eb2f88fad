"""Device-resident step of ONE slice of the sharded 1 M-sentence POS batch on one GPU (what rank r of an N-GPU run
decodes), e.g. the slice that holds the batch's longest sentence:   python tools/slice_step.py <world> <rank>"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import consistent_viterbi_b200 as cv  # noqa: E402
from consistent_viterbi_b200.dist import shard_bounds  # noqa: E402

world, rank = int(sys.argv[1]), int(sys.argv[2])
wl = bench.workload_pos(0, 1_000_000)
b = shard_bounds(wl["off"], world)
b0, b1 = int(b[rank]), int(b[rank + 1])
off = (wl["off"][b0:b1 + 1] - wl["off"][b0]).copy()
obs = wl["obs"][wl["off"][b0]:wl["off"][b1]].copy()
B, N, ml = b1 - b0, int(off[-1]), int(np.diff(off).max())
L = cv._lib.lib()
hmm = cv.HMM(wl["A"], wl["B"], wl["pi"])
h = hmm.device_handle(0)
d_obs = torch.from_numpy(obs.view(np.int32)).cuda(); d_off = torch.from_numpy(off).cuda()
d_path = torch.empty(N, dtype=torch.int32, device="cuda"); d_score = torch.empty(B, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream()


def run():
    cv._lib.check(L.cv_decode_batch_dev(h, d_obs.data_ptr(), d_off.data_ptr(), B, N, ml, d_path.data_ptr(), d_score.data_ptr(), st.cuda_stream, 0))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(10):
    run()
e1.record(st); torch.cuda.synchronize()
print(f"world {world} rank {rank}: B={B} max_len={ml} step {e0.elapsed_time(e1) / 10:.3f} ms  "
      f"env {dict((k, v) for k, v in os.environ.items() if k.startswith('CV_'))}  checksum {int(d_path.sum().item())}", flush=True)
