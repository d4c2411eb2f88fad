"""Times the one collective of the sharded decode in isolation: an in-place all_gather_into_tensor of a
[world][bytes] buffer with the row size of a 1 M-sentence POS batch cut into `world` slices.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/allgather_probe.py

Prints the time per collective for NCCL as configured by the environment (NCCL_ALGO, NCCL_PROTO, NCCL_NVLS_ENABLE ...)."""
import os

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
if os.environ.get("PROBE_EAGER"):          # as bench.py does: eager communicator bound to the device
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    dist.init_process_group("nccl")
if os.environ.get("PROBE_LIB"):            # a decode through the library first (its streams, carve-out settings, workspaces)
    import sys
    import numpy as np
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import consistent_viterbi_b200 as cv
    wl = bench.workload_pos(0, 250_000)
    hmm = cv.HMM(wl["A"], wl["B"], wl["pi"])
    for _ in range(3):
        cv.decode_batch(hmm, wl["obs"], wl["off"], device=local)
row = (1_000_000 // world) * 8 + (25_006_022 // world) + 64          # scores f64 + paths u8 of one slice
row = (row + 255) // 256 * 256
buf = torch.zeros(world, row, dtype=torch.uint8, device="cuda")
for _ in range(10):
    dist.all_gather_into_tensor(buf.view(-1), buf[rank])
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    dist.all_gather_into_tensor(buf.view(-1), buf[rank])
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 50], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    env = {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}
    print(f"world {world} row {row} B: all_gather_into_tensor {t.item():.4f} ms  env {env}", flush=True)
dist.destroy_process_group()
