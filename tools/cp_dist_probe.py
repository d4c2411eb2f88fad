"""Sharded constrained solve timing: run under torchrun (one rank per GPU), e.g.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
       tools/cp_dist_probe.py heavy 200
Prints, on rank 0, ms/node of the sharded solve and of the single-GPU solve on the same inputs (results compared)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench, consistent_viterbi_b200 as cv

kind = sys.argv[1]; budget = int(sys.argv[2])
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = bench.workload_cp(kind)
hm = cv.HMM(w["A"], w["B"], w["pi"])
args = (w["obs"], w["start"], w["comp"], w["ncomp"])
single = cv.cp_solve_arrays(hm, *args, max_nodes=2, device=local)
cv._lib.lib().cv_set_timing(1)
t0 = time.perf_counter()
single = cv.cp_solve_arrays(hm, *args, max_nodes=budget, device=local)
t_single = time.perf_counter() - t0
loop_single = cv._lib.lib().cv_last_kernel_ms(hm.device_handle(local))
grp = cv.CpDistGroup(hm, cap_N=w["N"], cap_terms=int((w["comp"] >= 0).sum()), device=local)
grp.solve(*args, max_nodes=2)
dist.barrier()
t0 = time.perf_counter()
r = grp.solve(*args, max_nodes=budget)
t_dist = time.perf_counter() - t0
loop_dist = cv._lib.lib().cv_last_kernel_ms(hm.device_handle(local))
same = bool((r["sol"] == single["sol"]).all() and r["obj"] == single["obj"] and r["explored"] == single["explored"])
tt = torch.tensor([t_dist, loop_dist], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"{kind} world {world} cuts {cv.plan_cuts(w['comp'], world).tolist()} nodes {r['explored']} identical {same} "
          f"sharded ms/node {1e3 * tt[0].item() / r['explored']:.4f} (device loop {tt[1].item() / r['explored']:.4f}) "
          f"single-GPU ms/node {1e3 * t_single / single['explored']:.4f} (device loop {loop_single / single['explored']:.4f})")
assert same
grp.close()
dist.barrier()
dist.destroy_process_group()
