"""Sharded constrained solve on 2+ GPUs (csrc/cp_dist.cuh): every rank must return exactly what the CPU oracle
and the single-GPU path return.  Needs >= 2 devices (skipped on the 1-GPU box; run with `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import consistent_viterbi_b200 as cv
    from oracle import pyoracle as po
    from util import random_hmm, random_superseq

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bad, n_sharded = [], 0
    try:
        for seed in range(24):
            rng = np.random.default_rng(7000 + seed)
            K = int(rng.choice([3, 6, 12, 16, 24, 45]))
            M = int(rng.integers(4, 40))
            A, B, pi = random_hmm(rng, K, M, zero_frac=0.1, ties=(seed % 5 == 0))
            nseq = int(rng.integers(5, 120))
            obs, start, comp, ncomp = random_superseq(rng, nseq, M, int(rng.integers(0, 4)), float(rng.choice([0.05, 0.2])), 1, 60)
            h = cv.HMM(A, B, pi)
            grp = cv.CpDistGroup(h, cap_N=len(obs), cap_terms=int((comp >= 0).sum()), device=rank)
            budget = 300
            res = grp.solve(obs, start, comp, ncomp, max_nodes=budget)
            res2 = grp.solve(obs, start, comp, ncomp, max_nodes=budget)        # buffers / epochs reused
            ref = po.cp_solve(A, B, pi, obs, start, comp, ncomp, max_nodes=budget)
            for got in (res, res2):
                same = (got["sol"] == ref["sol"]).all() and got["explored"] == ref["explored"] and \
                    np.float64(got["obj"]).tobytes() == np.float64(ref["obj"]).tobytes() and got["steps"] == ref["steps"]
                if not same:
                    bad.append(seed)
            cuts = cv.plan_cuts(comp, world)
            n_sharded += int(ncomp > 0 and cuts[1] < len(obs))
            grp.close()
            h.close()
        q.put((rank, bad, n_sharded))
    except Exception as e:  # noqa: BLE001
        q.put((rank, [repr(e)], -1))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_cp_matches_oracle(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, bad, n_sharded in res:
        assert bad == [], f"rank {rank}: {bad}"
        assert n_sharded >= 8
