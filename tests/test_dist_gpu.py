"""Batched decode sharded over 2+ GPUs (dist.ShardedDecoder): every rank must hold exactly the oracle's paths and
scores after the all-gather (NCCL).  Needs >= 2 devices (skipped on the 1-GPU box; run with `gpurun --gpus 2`)."""
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
    from consistent_viterbi_b200.dist import ShardedDecoder, decode_batch_sharded
    from oracle import pyoracle as po
    from util import random_batch, random_hmm

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    bad = []
    try:
        for case, (K, Bn) in enumerate([(45, 60000), (12, 3000), (70, 900)]):       # tile kernel / chain kernel / large-K (u32 paths)
            rng = np.random.default_rng(8800 + case)
            M = 50
            A, B, pi = random_hmm(rng, K, M, zero_frac=0.1)
            obs, off = random_batch(rng, Bn, M, 1, 40)
            rp, rs = po.decode_batch(A, B, obs, off, nthreads=4)
            h = cv.HMM(A, B, pi)
            sd = ShardedDecoder(h, off, device=rank)
            for it in range(2):                                                      # reused with new observations
                o2 = obs if it == 0 else np.roll(obs, 7)
                r2p, r2s = (rp, rs) if it == 0 else po.decode_batch(A, B, o2, off, nthreads=4)
                sd.load_obs(o2)
                sd.step()
                torch.cuda.synchronize()
                got = sd.paths().cpu().numpy()
                got = got.astype(np.uint32) if sd.narrow_paths else got.view(np.uint32)
                if not ((got == r2p).all() and sd.scores().cpu().numpy().tobytes() == r2s.tobytes()):
                    bad.append((case, it))
            p, s, _ = decode_batch_sharded(h, obs, off, device=rank)
            if not ((p == rp).all() and s.tobytes() == rs.tobytes()):
                bad.append((case, "helper"))
            if sd.narrow_paths != (K <= 64):
                bad.append((case, "narrow flag"))
            h.close()
        q.put((rank, bad))
    except Exception as e:  # noqa: BLE001
        q.put((rank, [repr(e)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_decode_matches_oracle(world):
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
    for rank, bad in res:
        assert bad == [], f"rank {rank}: {bad}"
