"""world_size-2 gloo test of the multi-GPU host logic (sharding + all-gather) on CPU.  The GPU decode is
replaced by the oracle here only to exercise the plumbing; the GPU path itself is covered by -m gpu tests."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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

    from consistent_viterbi_b200 import dist as cvd
    from oracle import pyoracle as po
    from util import random_batch, random_hmm

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(42)
    A, B, pi = random_hmm(rng, 6, 7)
    obs, off = random_batch(rng, 301, 7, 1, 25)

    def fake_gpu(hmm, o, f):
        return po.decode_batch(A, B, o, f)

    paths, scores, (b0, b1) = cvd.decode_batch_sharded(None, obs, off, gather=True, decode_fn=fake_gpu)
    ref_p, ref_s = po.decode_batch(A, B, obs, off)
    ok = bool((paths == ref_p).all() and scores.tobytes() == ref_s.tobytes())
    q.put((rank, ok, b0, b1))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_balanced():
    sys.path.insert(0, ROOT)
    from consistent_viterbi_b200.dist import local_slice, shard_bounds

    rng = np.random.default_rng(1)
    lens = rng.integers(1, 200, size=5000)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    for w in (1, 2, 4, 8):
        b = shard_bounds(off, w)
        assert b[0] == 0 and b[-1] == 5000 and (np.diff(b) >= 0).all()
        work = [off[b[r + 1]] - off[b[r]] for r in range(w)]
        assert max(work) - min(work) <= 2 * 200
    o, f, b0, b1 = local_slice(np.arange(off[-1]), off, 1, 4)
    assert f[0] == 0 and f[-1] == len(o) and o[0] == off[b0]
    # degenerate: more ranks than sequences
    b = shard_bounds(np.array([0, 3, 5]), 4)
    assert b[-1] == 2 and (np.diff(b) >= 0).all()


def test_two_rank_gloo_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res)
    res.sort()
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == 301
