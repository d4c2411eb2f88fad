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

    def fake_gpu(o, f):
        return po.decode_batch(A, B, o, f)

    paths, scores, (b0, b1) = cvd.decode_batch_sharded(None, obs, off, gather=True, decode_fn=fake_gpu)
    ref_p, ref_s = po.decode_batch(A, B, obs, off)
    ok = bool((paths == ref_p).all() and scores.tobytes() == ref_s.tobytes())
    # the reusable form: same layout, new observations every step; results stay in the padded gather buffers
    sd = cvd.ShardedDecoder(None, off, decode_fn=fake_gpu)
    for it in range(2):
        obs2 = rng.integers(0, 7, size=len(obs)).astype(np.uint32)          # same stream on both ranks
        sd.load_obs(obs2)
        sd.step()
        rp2, rs2 = po.decode_batch(A, B, obs2, off)
        ok = ok and bool((sd.paths().numpy().view(np.uint32) == rp2).all()) and not sd.narrow_paths and sd.scores().numpy().tobytes() == rs2.tobytes()
    ok = ok and sd.gpaths.shape[0] == world and (sd.b0, sd.b1) == (b0, b1)
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


# ---- sharded constrained solve (csrc/cp_dist.cuh): host logic + the ownership rules, on CPU ----------------
def _cp_instance(seed):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import py_restatement as pr
    from util import random_hmm, random_superseq

    rng = np.random.default_rng(seed)
    K, M = int(rng.integers(2, 5)), 5
    A, B, pi = random_hmm(rng, K, M, zero_frac=0.15, ties=(seed % 3 == 0))
    obs, start, comp, ncomp = random_superseq(rng, int(rng.integers(4, 9)), M, int(rng.integers(1, 4)), 0.3, 1, 7)
    hmm = pr.HMM(A.tolist(), B.tolist(), pi.tolist())

    def elements():
        els, t = [], 0
        for i in range(len(obs)):
            t = 0 if start[i] else t + 1
            els.append(pr.Element(t, int(obs[i]), int(comp[i]), comp[i] >= 0))
        return els

    return hmm, elements, comp, ncomp


def test_plan_cuts_matches_model_and_is_valid():
    """cv_cp_plan_cuts (C ABI, pure host) == the Python model's rule; inner cuts are positions of component 0."""
    sys.path.insert(0, ROOT)
    import consistent_viterbi_b200 as cv
    from oracle.py_sharded_model import plan_cuts as model_cuts

    rng = np.random.default_rng(5)
    for it in range(200):
        N = int(rng.integers(1, 400))
        comp = np.full(N, -1, dtype=np.int32)
        mask = rng.random(N) < rng.choice([0.02, 0.2, 0.6])
        comp[mask] = rng.integers(0, 3, size=int(mask.sum()))
        for R in (1, 2, 3, 4, 8):
            cuts = cv.plan_cuts(comp, R)
            ref = model_cuts(comp.tolist(), N, R)
            if ref is None:
                assert cuts.tolist() == [0] + [N] * R
                continue
            assert cuts.tolist() == ref
            assert cuts[0] == 0 and cuts[-1] == N and (np.diff(cuts) > 0).all()
            assert all(comp[c] == 0 for c in cuts[1:-1])
    big = np.full(640000, -1, dtype=np.int32)
    big[rng.random(640000) < 0.05] = 0
    cuts = cv.plan_cuts(big, 8)
    assert np.abs(np.diff(cuts) - 80000).max() < 200                     # balanced to within a few clamped gaps


def _cp_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import py_restatement as pr
    from oracle.py_sharded_model import ShardedCPSolver, plan_cuts

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def exchange(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    ok, n = True, 0
    for seed in range(60):
        hmm, elements, comp, ncomp = _cp_instance(seed)
        cuts = plan_cuts(comp.tolist(), len(comp), world) if ncomp > 0 else None
        if cuts is None:
            continue
        ref = pr.CPSolver(hmm, elements(), ncomp)
        ref.solve()
        s = ShardedCPSolver(hmm, elements(), ncomp, rank, world, cuts, exchange)
        s.solve()
        same_ub = [np.float64(u).tobytes() for u in s.ub_log] == [np.float64(u).tobytes() for u in ref.ub_log]
        ok = ok and s.best_sol == ref.best_sol and s.explored == ref.explored and same_ub
        ok = ok and np.float64(s.best_obj).tobytes() == np.float64(ref.best_obj).tobytes()
        n += 1
    q.put((rank, ok, n))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_cp_model():
    """Two gloo ranks run the sharded solve's ownership rules (rows a rank does not own are poisoned) and
    exchange only bound terms, backtrack maps and solution rows; the result must equal the sequential
    restatement of cp.rs bit for bit, node for node."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == res[1][2] and res[0][2] >= 20
