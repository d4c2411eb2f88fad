"""CPU tests: the C oracle against (1) brute-force path enumeration and (2) the
independent pure-Python restatement.  The reference ships no tests or golden
vectors (parity unpinned), so these are the pins the oracle has."""
import math

import numpy as np
import pytest

from oracle import py_restatement as pr
from oracle import pyoracle as po
from util import random_hmm, random_superseq


def _bits(x):
    return np.asarray(x, dtype=np.float64).tobytes()


@pytest.mark.parametrize("seed", range(6))
def test_r1_oracle_vs_restatement_and_bruteforce(seed):
    rng = np.random.default_rng(100 + seed)
    nbf = 0
    for it in range(120):
        K = int(rng.integers(1, 6)); M = int(rng.integers(1, 5)); T = int(rng.integers(1, 7))
        A, B, pi = random_hmm(rng, K, M, ties=(it % 3 == 0))
        obs = rng.integers(0, M, size=T).astype(np.uint32)
        path, score = po.decode(A, B, obs)
        h = pr.HMM(A, B, pi)
        seq = [int(o) for o in obs]
        p2, arr, bt = pr.decode(seq, h)
        assert list(path) == p2
        assert _bits(score) == _bits(arr[-1][p2[-1]])
        d, ps = po.decode_trace(A, B, obs)
        assert _bits(d) == _bits(arr)
        assert (ps == np.array(bt)).all()
        if K ** T <= 4096:
            bf = pr.r1_bruteforce_best(seq, h)
            assert bf == score                      # exact (fl(x+c) monotone in x)
            if score > -math.inf:
                assert pr.r1_path_score(p2, seq, h) == score
            nbf += 1
    assert nbf > 50


def test_r1_edge_cases():
    rng = np.random.default_rng(7)
    A, B, pi = random_hmm(rng, 4, 3)
    # T = 1 -> path [0], score 0.0 (viterbi.rs:6,24: argmax of an all-zero row)
    path, score = po.decode(A, B, np.array([2], dtype=np.uint32))
    assert list(path) == [0] and score == 0.0
    # T = 0 -> reference panics (usize underflow)
    with pytest.raises(po.OracleError) as e:
        po.decode(A, B, np.zeros(0, dtype=np.uint32))
    assert e.value.code == po.ERR_EMPTY
    # all emissions -inf -> every delta -inf, psi stays 0, path all zeros
    Binf = np.full_like(B, -np.inf)
    path, score = po.decode(A, Binf, np.array([0, 1, 2, 1], dtype=np.uint32))
    assert list(path) == [0, 0, 0, 0] and score == -np.inf
    # K = 1
    path, score = po.decode(np.array([[-0.5]]), np.array([[-1.0, -2.0]]), np.array([0, 1, 1], dtype=np.uint32))
    assert list(path) == [0, 0, 0] and score == ((0.0 + -0.5) + -2.0 + -0.5) + -2.0
    # observation out of range -> index panic
    with pytest.raises(po.OracleError) as e:
        po.decode(A, B, np.array([0, 3], dtype=np.uint32))
    assert e.value.code == po.ERR_ARG


def test_r1_batch_matches_single_and_threads():
    rng = np.random.default_rng(11)
    A, B, pi = random_hmm(rng, 7, 9)
    lens = rng.integers(1, 30, size=200)
    off = np.zeros(201, dtype=np.int64); off[1:] = np.cumsum(lens)
    obs = rng.integers(0, 9, size=int(off[-1])).astype(np.uint32)
    p1, s1 = po.decode_batch(A, B, obs, off, nthreads=1)
    p4, s4 = po.decode_batch(A, B, obs, off, nthreads=4)
    assert (p1 == p4).all() and _bits(s1) == _bits(s4)
    for b in (0, 17, 199):
        pb, sb = po.decode(A, B, obs[off[b]:off[b + 1]])
        assert (pb == p1[off[b]:off[b + 1]]).all() and _bits(sb) == _bits(s1[b])


def _elements(obs, start, comp):
    els, tt = [], 0
    for i in range(len(obs)):
        if start[i]:
            tt = 0
        els.append(pr.Element(tt, int(obs[i]), int(comp[i]), comp[i] >= 0))
        tt += 1
    return els


@pytest.mark.parametrize("seed", range(6))
def test_r2_cp_oracle_vs_restatement(seed):
    rng = np.random.default_rng(200 + seed)
    nok = 0
    for it in range(100):
        K = int(rng.integers(1, 5)); M = int(rng.integers(1, 4))
        A, B, pi = random_hmm(rng, K, M, ties=(it % 3 == 0), zero_frac=0.15)
        obs, start, comp, ncomp = random_superseq(rng, int(rng.integers(1, 5)), M, int(rng.integers(0, 4)), 0.35)
        s = pr.CPSolver(pr.HMM(A, B, pi), _elements(obs, start, comp), ncomp)
        try:
            s.solve(); perr = False
        except AssertionError:
            perr = True
        try:
            r = po.cp_solve(A, B, pi, obs, start, comp, ncomp, trace_nodes=4096, want_state=True); cerr = False
        except po.OracleError as e:
            assert e.code == po.ERR_ASSERT; cerr = True
        assert perr == cerr
        if perr:
            continue
        nok += 1
        assert list(r["sol"]) == s.best_sol
        assert _bits(r["obj"]) == _bits(s.best_obj)
        assert r["explored"] == s.explored and r["steps"] == s.steps
        assert _bits(s.ub_log) == _bits(r["ub"][: s.explored])
        assert _bits(r["delta"]) == _bits(s.final_state[0])
        assert (r["psi"] == np.array(s.final_state[1], dtype=np.uint64)).all()
    assert nok > 60


def test_r2_no_constraints_is_chain_viterbi():
    """ncomp == 0: obj = max(delta[N-1]) (cp.rs:139-142); with a single sequence whose pi
    row is all zero and T small the solution must be a best path of the chain score."""
    rng = np.random.default_rng(5)
    for _ in range(50):
        K = int(rng.integers(2, 4)); M = 3; T = int(rng.integers(2, 6))
        A, B, pi = random_hmm(rng, K, M, zero_frac=0.0)
        obs = rng.integers(0, M, size=T).astype(np.uint32)
        start = np.zeros(T, dtype=np.uint8); start[0] = 1
        comp = np.full(T, -1, dtype=np.int32)
        r = po.cp_solve(A, B, pi, obs, start, comp, 0)
        els = _elements(obs, start, comp)
        h = pr.HMM(A, B, pi)
        sc = pr.r2_chain_score([int(x) for x in r["sol"]], els, h)
        assert sc == r["obj"]
        import itertools
        best = max(pr.r2_chain_score(p, els, h) for p in itertools.product(range(K), repeat=T))
        assert abs(best - r["obj"]) <= 1e-12 * max(1.0, abs(best))   # selection uses fl(d+a): ulp-level only


def test_r2_max_nodes_budget():
    rng = np.random.default_rng(9)
    A, B, pi = random_hmm(rng, 4, 3, zero_frac=0.0)
    obs, start, comp, ncomp = random_superseq(rng, 4, 3, 3, 0.4)
    full = po.cp_solve(A, B, pi, obs, start, comp, ncomp)
    cut = po.cp_solve(A, B, pi, obs, start, comp, ncomp, max_nodes=3)
    assert cut["explored"] == min(3, full["explored"])


def test_oracle_mle_against_numpy_counts():
    """cvo_mle (literal hmm.rs:30-62 + log) on a zero initial model = count ratios; numpy's log is not glibc's, so
    values are compared to 1e-15 relative and zeros / -inf exactly.  A random initial model changes every entry."""
    rng = np.random.default_rng(11)
    K, M, B = 5, 7, 300
    lens = rng.integers(1, 9, size=B)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    obs = rng.integers(0, M, off[-1]).astype(np.uint32)
    tags = rng.integers(0, K, off[-1]).astype(np.int32)
    la, lb, lpi = po.mle(np.zeros((K, K)), np.zeros((K, M)), np.zeros(K), obs, tags, off)
    ca, cb, cpi, seen, end = np.zeros((K, K)), np.zeros((K, M)), np.zeros(K), np.zeros(K), np.zeros(K)
    for i in range(B):
        s, e = off[i], off[i + 1]
        cpi[tags[s]] += 1
        for t in range(s, e):
            cb[tags[t], obs[t]] += 1
            seen[tags[t]] += 1
            if t + 1 < e:
                ca[tags[t], tags[t + 1]] += 1
        end[tags[e - 1]] += 1
    with np.errstate(divide="ignore"):
        ra = np.log(ca / (seen - end)[:, None]) / np.log(10.0)
        rb = np.log(cb / seen[:, None]) / np.log(10.0)
        rpi = np.log(cpi / B) / np.log(10.0)
    for got, ref in ((la, ra), (lb, rb), (lpi, rpi)):
        assert (np.isneginf(got) == np.isneginf(ref)).all()
        fin = np.isfinite(ref)
        assert np.allclose(got[fin], ref[fin], rtol=1e-15, atol=0)
    a0 = rng.random((K, K))
    la2, _, _ = po.mle(a0, np.zeros((K, M)), np.zeros(K), obs, tags, off)
    assert not np.isneginf(la2).any() and (la2 != la).any()
    with pytest.raises(po.OracleError):
        po.mle(np.zeros((K, K)), np.zeros((K, M)), np.zeros(K), obs, np.where(np.arange(len(tags)) == 3, -1, tags), off)


def test_oracle_cfn_longest_path_bruteforce():
    """cvo_cfn_tables on a super-sequence with exactly two boundaries: the (c0, c1) table is longest_path
    (cfn.rs:11-35) for every (n1, n2); checked against enumeration of all state paths between the boundaries with
    the left-to-right rounded score (the DP's value, fl being monotone) and the clamps of cfn.rs:16-22."""
    import itertools

    rng = np.random.default_rng(21)
    for it in range(30):
        K, M = int(rng.integers(2, 4)), 4
        A, B, pi = random_hmm(rng, K, M, zero_frac=0.15, ties=(it % 4 == 0))
        L = int(rng.integers(1, 6))                                # rows between the two boundaries + 1
        N = L + 3
        obs = rng.integers(0, M, N).astype(np.uint32)
        start = np.zeros(N, dtype=np.uint8); start[0] = 1
        if it % 3 == 0 and L > 1:
            start[2] = 1                                          # a sequence start inside the segment: pi instead of a
        comp = np.full(N, -1, dtype=np.int32)
        t0, t1 = 1, 1 + L
        comp[t0], comp[t1] = 0, 1
        if L > 2 and it % 2 == 0:
            comp[t0 + 1] = 0                                      # same component again: clamped to n_from, no new boundary
        r = po.cfn_tables(A, B, pi, obs, start, comp, 2)
        assert r["nboundaries"] == 2
        for n1 in range(K):
            for n2 in range(K):
                best = -math.inf
                for mid in itertools.product(range(K), repeat=L - 1):
                    path = (n1,) + mid + (n2,)
                    if any(comp[t0 + i] >= 0 and path[i] != n1 for i in range(1, L)):
                        continue                                  # clamped rows only keep n_from
                    s = 0.0
                    for i in range(1, L + 1):
                        t = t0 + i
                        tr = pi[path[i]] if start[t] else A[path[i - 1], path[i]]
                        s = (s + tr) + B[path[i], obs[t]]
                    best = max(best, s)
                exp = 0.0 if best == -math.inf else best          # -inf costs are not accumulated (cfn.rs:123)
                assert r["tables"][0, 1, n1, n2] == exp and r["tables"][1, 0, n2, n1] == exp


def test_argmax_matches_ndarray_stats_published_cases():
    """The one third-party semantic on the path: ndarray-stats 0.5 `QuantileExt::argmax` (call sites viterbi.rs:16,24,
    cp.rs:39,53,74,86).  The crate is not under /root/reference; these are the cases of its own test suite
    (tests/quantile.rs::test_argmax, restated on the flattened row-major vector) plus the rule its implementation
    documents: the running maximum starts at the first element and only a strictly Greater element replaces it."""
    assert po.argmax([1, 5, 3, 2, 0, 6]) == 5                       # array![[1, 5, 3], [2, 0, 6]].argmax() == Ok((1, 2))
    assert po.argmax([1., 5., 3., 2., 0., 6.]) == 5
    with pytest.raises(po.OracleError) as e:                        # [[1., 5., 3.], [2., NAN, 6.]] -> Err(UndefinedOrder)
        po.argmax([1., 5., 3., 2., np.nan, 6.])
    assert e.value.code == 2
    with pytest.raises(po.OracleError) as e:                        # array![[], []] -> Err(EmptyInput)
        po.argmax([])
    assert e.value.code == 1
    assert po.argmax([3., 7., 7., 1.]) == 1                         # equal maxima: the first one stays
    assert po.argmax([-np.inf, -np.inf, -np.inf]) == 0
    assert po.argmax([-0.0, 0.0]) == 0 and po.argmax([0.0, -0.0]) == 0
