"""Randomised shape sweep (seeded): many (K, M, batch, length, sparsity, tie) combinations through every kernel
family against the oracle -- the GPU-side counterpart of the reference-less property testing SURVEY section 4 asks for."""
import numpy as np
import pytest

import consistent_viterbi_b200 as cv
from oracle import pyoracle as po
from util import random_batch, random_hmm, random_superseq

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(4))
def test_r1_random_shapes(seed):
    rng = np.random.default_rng(7000 + seed)
    L = cv._lib.lib()
    for it in range(14):
        K = int(rng.choice([1, 2, 3, 5, 8, 11, 16, 17, 25, 32, 33, 40, 47, 63, 64, 66, 90, 129, 131]))
        M = int(rng.integers(1, 50))
        Bn = int(rng.integers(1, 400)) if K <= 64 else int(rng.integers(1, 150))
        tmax = int(rng.integers(1, 45)) if K <= 64 else int(rng.integers(1, 12))
        A, B, pi = random_hmm(rng, K, M, zero_frac=float(rng.choice([0.0, 0.1, 0.5, 0.9])), ties=bool(it % 5 == 0))
        obs, off = random_batch(rng, Bn, M, 1, tmax)
        rp, rs = po.decode_batch(A, B, obs, off, nthreads=8)
        h = cv.HMM(A, B, pi)
        try:
            for chain_max in ((-1, 0) if K <= 64 else (-1,)):
                L.cv_debug_set_chain_max_batch(chain_max)
                p, s = cv.decode_batch(h, obs, off)
                assert (p == rp).all(), f"K={K} M={M} B={Bn} tmax={tmax} chain_max={chain_max}"
                assert s.tobytes() == rs.tobytes(), f"K={K} M={M} B={Bn} tmax={tmax} chain_max={chain_max}"
        finally:
            L.cv_debug_set_chain_max_batch(-1)
        h.close()


@pytest.mark.parametrize("seed", range(4))
def test_r2_random_shapes(seed):
    rng = np.random.default_rng(8000 + seed)
    for it in range(10):
        K = int(rng.choice([1, 2, 4, 7, 8, 9, 16, 17, 24, 31, 32, 33, 48, 64]))
        M = int(rng.integers(1, 30))
        A, B, pi = random_hmm(rng, K, M, zero_frac=float(rng.choice([0.0, 0.05, 0.3])), ties=bool(it % 4 == 0))
        obs, start, comp, ncomp = random_superseq(rng, int(rng.integers(1, 25)), M, int(rng.integers(0, 5)),
                                                  float(rng.choice([0.02, 0.1, 0.4])), 1, int(rng.integers(2, 40)))
        budget = 40
        h = cv.HMM(A, B, pi)
        try:
            ref = po.cp_solve(A, B, pi, obs, start, comp, ncomp, max_nodes=budget, trace_nodes=budget, want_state=True)
            ref_err = None
        except po.OracleError as e:
            ref_err = e.code
        try:
            got = cv.cp_solve_arrays(h, obs, start, comp, ncomp, max_nodes=budget, want_state=True, want_ub=budget)
            got_err = None
        except cv.CvError as e:
            got_err = e.code
        assert ref_err == got_err, f"K={K} M={M} ncomp={ncomp}: {ref_err} vs {got_err}"
        if ref_err is None:
            tag = f"K={K} M={M} N={len(obs)} ncomp={ncomp}"
            assert got["explored"] == ref["explored"] and got["steps"] == ref["steps"], tag
            n = int(ref["explored"])
            assert got["ub"][:n].tobytes() == ref["ub"][:n].tobytes(), tag
            assert np.float64(got["obj"]).tobytes() == np.float64(ref["obj"]).tobytes(), tag
            assert (got["sol"] == ref["sol"]).all(), tag
            assert got["delta"].tobytes() == ref["delta"].tobytes() and (got["psi"] == ref["psi"]).all(), tag
        h.close()
