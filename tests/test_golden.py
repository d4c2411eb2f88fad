"""Golden fixtures (tests/golden/, made by tools/make_golden.py): CPU tests pin the oracle and the independent
Python restatement to the committed outputs; GPU tests pin the CUDA path to the same outputs through the C ABI.
The reference itself has no golden vectors (parity unpinned); these are the oracle's outputs frozen at commit time."""
import os

import numpy as np
import pytest

from oracle import py_restatement as pr
from oracle import pyoracle as po

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(x):
    return np.asarray(x, dtype=np.float64).tobytes()


def _r1_cases():
    z = np.load(os.path.join(G, "r1_small.npz"))
    for i in range(int(z["ncases"])):
        yield {k: z[f"c{i}_{k}"] for k in ("A", "B", "pi", "obs", "off", "paths", "scores")}


def _r2_cases():
    z = np.load(os.path.join(G, "r2_small.npz"))
    for i in range(int(z["ncases"])):
        yield {k: z[f"c{i}_{k}"] for k in ("A", "B", "pi", "obs", "start", "comp", "ncomp", "sol", "obj", "explored", "steps", "ub")}


def test_oracle_matches_golden_r1():
    for c in _r1_cases():
        p, s = po.decode_batch(c["A"], c["B"], c["obs"], c["off"])
        assert (p == c["paths"]).all() and _bits(s) == _bits(c["scores"])


def test_python_restatement_matches_golden_r1():
    for c in list(_r1_cases())[:3]:
        h = pr.HMM(c["A"], c["B"], c["pi"])
        for b in range(0, len(c["off"]) - 1, 7):
            seq = [int(o) for o in c["obs"][c["off"][b]:c["off"][b + 1]]]
            p, arr, _ = pr.decode(seq, h)
            assert p == list(c["paths"][c["off"][b]:c["off"][b + 1]])
            assert _bits(arr[-1][p[-1]]) == _bits(c["scores"][b])


def test_oracle_matches_golden_r2():
    for c in _r2_cases():
        r = po.cp_solve(c["A"], c["B"], c["pi"], c["obs"], c["start"], c["comp"], int(c["ncomp"]), max_nodes=200,
                        trace_nodes=200)
        assert (r["sol"] == c["sol"]).all() and _bits(r["obj"]) == _bits(c["obj"])
        assert r["explored"] == int(c["explored"]) and r["steps"] == int(c["steps"])
        assert _bits(r["ub"][: r["explored"]]) == _bits(c["ub"])


@pytest.mark.parametrize("house", "ABC")
def test_oracle_matches_golden_ar(house):
    z = np.load(os.path.join(G, f"ar_house_{house}.npz"))
    p, s = po.decode_batch(z["logA"], z["logB"], z["obs"], z["seq_off"])
    assert (p == z["paths"]).all() and _bits(s) == _bits(z["scores"])
    r = po.cp_solve(z["logA"], z["logB"], z["logPi"], z["obs"], z["cp_start"], z["cp_comp"], int(z["cp_ncomp"]),
                    max_nodes=int(z["cp_max_nodes"]))
    assert (r["sol"] == z["cp_sol"]).all() and _bits(r["obj"]) == _bits(z["cp_obj"])
    assert r["explored"] == int(z["cp_explored"]) and r["steps"] == int(z["cp_steps"])


# ---------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_matches_golden_r1():
    import consistent_viterbi_b200 as cv
    L = cv._lib.lib()
    for c in _r1_cases():
        h = cv.HMM(c["A"], c["B"], c["pi"])
        try:
            for chain_max in (-1, 0):           # warp-per-sequence kernel, then the tile kernel
                L.cv_debug_set_chain_max_batch(chain_max)
                p, s = cv.decode_batch(h, c["obs"], c["off"])
                assert (p == c["paths"]).all() and _bits(s) == _bits(c["scores"])
        finally:
            L.cv_debug_set_chain_max_batch(-1)
        h.close()


@pytest.mark.gpu
def test_gpu_matches_golden_r2():
    import consistent_viterbi_b200 as cv
    for c in _r2_cases():
        h = cv.HMM(c["A"], c["B"], c["pi"])
        r = cv.cp_solve_arrays(h, c["obs"], c["start"], c["comp"], int(c["ncomp"]), max_nodes=200, want_ub=200)
        assert (r["sol"] == c["sol"]).all() and _bits(r["obj"]) == _bits(c["obj"])
        assert r["explored"] == int(c["explored"]) and r["steps"] == int(c["steps"])
        assert _bits(r["ub"][: r["explored"]]) == _bits(c["ub"])
        h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("house", "ABC")
def test_gpu_matches_golden_ar(house):
    """BASELINE configs[1]: the activity-recognition data set, full batched decode + a budgeted constrained solve."""
    import consistent_viterbi_b200 as cv
    z = np.load(os.path.join(G, f"ar_house_{house}.npz"))
    h = cv.HMM(z["logA"], z["logB"], z["logPi"])
    p, s = cv.decode_batch(h, z["obs"], z["seq_off"])
    assert (p == z["paths"]).all() and _bits(s) == _bits(z["scores"])
    r = cv.cp_solve_arrays(h, z["obs"], z["cp_start"], z["cp_comp"], int(z["cp_ncomp"]), max_nodes=int(z["cp_max_nodes"]))
    assert (r["sol"] == z["cp_sol"]).all() and _bits(r["obj"]) == _bits(z["cp_obj"])
    assert r["explored"] == int(z["cp_explored"]) and r["steps"] == int(z["cp_steps"])
    h.close()
