"""The bound of a B&B node is a serial f64 sum whose order is part of the result (cp.rs:103-116).  The GPU path
has a one-thread loop and a parallel exact-order kernel (binade scan, cp_kernels.cuh); both must equal the
sequential IEEE sum bit for bit -- including exact ties, zeros, -inf and the serial fall-backs."""
import ctypes as C

import numpy as np
import pytest

import consistent_viterbi_b200 as cv

pytestmark = pytest.mark.gpu


def _serial(x):
    s = 0.0
    for v in x.tolist():
        s = s + v
    return s


def _gpu(x, mode):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = C.c_double(0.0)
    cv._lib.check(cv._lib.lib().cv_debug_ordered_sum(x.ctypes.data, len(x), mode, C.byref(out)))
    return out.value


def _same(a, b):
    return np.float64(a).tobytes() == np.float64(b).tobytes() or (np.isnan(a) and np.isnan(b))


@pytest.mark.parametrize("seed", range(6))
def test_ordered_sum_random(seed):
    rng = np.random.default_rng(900 + seed)
    for it in range(40):
        n = int(rng.integers(0, 30000))
        kind = it % 5
        if kind == 0:
            x = -rng.random(n) * 10                               # log-probability-like
        elif kind == 1:
            x = -rng.integers(0, 64, n) / 8.0                     # exact ties all the time
        elif kind == 2:
            x = -np.ldexp(rng.random(n), rng.integers(-60, 20, n))   # wildly different magnitudes
        elif kind == 3:
            x = -(rng.integers(0, 1 << 20, n) * 2.0 ** -rng.integers(0, 60, n).astype(np.float64))
        else:
            x = np.log10(np.maximum(rng.dirichlet(np.full(8, 0.3), n)[:, 0], 1e-300)) if n else np.zeros(0)
        if n and it % 7 == 0:
            x[rng.integers(0, n)] = -0.0
        ref = _serial(x)
        assert _same(_gpu(x, 0), ref), f"serial kernel differs (kind {kind}, n {n})"
        assert _same(_gpu(x, 1), ref), f"block-structured kernel differs (kind {kind}, n {n})"
        assert _same(_gpu(x, 2), ref), f"single-CTA scan kernel differs (kind {kind}, n {n})"


def test_ordered_sum_large_and_special():
    rng = np.random.default_rng(77)
    x = np.log10(rng.random(300000))                              # 3e5 terms like a config-5 node
    ref = float(np.cumsum(x)[-1])                                 # numpy cumsum is the sequential sum
    assert _same(_gpu(x, 1), ref) and _same(_gpu(x, 0), ref) and _same(_gpu(x, 2), ref)
    y = x[:5000].copy(); y[1234] = -np.inf
    assert _gpu(y, 1) == -np.inf and _gpu(y, 0) == -np.inf
    z = x[:5000].copy(); z[77] = 3.5                              # positive term -> serial fall-back inside the kernel
    assert _same(_gpu(z, 1), _serial(z)) and _same(_gpu(z, 2), _serial(z))
    z2 = x[:5000].copy(); z2[4999] = np.nan
    assert np.isnan(_gpu(z2, 1)) and np.isnan(_gpu(z2, 0))
    tiny = -np.full(4000, 5e-324)                                 # subnormal running sum -> fall-back
    assert _same(_gpu(tiny, 1), _serial(tiny)) and _same(_gpu(tiny, 2), _serial(tiny))
    assert _gpu(np.zeros(0), 1) == 0.0 and _gpu(np.zeros(0), 0) == 0.0
    assert _same(_gpu(-np.zeros(10), 1), _serial(-np.zeros(10)))


def test_ordered_sum_block_edges():
    """Sizes around the 128-term blocks / 32-block groups of the block-structured kernel, sums sitting right at a
    binade boundary (the prediction must be distrusted there), and a list longer than SUM_MAX_BLOCKS blocks."""
    rng = np.random.default_rng(4242)
    for n in (1, 2, 127, 128, 129, 255, 256, 4095, 4096, 4097, 8191, 8193, 128 * 32 * 3 + 5):
        x = -rng.random(n) * 3
        assert _same(_gpu(x, 1), _serial(x)), n
        y = -np.ones(n) * 0.5                                     # partial sums hit powers of two exactly
        assert _same(_gpu(y, 1), _serial(y)), n
    for total in (2.0 ** 10, 2.0 ** 14):                          # creep over a binade boundary in tiny steps
        x = np.concatenate([[-(total - 1e-9)], -np.full(6000, 1e-12), -rng.random(3000)])
        assert _same(_gpu(x, 1), _serial(x))
    for n in (128 * 2048 - 5, 128 * 2048 + 5,                     # around SUM_PREFIX_MIN (prefix kernel on / off)
              128 * 4096 - 77, 128 * 4096 + 300,                  # one tile / two tiles of the chain kernel
              128 * 4096 * 3 + 5):                                # four tiles
        x = -rng.random(n) * 2
        assert _same(_gpu(x, 1), float(np.cumsum(x)[-1])), n
    x = np.log10(rng.random(128 * 65536 + 5))                     # > SUM_MAX_BLOCKS blocks -> single-CTA path
    assert _same(_gpu(x, 1), float(np.cumsum(x)[-1]))
