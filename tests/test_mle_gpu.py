"""cv_mle (device event counts + the reference's `+= 1.0` / divide / ln(x)/ln(10) arithmetic replayed on the
host) against the literal CPU restatement of HMM::maximum_likelihood_estimation + log (hmm.rs:30-62,192-205):
bit-exact for a zero initial model and for a random one (the reference starts from HMM::new's random model)."""
import numpy as np
import pytest

import consistent_viterbi_b200 as cv
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _data(rng, K, M, B, tmin, tmax, skew=False):
    lens = rng.integers(tmin, tmax + 1, size=B)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    N = int(off[-1])
    if skew:                                                       # a few hot (tag, word) pairs: heavy atomics contention
        obs = np.minimum(rng.zipf(1.3, N) - 1, M - 1).astype(np.uint32)
        tags = np.minimum(rng.zipf(1.5, N) - 1, K - 1).astype(np.int32)
    else:
        obs = rng.integers(0, M, N).astype(np.uint32)
        tags = rng.integers(0, K, N).astype(np.int32)
    return obs, tags, off


def _same(x, y):
    return np.ascontiguousarray(x).tobytes() == np.ascontiguousarray(y).tobytes()


@pytest.mark.parametrize("K,M,B,tmax", [(3, 4, 7, 5), (8, 14, 25, 60), (45, 2000, 3000, 40), (100, 50, 400, 30)])
@pytest.mark.parametrize("init", ["zeros", "random"])
def test_mle_matches_oracle(K, M, B, tmax, init):
    rng = np.random.default_rng(K * 1000 + M)
    obs, tags, off = _data(rng, K, M, B, 1, tmax, skew=(K == 45))
    h = cv.HMM.new(K, (M,), None if init == "zeros" else rng)
    a0, b0, pi0 = h.a.copy(), h.b.copy(), h.pi.copy()
    h.mle_arrays(obs, tags, off)
    ra, rb, rpi = po.mle(a0, b0.reshape(K, -1), pi0, obs, tags, off)
    assert _same(h.a, ra) and _same(h.b.reshape(K, -1), rb) and _same(h.pi, rpi)


def test_mle_many_increments_on_random_init():
    """One entry incremented ~1e6 times on top of a random value: every binade crossing of the running value
    rounds once in the reference's loop; the closed form must reproduce it."""
    rng = np.random.default_rng(99)
    K, M, B = 2, 2, 400000
    off = np.arange(0, 3 * B + 1, 3, dtype=np.int64)
    obs = np.zeros(3 * B, dtype=np.uint32)
    tags = np.zeros(3 * B, dtype=np.int32)
    tags[rng.integers(0, 3 * B, 50)] = 1
    h = cv.HMM.new(K, (M,), rng)
    a0, b0, pi0 = h.a.copy(), h.b.copy(), h.pi.copy()
    h.mle_arrays(obs, tags, off)
    ra, rb, rpi = po.mle(a0, b0, pi0, obs, tags, off)
    assert _same(h.a, ra) and _same(h.b, rb) and _same(h.pi, rpi)


def test_mle_reference_interface_and_errors():
    rng = np.random.default_rng(3)
    K, bd = 4, (3, 2)
    seqs = [rng.integers(0, [3, 2], size=(int(rng.integers(1, 9)), 2)) for _ in range(12)]
    tags = [[int(t) for t in rng.integers(0, K, len(s))] for s in seqs]
    h = cv.HMM.new(K, bd)
    h.maximum_likelihood_estimation(seqs, tags)
    flat = np.concatenate([h.flatten_obs(s) for s in seqs])
    off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])])
    ra, rb, rpi = po.mle(np.zeros((K, K)), np.zeros((K, 6)), np.zeros(K), flat, np.concatenate(tags), off)
    assert _same(h.a, ra) and _same(h.b.reshape(K, -1), rb) and _same(h.pi, rpi)
    # the trained model decodes on the GPU like any other
    p, _ = cv.decode_batch(h, flat, off)
    rp, _ = po.decode_batch(h.a, h.b.reshape(K, -1), flat, off)
    assert (p == rp).all()
    bad = [list(t) for t in tags]
    bad[3][0] = None                                              # reference: tag[t].unwrap() panics
    with pytest.raises(cv.CvError) as e:
        cv.HMM.new(K, bd).maximum_likelihood_estimation(seqs, bad)
    assert e.value.code == cv._lib.ERR_ARG
    with pytest.raises(cv.CvError) as e:                          # reference: tag[0] on an empty sequence panics
        cv.HMM.new(K, bd).mle_arrays(flat, np.concatenate(tags), np.array([0, 0, len(flat)]))
    assert e.value.code == cv._lib.ERR_EMPTY
