"""Repeated runs of the streamed / concurrent decode pipeline on a ragged batch; every run must reproduce the oracle
result bit for bit (memory-ordering bugs in the tile_done / arrived / chunk_done hand-offs would show up here)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import consistent_viterbi_b200 as cv
from oracle import pyoracle as po
from util import random_hmm
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(7)
K, M, Bn = 45, 500, 300000
A, B, pi = random_hmm(rng, K, M, zero_frac=0.05)
lens = np.clip(np.rint(rng.gamma(2.5, 10.0, size=Bn)), 1, 200).astype(np.int64)
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
obs = rng.integers(0, M, int(off[-1])).astype(np.uint32)
t0 = time.time()
rp, rs = po.decode_batch(A, B, obs, off, nthreads=os.cpu_count() or 8)
print("oracle", round(time.time() - t0, 1), "s")
h = cv.HMM(A, B, pi)
L = cv._lib.lib()
bad = 0
for it in range(iters):
    L.cv_debug_set_chunks((2, 3, 4, 7)[it % 4])
    p, s = cv.decode_batch(h, obs, off)
    if not ((p == rp).all() and s.tobytes() == rs.tobytes()):
        bad += 1
        print("MISMATCH at iteration", it, "paths differing:", int((p != rp).sum()))
L.cv_debug_set_chunks(-1)
print("iterations", iters, "mismatches", bad)
sys.exit(1 if bad else 0)
