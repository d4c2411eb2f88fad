"""Tiny run of every kernel family (for compute-sanitizer): tile kernel, chain kernel, large-K kernel, CP solver."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import consistent_viterbi_b200 as cv
from util import random_batch, random_hmm, random_superseq
rng = np.random.default_rng(1)
L = cv._lib.lib()
A, B, pi = random_hmm(rng, 13, 7)
obs, off = random_batch(rng, 150, 7, 1, 12)
h = cv.HMM(A, B, pi)
p1, s1 = cv.decode_batch(h, obs, off)             # chain kernel
L.cv_set_chain_max_batch(0)
p2, s2 = cv.decode_batch(h, obs, off)             # tile kernel + backtrace
L.cv_set_chain_max_batch(-1)
assert (p1 == p2).all() and s1.tobytes() == s2.tobytes()
o2, st, comp, nc = random_superseq(rng, 6, 7, 2, 0.2, 3, 15)
r = cv.cp_solve_arrays(h, o2, st, comp, nc, max_nodes=30)
h.close()
A, B, pi = random_hmm(rng, 70, 5)
obs, off = random_batch(rng, 70, 5, 1, 6)
h = cv.HMM(A, B, pi)
cv.decode_batch(h, obs, off)                      # large-K kernel
h.close()
print("sanitize run ok", r["explored"])
