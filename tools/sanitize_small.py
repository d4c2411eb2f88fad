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
L.cv_debug_set_chain_max_batch(0)
p2, s2 = cv.decode_batch(h, obs, off)             # tile kernel + backtrace
L.cv_debug_set_chain_max_batch(-1)
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
# second session: block-structured ordered sum (all three paths), CFN tables, MLE counts, K <= 16 half-warp sweeps
import ctypes as C
def osum(x, mode):
    out = C.c_double(0.0)
    cv._lib.check(L.cv_debug_ordered_sum(np.ascontiguousarray(x).ctypes.data, len(x), mode, C.byref(out)))
    return out.value
x = -rng.random(70000) * 3
assert osum(x, 0) == osum(x, 1) == osum(x, 2) == float(np.cumsum(x)[-1])
x = -rng.random(128 * 4096 + 777)
assert osum(x, 1) == float(np.cumsum(x)[-1])
A, B, pi = random_hmm(rng, 12, 9, zero_frac=0.05)
h = cv.HMM(A, B, pi)
o2, st, comp, nc = random_superseq(rng, 40, 9, 3, 0.25, 3, 30)
r2 = cv.cp_solve_arrays(h, o2, st, comp, nc, max_nodes=200)
c = cv.cfn_tables(h, o2, st, comp, nc)
h.close()
obs, off = random_batch(rng, 500, 9, 1, 20)
tags = rng.integers(0, 12, len(obs)).astype(np.int32)
cv.HMM.new(12, (9,), rng).mle_arrays(obs, tags, off)
cv.HMM.new(70, (9,)).mle_arrays(obs, rng.integers(0, 70, len(obs)).astype(np.int32), off)
print("sanitize run 2 ok", r2["explored"], c["nboundaries"])
