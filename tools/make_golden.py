"""Generate the committed fixtures under tests/golden/ (run HERE, where /root/reference exists).

1. ar_house_{A,B,C}.npz -- BASELINE configs[1]: datasets/ar/house-*.csv turned into the hot path's inputs.
   The reference ships no preprocessing for these CSVs, so the mapping is the builder's:
   state = `activity` (sorted unique), observation = `sensor` id (D=2 with a constant second feature, bdims =
   [#sensors, 1]), one sequence per `day`.  The HMM is a deterministic add-one-smoothed count MLE in log10
   (the reference's MLE adds counts to a thread_rng init, hmm.rs:22-47, and cannot be reproduced).
   Stored: obs, seq_off, tags, logA/logB/logPi and the ORACLE's decode paths/scores (golden outputs).
2. r1_small.npz / r2_small.npz -- seeded random instances with the oracle's outputs, as regression pins for
   both restatements and the GPU path.

The golden outputs come from oracle/ (C restatement); the reference itself cannot run here (Rust, no cargo).
"""
import csv
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as po  # noqa: E402
from util import random_batch, random_hmm, random_superseq  # noqa: E402

REF = "/root/reference/datasets/ar"
OUT = os.path.join(ROOT, "tests", "golden")


def mle_log10(obs, off, tags, K, M):
    """Add-one smoothed counts -> log10 probabilities (deterministic)."""
    A = np.ones((K, K)); Bm = np.ones((K, M)); pi = np.ones(K)
    for b in range(len(off) - 1):
        s = tags[off[b]:off[b + 1]]; o = obs[off[b]:off[b + 1]]
        pi[s[0]] += 1
        for t in range(len(s)):
            Bm[s[t], o[t]] += 1
            if t:
                A[s[t - 1], s[t]] += 1
    return (np.log10(A / A.sum(1, keepdims=True)), np.log10(Bm / Bm.sum(1, keepdims=True)), np.log10(pi / pi.sum()))


def house(name):
    rows = list(csv.DictReader(open(os.path.join(REF, f"house-{name}.csv"))))
    sensors = sorted({r["sensor"] for r in rows})
    acts = sorted({r["activity"] for r in rows})
    sid = {s: i for i, s in enumerate(sensors)}
    aid = {a: i for i, a in enumerate(acts)}
    obs, tags, off, last = [], [], [0], None
    for r in rows:
        if last is not None and r["day"] != last:
            off.append(len(obs))
        last = r["day"]
        obs.append(sid[r["sensor"]]); tags.append(aid[r["activity"]])
    off.append(len(obs))
    obs = np.array(obs, dtype=np.uint32); tags = np.array(tags, dtype=np.int32); off = np.array(off, dtype=np.int64)
    K, M = len(acts), len(sensors)
    A, Bm, pi = mle_log10(obs, off, tags, K, M)
    paths, scores = po.decode_batch(A, Bm, obs, off)
    # constrained variant: control tags = true activity on every 50th element, 3 most frequent activities only
    top = np.argsort(-np.bincount(tags, minlength=K))[:3]
    comp = np.full(len(obs), -1, dtype=np.int32)
    for c, a in enumerate(top):
        idx = np.nonzero(tags == a)[0][::50]
        comp[idx] = c
    used = sorted(set(int(c) for c in comp if c >= 0))
    remap = {c: i for i, c in enumerate(used)}
    comp = np.array([remap[int(c)] if c >= 0 else -1 for c in comp], dtype=np.int32)
    start = np.zeros(len(obs), dtype=np.uint8); start[off[:-1]] = 1
    cp = po.cp_solve(A, Bm, pi, obs, start, comp, len(used), max_nodes=40)
    np.savez_compressed(os.path.join(OUT, f"ar_house_{name}.npz"), obs=obs, seq_off=off, tags=tags, logA=A, logB=Bm,
                        logPi=pi, paths=paths, scores=scores, cp_comp=comp, cp_start=start, cp_ncomp=len(used),
                        cp_max_nodes=40, cp_sol=cp["sol"], cp_obj=cp["obj"], cp_explored=cp["explored"],
                        cp_steps=cp["steps"], n_sensors=M, n_activities=K)
    print(name, "N", len(obs), "seqs", len(off) - 1, "K", K, "M", M, "cells", int(((np.diff(off) - 1) * K * K).sum()),
          "cp explored", cp["explored"], "obj", cp["obj"])


def small():
    rng = np.random.default_rng(20260101)
    r1 = {}
    for i, (K, M, Bn) in enumerate([(3, 4, 40), (8, 5, 60), (17, 9, 50), (45, 50, 80), (64, 6, 30), (70, 7, 20), (130, 5, 12)]):
        A, Bm, pi = random_hmm(rng, K, M, zero_frac=0.2, ties=(i == 1))
        obs, off = random_batch(rng, Bn, M, 1, 30)
        p, s = po.decode_batch(A, Bm, obs, off)
        for k, v in dict(A=A, B=Bm, pi=pi, obs=obs, off=off, paths=p, scores=s).items():
            r1[f"c{i}_{k}"] = v
    r1["ncases"] = 7
    np.savez_compressed(os.path.join(OUT, "r1_small.npz"), **r1)
    r2 = {}
    for i, (K, M, ns, nc, pa) in enumerate([(2, 3, 4, 2, 0.3), (4, 4, 6, 3, 0.2), (6, 9, 12, 3, 0.15), (12, 20, 20, 4, 0.1),
                                             (5, 4, 5, 0, 0.0), (16, 12, 10, 2, 0.1)]):
        A, Bm, pi = random_hmm(rng, K, M, zero_frac=0.15, ties=(i == 1))
        obs, start, comp, ncomp = random_superseq(rng, ns, M, nc, pa, 2, 25)
        r = po.cp_solve(A, Bm, pi, obs, start, comp, ncomp, max_nodes=200, trace_nodes=200)
        for k, v in dict(A=A, B=Bm, pi=pi, obs=obs, start=start, comp=comp, ncomp=ncomp, sol=r["sol"], obj=r["obj"],
                         explored=r["explored"], steps=r["steps"], ub=r["ub"][: r["explored"]]).items():
            r2[f"c{i}_{k}"] = v
    r2["ncases"] = 6
    np.savez_compressed(os.path.join(OUT, "r2_small.npz"), **r2)


def stdrng():
    """The stream `recompute_constraints` consumes: StdRng::seed_from_u64(3019) (viterbi_solver/utils.rs:101).  The
    generator is pinned to the rand crates' published vectors by tests/test_cli.py; this freezes its first draws."""
    import json

    from consistent_viterbi_b200.superseq import StdRng
    a, b = StdRng(3019), StdRng(3019)
    json.dump({"seed": 3019, "next_u64": [a.next_u64() for _ in range(16)], "gen_f64_hex": [b.gen_f64().hex() for _ in range(16)]},
              open(os.path.join(OUT, "stdrng_3019.json"), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--stdrng-only" in sys.argv:
        stdrng()
        sys.exit(0)
    for n in "ABC":
        house(n)
    small()
    stdrng()
    print("golden fixtures written to", OUT)
