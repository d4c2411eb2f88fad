#!/usr/bin/env python
"""bench.py -- Viterbi cells/s of the B200 hot path, one JSON line (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload pos|large] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload = BASELINE.json
configs[2], the POS-tagging shape (K=45 tags, 20k vocab, 1M sentences, avg T=25): the batched, sharding
configuration the metric "cells/s at 1/2/4/8 B200" is quoted on (configs[1], datasets/ar, is 60 sequences /
1.5e7 cells, latency bound and unshardable -- it is reported under "other" together with the constrained-decode
configs and a reduced large-state run).
N > 1: launched by torchrun, one rank per GPU, every rank decodes its own shard of 1M sentences (weak
scaling, no data-path collective), max-over-ranks timing.

value    cells/s with inputs resident in HBM (cv_decode_batch_dev on torch's current stream)
e2e      cells/s through the C ABI call a user makes (cv_decode_batch, pinned HOST buffers, H2D + D2H inside)
roofline dominant kernel (forward recurrence) vs the FP64 issue peak measured in the same run, plus HBM view
cpu_baseline  the C oracle (port of the reference loops) on this box's cores, bounded sample
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "viterbi_cells_per_s"
UNIT = "cells/s"


# ----------------------------------------------------------------------------------------------
# synthetic workloads, seeded
# ----------------------------------------------------------------------------------------------
def _log_dirichlet_rows(rng, n, m, alpha, zero_frac):
    # gamma draws normalised per row = Dirichlet(alpha); chunked to bound memory
    out = np.empty((n, m), dtype=np.float64)
    for r0 in range(0, n, 64):
        r1 = min(n, r0 + 64)
        g = rng.standard_gamma(alpha, size=(r1 - r0, m))
        g[g == 0.0] = np.finfo(np.float64).tiny
        p = g / g.sum(axis=1, keepdims=True)
        p[rng.random((r1 - r0, m)) < zero_frac] = 0.0
        with np.errstate(divide="ignore"):
            out[r0:r1] = np.log10(p)
    return out


def make_hmm(seed, K, M, alpha, zero_frac):
    rng = np.random.default_rng(seed)
    A = _log_dirichlet_rows(rng, K, K, alpha, zero_frac)
    B = _log_dirichlet_rows(rng, K, M, alpha, zero_frac)
    pi = _log_dirichlet_rows(rng, 1, K, alpha, zero_frac)[0]
    return A, B, pi


def workload_pos(rank, nseq):
    """K=45, V=20000, T = clamp(round(Gamma(2.5, 10)), 1, 200), Zipf(1.1) word ids; seed 3019 (+rank)."""
    K, M = 45, 20000
    A, B, pi = make_hmm(3019, K, M, 0.1, 0.05)
    rng = np.random.default_rng(3019 + 1000 * (rank + 1))
    lens = np.clip(np.rint(rng.gamma(2.5, 10.0, size=nseq)), 1, 200).astype(np.int64)
    off = np.zeros(nseq + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    obs = (rng.zipf(1.1, size=int(off[-1])) % M).astype(np.uint32)
    cells = float(((lens - 1) * K * K).sum())
    return dict(name="pos_K45_V20k", K=K, M=M, A=A, B=B, pi=pi, obs=obs, off=off, cells=cells,
                steps=float((lens - 1).sum()), desc=f"POS shape K=45 V=20000 B={nseq} avgT=25 (configs[2])")


def workload_large(rank, nseq, T, K=1024, M=4096):
    A, B, pi = make_hmm(3019, K, M, 0.05, 0.0)
    rng = np.random.default_rng(3019 + 1000 * (rank + 1))
    off = np.arange(nseq + 1, dtype=np.int64) * T
    obs = rng.integers(0, M, size=nseq * T).astype(np.uint32)
    cells = float(nseq) * (T - 1) * K * K
    return dict(name=f"large_K{K}", K=K, M=M, A=A, B=B, pi=pi, obs=obs, off=off, cells=cells,
                steps=float(nseq) * (T - 1), desc=f"large-state K={K} M={M} T={T} B={nseq} (configs[3])")


def workload_ar():
    """BASELINE configs[1]: datasets/ar houses A/B/C as committed fixtures (tests/golden/ar_house_*.npz, made by
    tools/make_golden.py from the reference's CSVs; the reference has no preprocessing of its own)."""
    out = []
    for hname in "ABC":
        z = np.load(os.path.join(ROOT, "tests", "golden", f"ar_house_{hname}.npz"))
        K = int(z["n_activities"])
        off = z["seq_off"]
        out.append(dict(name=f"ar_house_{hname}", K=K, M=int(z["n_sensors"]), A=z["logA"], B=z["logB"], pi=z["logPi"],
                        obs=z["obs"], off=off, cells=float(((np.diff(off) - 1) * K * K).sum()), golden=z["paths"]))
    return out


def workload_cp(kind):
    """Constrained decode inputs, sampled from the model itself so that the constraints are satisfiable (as in the
    reference's pipeline, where control tags are true tags): hidden paths and observations are drawn from the HMM,
    four states are "control" states and a position carrying one of them is tagged with that state with some
    probability; component = tag value (Constraints::from_tags), prop = 1.
    trucks: BASELINE configs[0] stand-in (real datasets/trucks is absent): D=2 bdims [16,8] (M=128), K=12, 200
    sequences T~U[50,400], ~10 % of the positions clamped.  (With zero probabilities in the model the reference's
    bound -- stale backpointers into fresh rows, SURVEY Q5 -- is -inf at depth 2 and the search ends after 2K nodes.)
    heavy: configs[4]: K=16, M=64, 64 sequences x T=10000, ~20 % of the positions clamped."""
    rng = np.random.default_rng(3019)
    if kind == "trucks":
        K, M, nseq, pact, tlo, thi, zf = 12, 128, 200, 0.30, 50, 400, 0.0
    elif kind == "heavy10":                                   # configs[4] with ten times the sequences (N = 6.4M)
        K, M, nseq, pact, tlo, thi, zf = 16, 64, 640, 0.80, 10000, 10000, 0.0
    else:
        K, M, nseq, pact, tlo, thi, zf = 16, 64, 64, 0.80, 10000, 10000, 0.0
    A, B, pi = make_hmm(3019, K, M, 0.5, zf)
    PA, PB, Ppi = 10.0 ** A, 10.0 ** B, 10.0 ** pi
    PA /= PA.sum(1, keepdims=True); PB /= PB.sum(1, keepdims=True); Ppi /= Ppi.sum()
    cA, cB = np.cumsum(PA, axis=1), np.cumsum(PB, axis=1)
    lens = rng.integers(tlo, thi + 1, size=nseq)
    N = int(lens.sum())
    start = np.zeros(N, dtype=np.uint8)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    start[starts] = 1
    states = np.zeros(N, dtype=np.int64)
    u = rng.random(N)
    cur = np.minimum((np.cumsum(Ppi)[None, :] < u[starts][:, None]).sum(1), K - 1)   # first state of every sequence
    pos = starts.copy()
    alive = np.ones(nseq, dtype=bool)
    for t in range(int(lens.max())):                       # all sequences advance together (vectorised over sequences)
        idx = pos[alive]
        states[idx] = cur[alive]
        nxt = np.minimum((cA[cur[alive]] < u[np.minimum(idx + 1, N - 1)][:, None]).sum(1), K - 1)
        cur[alive] = nxt
        pos[alive] += 1
        alive = (pos - starts) < lens
        if not alive.any():
            break
    obs = np.minimum((cB[states] < rng.random(N)[:, None]).sum(1), M - 1).astype(np.uint32)
    control = rng.choice(K, size=4, replace=False)
    comp = np.full(N, -1, dtype=np.int32)
    tagged = rng.random(N) < pact
    for c, st in enumerate(control):
        comp[(states == st) & tagged] = c
    used = sorted(set(int(c) for c in comp if c >= 0))
    remap = {c: i for i, c in enumerate(used)}
    comp = np.array([remap[int(c)] if c >= 0 else -1 for c in comp], dtype=np.int32)
    return dict(name=f"cp_{kind}", K=K, M=M, A=A, B=B, pi=pi, obs=obs, start=start, comp=comp, ncomp=len(used), N=N,
                clamped=float((comp >= 0).mean()))


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi fields via NVML) during the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max = None

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.ok:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------
# CPU baseline: the C oracle (port of the reference loops), bounded sample
# ----------------------------------------------------------------------------------------------
def cpu_baseline(wl, budget_s=12.0, threads=None):
    from oracle import pyoracle as po

    threads = threads or (os.cpu_count() or 1)
    off, obs, K = wl["off"], wl["obs"], wl["K"]
    B = len(off) - 1

    def run(nb):
        o = off[: nb + 1]
        t0 = time.perf_counter()
        po.decode_batch(wl["A"], wl["B"], obs[: o[-1]], o, nthreads=threads)
        dt = time.perf_counter() - t0
        return float(((np.diff(o) - 1) * K * K).sum()), dt

    nb = max(1, min(B, 64 if K > 64 else 2000))
    if K > 64:
        # one long sequence already costs seconds on the CPU: bound T as well
        T = int(off[1] - off[0])
        Tc = min(T, 64)
        o = np.arange(min(B, threads) + 1, dtype=np.int64) * Tc
        ob = np.concatenate([obs[off[b]: off[b] + Tc] for b in range(len(o) - 1)])
        t0 = time.perf_counter()
        po.decode_batch(wl["A"], wl["B"], ob, o, nthreads=threads)
        dt = time.perf_counter() - t0
        cells = float(len(o) - 1) * (Tc - 1) * K * K
        return {"value": cells / dt, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{len(o) - 1} sequences x first {Tc} steps of the workload, {dt:.1f} s, C oracle (port of "
                          f"viterbi.rs:5-32), {threads} OpenMP threads"}
    cells, dt = run(nb)
    nb2 = int(max(nb, min(B, nb * budget_s / max(dt, 1e-3))))
    tot_c, tot_t, reps = 0.0, 0.0, 0
    while tot_t < budget_s and reps < 8:          # ~10-30 s of CPU work in total
        cells, dt = run(nb2)
        tot_c += cells; tot_t += dt; reps += 1
    return {"value": tot_c / tot_t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {nb2} sequences of the workload x {reps} passes, {tot_t:.1f} s, C oracle (port of "
                      f"viterbi.rs:5-32), {threads} OpenMP threads"}


# ----------------------------------------------------------------------------------------------
def run_other(cv, L, device):
    """Short, bounded runs of the other BASELINE configs (reported under "other"; parity for each is in tests/)."""
    from oracle import pyoracle as po
    other = {}
    # configs[1]: datasets/ar, full data set, batched decode (three models, 60 day-sequences in total)
    ar, t_gpu, t_cpu, cells = workload_ar(), 0.0, 0.0, 0.0
    for w in ar:
        hm = cv.HMM(w["A"], w["B"], w["pi"])
        cv.decode_batch(hm, w["obs"], w["off"], device=device)
        t0 = time.perf_counter()
        for _ in range(5):
            paths, _ = cv.decode_batch(hm, w["obs"], w["off"], device=device)
        t_gpu += (time.perf_counter() - t0) / 5
        assert (paths == w["golden"]).all()
        t0 = time.perf_counter()
        po.decode_batch(w["A"], w["B"], w["obs"], w["off"], nthreads=os.cpu_count() or 1)
        t_cpu += time.perf_counter() - t0
        cells += w["cells"]
        hm.close()
    other["ar_full_dataset"] = {"cells": cells, "e2e_ms": 1e3 * t_gpu, "e2e_cells_per_s": cells / t_gpu,
                                "cpu_port_cells_per_s": cells / t_cpu,
                                "note": "60 sequences, 1.5e7 cells: latency bound (serial in t), paths equal the golden fixture"}
    # configs[3] shape at reduced batch/length (the full B=4096, T=4096 run is `--workload large`: 3.95 s/step)
    w = workload_large(0, 2048, 64)
    hm = cv.HMM(w["A"], w["B"], w["pi"])
    cv.decode_batch(hm, w["obs"], w["off"], device=device)
    t0 = time.perf_counter()
    cv.decode_batch(hm, w["obs"], w["off"], device=device)
    dt = time.perf_counter() - t0
    other["large_K1024_B2048_T64"] = {"cells": w["cells"], "e2e_ms": 1e3 * dt, "e2e_cells_per_s": w["cells"] / dt,
                                      "note": "full configs[3] (B=4096, T=4096, `--workload large`): 3.95 s/step = 4.45e12 cells/s, 52 % of the FP64 roofline (DESIGN.md)"}
    hm.close()
    # supervised MLE event counts (hmm.rs:35-48) at the POS shape, 1M sentences, random tags
    w = workload_pos(0, 1000000)
    tg = np.random.default_rng(3019).integers(0, w["K"], len(w["obs"])).astype(np.int32)
    hm = cv.HMM.new(w["K"], (w["B"].shape[1],))
    hm.mle_arrays(w["obs"], tg, w["off"], device=device)
    hm2 = cv.HMM.new(w["K"], (w["B"].shape[1],))
    t0 = time.perf_counter()
    cms = hm2.mle_arrays(w["obs"], tg, w["off"], device=device)
    dt = time.perf_counter() - t0
    nel = len(w["obs"])
    other["mle_counts_pos_1M"] = {"elements": nel, "count_kernels_ms": cms, "count_GBps": 10.0 * nel / (cms * 1e-3) / 1e9,
                                    "e2e_ms": 1e3 * dt, "bytes_per_element": 10,
                                    "note": "cv_mle: device event counts (u64 atomics; 10 B/element: obs u32, tag i32, start flag written + read) + host replay of the "
                                            "reference's += 1.0 / divide / ln(x)/ln(10); e2e includes H2D of obs+tags and the K*M finalisation"}
    # CFN cost tables (cfn.rs:82-167) on the configs[4] super-sequence; CPU port on a prefix
    w = workload_cp("heavy")
    hm = cv.HMM(w["A"], w["B"], w["pi"])
    cv.cfn_tables(hm, w["obs"][:5000], w["start"][:5000], w["comp"][:5000], w["ncomp"], device=device)
    t0 = time.perf_counter()
    r = cv.cfn_tables(hm, w["obs"], w["start"], w["comp"], w["ncomp"], device=device)
    dt = time.perf_counter() - t0
    act = np.flatnonzero(w["comp"] >= 0)
    chg = np.concatenate([[True], w["comp"][act][1:] != w["comp"][act][:-1]])
    bnd = act[chg]
    K = w["K"]
    cells = float(np.diff(bnd).sum()) * K ** 3
    npre = 40000
    t0 = time.perf_counter()
    rc = po.cfn_tables(w["A"], w["B"], w["pi"], w["obs"][:npre], w["start"][:npre], w["comp"][:npre], w["ncomp"])
    dtc = time.perf_counter() - t0
    bpre = bnd[bnd < npre]
    other["cfn_tables_heavy"] = {"N": w["N"], "K": K, "boundaries": int(r["nboundaries"]), "cells": cells,
                                 "device_ms": r["device_ms"], "device_cells_per_s": cells / (r["device_ms"] * 1e-3),
                                 "e2e_ms": 1e3 * dt, "cpu_port_cells_per_s": float(np.diff(bpre).sum()) * K ** 3 / dtc * 1.0,
                                 "cpu_sample": f"first {npre} elements, single thread; the literal port runs K*K sweeps per pair "
                                               "(K-fold the device's work for the same tables), cells counted as K sweeps per pair",
                                 "lower_bound": r["lower_bound"]}
    hm.close()
    # configs[0] stand-in and configs[4]: constrained decode with a node budget
    for kind, budget, cpu_budget in (("trucks", 0, 150), ("heavy", 2000, 6)):    # trucks-like: complete search
        w = workload_cp(kind)
        hm = cv.HMM(w["A"], w["B"], w["pi"])
        cv.cp_solve_arrays(hm, w["obs"], w["start"], w["comp"], w["ncomp"], max_nodes=3, device=device)
        L.cv_set_timing(1)
        t0 = time.perf_counter()
        r = cv.cp_solve_arrays(hm, w["obs"], w["start"], w["comp"], w["ncomp"], max_nodes=budget, device=device)
        dt = time.perf_counter() - t0
        loop_ms = L.cv_last_kernel_ms(hm.device_handle(device))
        L.cv_set_timing(0)
        t0 = time.perf_counter()
        rc = po.cp_solve(w["A"], w["B"], w["pi"], w["obs"], w["start"], w["comp"], w["ncomp"], max_nodes=cpu_budget)
        dtc = time.perf_counter() - t0
        K = w["K"]
        other[w["name"]] = {"N": w["N"], "K": K, "nodes": int(r["explored"]), "sweep_steps": int(r["steps"]),
                            "cells": float(r["steps"]) * K * K, "e2e_ms": 1e3 * dt,
                            "e2e_cells_per_s": float(r["steps"]) * K * K / dt, "ms_per_node": 1e3 * dt / max(1, r["explored"]),
                            "device_loop_ms_per_node": loop_ms / max(1, r["explored"]),
                            "cpu_port_cells_per_s": float(rc["steps"]) * K * K / dtc, "cpu_nodes": int(rc["explored"]),
                            "cpu_ms_per_node": 1e3 * dtc / max(1, rc["explored"]), "objective": r["obj"],
                            "max_nodes": budget, "clamped_fraction": w["clamped"],
                            "note": "max_nodes = 0: complete branch and bound; the CPU port (single thread, like the "
                                    "reference) runs a node-budgeted prefix of the same search"}
        hm.close()
    return other


def run_cp_sharded(cv, L, spec, local, rank, world, dist, torch):
    """Constrained decode sharded over the ranks (NVLink peer stores, no NCCL on the data path) next to the same
    solve on one GPU; every rank must return the single-GPU result bit for bit.  Device-loop times (CUDA events,
    set-up and copies excluded), max over ranks."""
    kind, _, budget = spec.partition(":")
    budget = int(budget or 0)
    w = workload_cp(kind)
    hm = cv.HMM(w["A"], w["B"], w["pi"])
    args = (w["obs"], w["start"], w["comp"], w["ncomp"])
    h = hm.device_handle(local)
    L.cv_set_timing(1)
    cv.cp_solve_arrays(hm, *args, max_nodes=3, device=local)
    single = cv.cp_solve_arrays(hm, *args, max_nodes=budget, device=local)
    ms_single = L.cv_last_kernel_ms(h)
    grp = cv.CpDistGroup(hm, cap_N=w["N"], cap_terms=int((w["comp"] >= 0).sum()), device=local)
    grp.solve(*args, max_nodes=3)
    dist.barrier()
    t0 = time.perf_counter()
    r = grp.solve(*args, max_nodes=budget)
    wall = time.perf_counter() - t0
    ms_sharded = L.cv_last_kernel_ms(h)
    L.cv_set_timing(0)
    same = bool((r["sol"] == single["sol"]).all() and r["obj"] == single["obj"] and r["explored"] == single["explored"])
    t = torch.tensor([ms_sharded, ms_single, 1e3 * wall, 0.0 if same else 1.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cuts = cv.plan_cuts(w["comp"], world).tolist()
    grp.close()
    hm.close()
    nodes = max(1, int(r["explored"]))
    return {"workload": w["name"], "N": w["N"], "K": w["K"], "nodes": nodes, "row_cuts": cuts,
            "identical_to_single_gpu_on_every_rank": bool(t[3].item() == 0.0),
            "sharded_ms_per_node": t[0].item() / nodes, "single_gpu_ms_per_node": t[1].item() / nodes,
            "sharded_wall_ms": t[2].item(),
            "exchange": "bound terms, backtrack maps and solution rows by NVLink peer stores + flags (CUDA IPC); no NCCL"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pos", choices=["pos", "large"])
    ap.add_argument("--nseq", type=int, default=0, help="sequences per GPU (0 = the config's size)")
    ap.add_argument("--seqlen", type=int, default=0, help="T for --workload large (0 = 4096)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-other", action="store_true", help="skip the short runs of the other configs")
    ap.add_argument("--cp-sharded", default="", metavar="KIND:NODES",
                    help="with --gpus > 1: also time the constrained decode sharded over the ranks (csrc/cp_dist.cuh), "
                         "e.g. heavy:2000 (configs[4]) or trucks:0; reported under \"cp_sharded\"")
    return ap.parse_args()


def build_workload(args, rank):
    if args.workload == "pos":
        return workload_pos(rank, args.nseq or 1_000_000)
    return workload_large(rank, args.nseq or 4096, args.seqlen or 4096)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (C oracle port; the Rust binary cannot be
    built here) with all host threads, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import pyoracle as po

    wl = build_workload(args, 0)
    threads = os.cpu_count() or 1
    K, off, obs = wl["K"], wl["off"], wl["obs"]
    if K > 64:
        Tc, nb = 32, min(len(off) - 1, threads)
        o = np.arange(nb + 1, dtype=np.int64) * Tc
        ob = np.concatenate([obs[off[b]: off[b] + Tc] for b in range(nb)])
        sample = f"{nb} sequences x first {Tc} steps per step"
    else:
        nb = min(len(off) - 1, 20000)
        o = off[: nb + 1]
        ob = obs[: o[-1]]
        sample = f"first {nb} sequences per step"
    cells = float(((np.diff(o) - 1) * K * K).sum())
    for _ in range(args.warmup):
        po.decode_batch(wl["A"], wl["B"], ob, o, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        po.decode_batch(wl["A"], wl["B"], ob, o, nthreads=threads)
    dt = time.perf_counter() - t0
    v = cells * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "desc": wl["desc"]},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample + f", C oracle (port of viterbi.rs:5-32), {threads} OpenMP threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import consistent_viterbi_b200 as cv

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = cv._lib.lib()
    if os.environ.get("CV_CHUNKS"):
        L.cv_debug_set_chunks(int(os.environ["CV_CHUNKS"]))
    if os.environ.get("CV_SMALL_CFG"):
        L.cv_debug_set_small_config(int(os.environ["CV_SMALL_CFG"]))
    wl = build_workload(args, rank)
    hmm = cv.HMM(wl["A"], wl["B"], wl["pi"])
    h = hmm.device_handle(local)
    obs_np, off_np = wl["obs"], wl["off"]
    B, N = len(off_np) - 1, int(off_np[-1])
    max_len = int(np.diff(off_np).max())

    # ---- FP64 issue peak of this GPU, measured now (roofline denominator) ----
    ops, ms = C.c_double(), C.c_double()
    cv._lib.check(L.cv_debug_probe_fp64(local, 0, 20000, C.byref(ops), C.byref(ms)))
    peak_dadd = ops.value
    cv._lib.check(L.cv_debug_probe_fp64(local, 1, 20000, C.byref(ops), C.byref(ms)))
    peak_mix = ops.value

    # ---- device-resident inputs ----
    d_obs = torch.from_numpy(obs_np.view(np.int32)).cuda()
    d_off = torch.from_numpy(off_np).cuda()
    d_path = torch.empty(N, dtype=torch.int32, device="cuda")
    d_score = torch.empty(B, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()

    def step_dev(timing=False):
        rc = L.cv_decode_batch_dev(h, d_obs.data_ptr(), d_off.data_ptr(), B, N, max_len, d_path.data_ptr(),
                                   d_score.data_ptr(), stream.cuda_stream, 0)
        cv._lib.check(rc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_dev()
    barrier()
    launches0 = L.cv_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_dev()
        e1.record(stream)
        barrier()
    launches = L.cv_launch_count() - launches0
    dev_ms = e0.elapsed_time(e1)

    # ---- dominant-kernel duration: CUDA events on the launching stream around the forward kernel ----
    L.cv_set_timing(1)
    fwd_ms, bt_ms = [], []
    for _ in range(max(3, min(args.steps, 5))):
        rc = L.cv_decode_batch_dev(h, d_obs.data_ptr(), d_off.data_ptr(), B, N, max_len, d_path.data_ptr(),
                                   d_score.data_ptr(), stream.cuda_stream, 1)
        cv._lib.check(rc)
        fwd_ms.append(L.cv_last_kernel_ms(h))
        bt_ms.append(L.cv_last_backtrace_ms(h))
    L.cv_set_timing(0)
    fwd = float(np.mean(fwd_ms))

    # ---- end to end through the C ABI with pinned host buffers ----
    def pinned(arr):
        p = L.cv_host_alloc(arr.nbytes)
        assert p, "cv_host_alloc failed"
        buf = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(arr.nbytes,))
        buf[:] = arr.view(np.uint8).reshape(-1)
        return p, buf
    p_obs, _ = pinned(obs_np)
    p_off, _ = pinned(off_np)
    p_path = L.cv_host_alloc(4 * N)
    p_score = L.cv_host_alloc(8 * B)

    def step_e2e():
        cv._lib.check(L.cv_decode_batch(h, p_obs, p_off, B, p_path, p_score))
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(2, min(args.steps, 5))
    for _ in range(n_e2e):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / n_e2e
    # check the end-to-end result against the device-resident one
    host_paths = np.ctypeslib.as_array(C.cast(p_path, C.POINTER(C.c_uint32)), shape=(N,))
    assert (host_paths == d_path.cpu().numpy().view(np.uint32)).all(), "e2e and device-resident paths differ"

    # ---- reduce over ranks ----
    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    cp_sharded = None
    if args.cp_sharded and world > 1:
        try:
            cp_sharded = run_cp_sharded(cv, L, args.cp_sharded, local, rank, world, dist, torch)
        except Exception as e:  # noqa: BLE001
            cp_sharded = {"error": repr(e)}

    total_cells = allsum(wl["cells"])
    dev_ms = allmax(dev_ms)
    e2e_s = allmax(e2e_s)
    launches = int(allsum(float(launches)))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        K = wl["K"]
        # algorithmic HBM bytes per step of one sequence (DESIGN.md "Roofline"): obs u32 + logB^T row +
        # delta-history row written by the forward kernel and read back by the backtrace + path u32
        if K <= 64:
            bytes_per_step = 4 + 8 * K + 8 * K + 8 * K + 4
        else:
            w = 1 if ((K + 127) // 128) * 128 <= 256 else 2
            bytes_per_step = 4 + 8 * K + w * K + w + 4
        fp64_ops = 2.0 * wl["cells"]                      # 1 DADD + 1 compare per cell
        achieved_alu = fp64_ops / (fwd * 1e-3)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wl["name"])
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": total_cells * args.steps / (dev_ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "desc": wl["desc"], "per_gpu_sequences": B, "per_gpu_elements": N,
                       "sharding": f"{world} x independent shards, no data-path collective",
                       "l2": "inputs + delta history per step >> 126 MB L2 (no flush needed)"},
            "e2e": {"value": total_cells / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(obs_np.nbytes + off_np.nbytes),
                    "d2h_bytes_per_step": int(4 * N + 8 * B), "ms_per_step": 1e3 * e2e_s,
                    "api": "cv_decode_batch (C ABI, pinned host buffers)"},
            "gpu_launches": launches,
            "clocks": clk.summary(),
            "roofline": {
                "bound": "fp64_alu", "kernel": "decode_small_fwd_kernel" if K <= 64 else "decode_large_kernel",
                "achieved": achieved_alu / 1e12, "peak": peak_mix / 1e12, "unit": "TFLOP/s",
                "unit_note": "FP64 add+compare operations (2 per cell), not tensor FLOPs; frac = max(ALU view, HBM view) as SURVEY 8d defines",
                "frac": achieved_alu / peak_mix, "traffic": traffic,
                "peak_source": "measured in this run: DADD+DSETP issue rate over all SMs (cv_debug_probe_fp64 mode 1); "
                               f"DADD alone {peak_dadd / 1e12:.2f}",
                "kernel_ms": fwd, "backtrace_ms": float(np.mean(bt_ms)),
                "hbm": {"achieved": wl["steps"] * bytes_per_step / ((fwd + float(np.mean(bt_ms))) * 1e-3) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s", "bytes_per_step": bytes_per_step,
                        "frac": wl["steps"] * bytes_per_step / ((fwd + float(np.mean(bt_ms))) * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
            },
        }
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline(wl)
        if not args.no_other and world == 1 and args.workload == "pos":
            try:
                line["other"] = run_other(cv, L, local)
            except Exception as e:  # the headline numbers stand on their own
                line["other"] = {"error": repr(e)}
        if cp_sharded is not None:
            line["cp_sharded"] = cp_sharded
        print(json.dumps(line), flush=True)
    for p in (p_obs, p_off, p_path, p_score):
        L.cv_host_free(p)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
