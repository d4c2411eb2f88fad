"""Constrained decode: mirror of the reference's `Solver` trait (src/viterbi_solver.rs:11-16) and
`CPSolver` (src/viterbi_solver/cp.rs:8-152) over the C ABI entry cv_cp_solve.  GPU only."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .hmm import HMM
from .superseq import SuperSequence


def cp_solve_arrays(hmm: HMM, obs, is_seq_start, comp, ncomp, max_nodes: int = 0, device: int = -1,
                    want_state: bool = False, want_ub: int = 0):
    """cv_cp_solve on raw arrays.  Returns dict(sol u64[N], obj, explored, steps[, delta, psi, ub])."""
    obs = np.ascontiguousarray(obs, dtype=np.uint32)
    start = np.ascontiguousarray(is_seq_start, dtype=np.uint8)
    comp = np.ascontiguousarray(comp, dtype=np.int32)
    N = obs.shape[0]
    sol = np.zeros(max(N, 1), dtype=np.uint64)
    obj = C.c_double(0.0)
    explored, steps = C.c_uint64(0), C.c_uint64(0)
    L = _lib.lib()
    h = hmm.device_handle(device)
    rc = L.cv_cp_solve(h, obs.ctypes.data, start.ctypes.data, comp.ctypes.data, N, int(ncomp), int(max_nodes),
                       sol.ctypes.data, C.byref(obj), C.byref(explored), C.byref(steps))
    _lib.check(rc)
    out = dict(sol=sol[:N], obj=obj.value, explored=explored.value, steps=steps.value)
    if want_state:
        K = hmm.nstates()
        delta = np.zeros((N, K), dtype=np.float64)
        psi = np.zeros((N, K), dtype=np.uint64)
        _lib.check(L.cv_debug_cp_last_state(h, delta.ctypes.data, psi.ctypes.data))
        out["delta"], out["psi"] = delta, psi
    if want_ub:
        ub = np.zeros(want_ub, dtype=np.float64)
        n = C.c_uint64(0)
        _lib.check(L.cv_debug_cp_last_ub(h, ub.ctypes.data, want_ub, C.byref(n)))
        out["ub"] = ub[: min(want_ub, n.value)]
    return out


def cfn_tables(hmm: HMM, obs, is_seq_start, comp, k: int, device: int = -1):
    """cv_cfn_tables: write_cfn's cost tables [k,k,K,K], unary costs [k,K], lower bound (cfn.rs:82-167)."""
    obs = np.ascontiguousarray(obs, dtype=np.uint32)
    start = np.ascontiguousarray(is_seq_start, dtype=np.uint8)
    comp = np.ascontiguousarray(comp, dtype=np.int32)
    K = hmm.nstates()
    tables = np.zeros((max(k, 0), max(k, 0), K, K), dtype=np.float64)
    unary = np.zeros((max(k, 0), K), dtype=np.float64)
    lb, ms, nb = C.c_double(0.0), C.c_double(0.0), C.c_int64(0)
    rc = _lib.lib().cv_cfn_tables(hmm.device_handle(device), obs.ctypes.data, start.ctypes.data, comp.ctypes.data,
                                  obs.shape[0], int(k), tables.ctypes.data, unary.ctypes.data, C.byref(lb), C.byref(nb),
                                  C.byref(ms))
    _lib.check(rc)
    return dict(tables=tables, unary=unary, lower_bound=lb.value, nboundaries=nb.value, device_ms=ms.value)


IPC_HANDLE_BYTES = 64


def plan_cuts(comp, nranks: int):
    """Row cuts of the sharded solve (cv_cp_plan_cuts): int64[nranks + 1]; {0, N, N, ..} = cannot be cut."""
    comp = np.ascontiguousarray(comp, dtype=np.int32)
    cuts = np.zeros(nranks + 1, dtype=np.int64)
    _lib.check(_lib.lib().cv_cp_plan_cuts(comp.ctypes.data, comp.shape[0], int(nranks), cuts.ctypes.data))
    return cuts


class CpDistGroup:
    """This rank's exchange buffer for sharded constrained solves (cv_cp_dist_*).  The IPC handles travel over
    torch.distributed (any backend); the per-node exchange itself is NVLink peer stores inside the kernels."""

    def __init__(self, hmm: HMM, cap_N: int, cap_terms: int, device: int = -1, group=None):
        import torch.distributed as dist

        self.hmm, self.cap_N, self.cap_terms = hmm, int(cap_N), int(cap_terms)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        L = _lib.lib()
        self._h = C.c_void_p()
        mine = C.create_string_buffer(IPC_HANDLE_BYTES)
        _lib.check(L.cv_cp_dist_create(hmm.device_handle(device), self.rank, self.world, self.cap_N, self.cap_terms,
                                       mine, C.byref(self._h)))
        handles = [None] * self.world
        dist.all_gather_object(handles, mine.raw, group=group)
        _lib.check(L.cv_cp_dist_connect(self._h, b"".join(handles)))
        dist.barrier(group=group)

    def solve(self, obs, is_seq_start, comp, ncomp, max_nodes: int = 0):
        """Collective cv_cp_solve_dist; every rank gets the full result."""
        obs = np.ascontiguousarray(obs, dtype=np.uint32)
        start = np.ascontiguousarray(is_seq_start, dtype=np.uint8)
        comp = np.ascontiguousarray(comp, dtype=np.int32)
        N = obs.shape[0]
        sol = np.zeros(max(N, 1), dtype=np.uint64)
        obj = C.c_double(0.0)
        explored, steps = C.c_uint64(0), C.c_uint64(0)
        rc = _lib.lib().cv_cp_solve_dist(self._h, obs.ctypes.data, start.ctypes.data, comp.ctypes.data, N, int(ncomp),
                                         int(max_nodes), sol.ctypes.data, C.byref(obj), C.byref(explored), C.byref(steps))
        _lib.check(rc)
        return dict(sol=sol[:N], obj=obj.value, explored=explored.value, steps=steps.value)

    def close(self):
        if self._h:
            _lib.lib().cv_cp_dist_destroy(self._h)
            self._h = C.c_void_p()


class Solver:
    """trait Solver (viterbi_solver.rs:11-16)."""

    def solve(self):
        raise NotImplementedError

    def get_solution(self):
        raise NotImplementedError

    def get_objective(self):
        raise NotImplementedError

    def get_name(self):
        raise NotImplementedError


class CPSolver(Solver):
    """CPSolver::new(&hmm, &super_seq) (cp.rs:20); solve / get_solution / get_objective / get_name /
    get_explored_nodes keep the reference's names and meaning."""

    def __init__(self, hmm: HMM, sequence: SuperSequence, device: int = -1, max_nodes: int = 0):
        self.hmm, self.sequence, self.device, self.max_nodes = hmm, sequence, device, max_nodes
        self.best_obj = -np.inf                                      # cp.rs:29
        self.best_sol = np.zeros(len(sequence), dtype=np.uint64)
        self.explored_nodes = 0
        self.sweep_steps = 0

    def solve(self):                                                # cp.rs:133-143
        obs, start, comp, ncomp = self.sequence.solver_inputs()
        r = cp_solve_arrays(self.hmm, obs, start, comp, ncomp, self.max_nodes, self.device)
        self.best_sol, self.best_obj = r["sol"], r["obj"]
        self.explored_nodes, self.sweep_steps = r["explored"], r["steps"]

    def get_solution(self):                                         # cp.rs:145-147
        return self.best_sol

    def get_objective(self):                                        # cp.rs:149
        return self.best_obj

    def get_name(self):                                             # cp.rs:151
        return "cp"

    def get_explored_nodes(self):                                   # cp.rs:128
        return self.explored_nodes
