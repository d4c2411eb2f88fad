"""ctypes binding of the C oracle (oracle/libcv_oracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/cv_oracle.h.  PARITY UNPINNED: the
reference has no tests or fixtures and cannot be built here (Rust, no cargo).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import
this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcv_oracle.so")

OK, ERR_EMPTY, ERR_NAN, ERR_ARG, ERR_ASSERT = 0, 1, 2, 3, 4


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc -O2, no fast-math)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
        os.path.join(_HERE, "cv_oracle.c")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, u32p, i64p, u64p, u8p, i32p = (
            C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_int64),
            C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.POINTER(C.c_int32),
        )
        L.cvo_argmax.argtypes = [dp, C.c_int64, i64p]
        L.cvo_argmax.restype = C.c_int
        L.cvo_decode.argtypes = [C.c_int, C.c_int64, dp, dp, u32p, C.c_int64, u32p, dp]
        L.cvo_decode.restype = C.c_int
        L.cvo_decode_trace.argtypes = [C.c_int, C.c_int64, dp, dp, u32p, C.c_int64, dp, u32p]
        L.cvo_decode_trace.restype = C.c_int
        L.cvo_decode_batch.argtypes = [C.c_int, C.c_int64, dp, dp, u32p, i64p, C.c_int64, u32p, dp, C.c_int]
        L.cvo_decode_batch.restype = C.c_int
        L.cvo_cp_solve.argtypes = [
            C.c_int, C.c_int64, dp, dp, dp, C.c_int64, u32p, u8p, i32p, C.c_int32, C.c_uint64,
            u64p, dp, u64p, u64p, u64p, C.c_uint64, dp, dp, u64p,
        ]
        L.cvo_cp_solve.restype = C.c_int
        L.cvo_mle.argtypes = [C.c_int, C.c_int64, dp, dp, dp, u32p, i32p, i64p, C.c_int64]
        L.cvo_mle.restype = C.c_int
        L.cvo_cfn_tables.argtypes = [C.c_int, C.c_int64, dp, dp, dp, C.c_int64, u32p, u8p, i32p, C.c_int32, dp, dp, dp, i64p]
        L.cvo_cfn_tables.restype = C.c_int
        _lib = L
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle status {code}")
        self.code = code


def argmax(v):
    """ndarray-stats 0.5 `QuantileExt::argmax` of a 1-D f64 vector (OracleError on empty input / NaN)."""
    v = _f64(np.asarray(v, dtype=np.float64).reshape(-1))
    idx = C.c_int64(0)
    rc = lib().cvo_argmax(_p(v, C.c_double), v.shape[0], C.byref(idx))
    if rc:
        raise OracleError(rc)
    return int(idx.value)


def decode(logA, logB, obs):
    """viterbi::decode (viterbi.rs:5-32). Returns (path u32[T], score)."""
    logA, logB = _f64(logA), _f64(logB)
    K, M = logB.shape
    obs = np.ascontiguousarray(obs, dtype=np.uint32)
    T = obs.shape[0]
    path = np.zeros(max(T, 1), dtype=np.uint32)
    score = C.c_double(0.0)
    rc = lib().cvo_decode(K, M, _p(logA, C.c_double), _p(logB, C.c_double), _p(obs, C.c_uint32), T,
                          _p(path, C.c_uint32), C.byref(score))
    if rc:
        raise OracleError(rc)
    return path[:T], score.value


def decode_trace(logA, logB, obs):
    logA, logB = _f64(logA), _f64(logB)
    K, M = logB.shape
    obs = np.ascontiguousarray(obs, dtype=np.uint32)
    T = obs.shape[0]
    delta = np.zeros((T, K), dtype=np.float64)
    psi = np.zeros((T, K), dtype=np.uint32)
    rc = lib().cvo_decode_trace(K, M, _p(logA, C.c_double), _p(logB, C.c_double), _p(obs, C.c_uint32), T,
                                _p(delta, C.c_double), _p(psi, C.c_uint32))
    if rc:
        raise OracleError(rc)
    return delta, psi


def decode_batch(logA, logB, obs_flat, seq_off, nthreads: int = 1):
    """B independent viterbi::decode calls. Returns (paths u32[N], scores f64[B])."""
    logA, logB = _f64(logA), _f64(logB)
    K, M = logB.shape
    obs_flat = np.ascontiguousarray(obs_flat, dtype=np.uint32)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
    B = seq_off.shape[0] - 1
    paths = np.zeros(max(obs_flat.shape[0], 1), dtype=np.uint32)
    scores = np.zeros(max(B, 1), dtype=np.float64)
    rc = lib().cvo_decode_batch(K, M, _p(logA, C.c_double), _p(logB, C.c_double), _p(obs_flat, C.c_uint32),
                                _p(seq_off, C.c_int64), B, _p(paths, C.c_uint32), _p(scores, C.c_double),
                                int(nthreads))
    if rc:
        raise OracleError(rc)
    return paths[: obs_flat.shape[0]], scores[:B]


def cp_solve(logA, logB, logPi, obs, is_seq_start, comp, ncomp, max_nodes: int = 0,
             trace_nodes: int = 0, want_state: bool = False):
    """CPSolver::new + solve (cp.rs:20-152).

    Returns dict(sol u64[N], obj, explored, steps, node_hash, ub, delta, psi).
    """
    logA, logB, logPi = _f64(logA), _f64(logB), _f64(logPi)
    K, M = logB.shape
    obs = np.ascontiguousarray(obs, dtype=np.uint32)
    N = obs.shape[0]
    start = np.ascontiguousarray(is_seq_start, dtype=np.uint8)
    comp = np.ascontiguousarray(comp, dtype=np.int32)
    sol = np.zeros(max(N, 1), dtype=np.uint64)
    obj = C.c_double(0.0)
    explored, steps = C.c_uint64(0), C.c_uint64(0)
    nh = np.zeros(trace_nodes, dtype=np.uint64) if trace_nodes else None
    ub = np.zeros(trace_nodes, dtype=np.float64) if trace_nodes else None
    delta = np.zeros((N, K), dtype=np.float64) if want_state else None
    psi = np.zeros((N, K), dtype=np.uint64) if want_state else None
    rc = lib().cvo_cp_solve(K, M, _p(logA, C.c_double), _p(logB, C.c_double), _p(logPi, C.c_double), N,
                            _p(obs, C.c_uint32), _p(start, C.c_uint8), _p(comp, C.c_int32), int(ncomp),
                            int(max_nodes), _p(sol, C.c_uint64), C.byref(obj), C.byref(explored),
                            C.byref(steps), _p(nh, C.c_uint64), int(trace_nodes), _p(ub, C.c_double),
                            _p(delta, C.c_double), _p(psi, C.c_uint64))
    if rc:
        raise OracleError(rc)
    return dict(sol=sol[:N], obj=obj.value, explored=explored.value, steps=steps.value,
                node_hash=nh, ub=ub, delta=delta, psi=psi)


def mle(a, b, pi, obs_flat, tags_flat, seq_off):
    """HMM::maximum_likelihood_estimation + log (hmm.rs:30-62,192-205) on top of the model (a, b, pi);
    returns new (logA, logB, logPi)."""
    a, b, pi = _f64(a).copy(), _f64(b).copy(), _f64(pi).copy()
    K, M = b.shape
    obs = np.ascontiguousarray(obs_flat, dtype=np.uint32)
    tags = np.ascontiguousarray(tags_flat, dtype=np.int32)
    off = np.ascontiguousarray(seq_off, dtype=np.int64)
    rc = lib().cvo_mle(K, M, _p(a, C.c_double), _p(b, C.c_double), _p(pi, C.c_double), _p(obs, C.c_uint32),
                       _p(tags, C.c_int32), _p(off, C.c_int64), len(off) - 1)
    if rc:
        raise OracleError(rc)
    return a, b, pi


def cfn_tables(logA, logB, logPi, obs, is_seq_start, comp, k):
    """write_cfn's cost tables / unary costs / lower bound (cfn.rs:11-167)."""
    logA, logB, logPi = _f64(logA), _f64(logB), _f64(logPi)
    K, M = logB.shape
    obs = np.ascontiguousarray(obs, dtype=np.uint32)
    start = np.ascontiguousarray(is_seq_start, dtype=np.uint8)
    comp = np.ascontiguousarray(comp, dtype=np.int32)
    tables = np.zeros((k, k, K, K), dtype=np.float64)
    unary = np.zeros((k, K), dtype=np.float64)
    lb = C.c_double(0.0)
    nb = C.c_int64(0)
    rc = lib().cvo_cfn_tables(K, M, _p(logA, C.c_double), _p(logB, C.c_double), _p(logPi, C.c_double), len(obs),
                              _p(obs, C.c_uint32), _p(start, C.c_uint8), _p(comp, C.c_int32), int(k),
                              _p(tables, C.c_double), _p(unary, C.c_double), C.byref(lb), C.byref(nb))
    if rc:
        raise OracleError(rc)
    return dict(tables=tables, unary=unary, lower_bound=lb.value, nboundaries=nb.value)
