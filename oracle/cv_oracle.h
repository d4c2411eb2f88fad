/*
 * cv_oracle.h -- CPU oracle for the consistent-viterbi hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the
 * CPU baseline.  The product path (consistent_viterbi_b200/, include/) never
 * links or calls it.
 *
 * PARITY UNPINNED: the reference (AlexandreDubray/consistent-viterbi) ships no
 * tests, golden vectors or fixtures, and it is Rust with un-vendored crates
 * (no cargo/rustc in this image), so it can neither be run nor compiled here.
 * This file is a literal restatement of the reference's loops, written from
 * the cited lines, pinned only by (1) brute-force path enumeration KATs,
 * (2) an independent pure-Python restatement (oracle/py_restatement.py) and
 * (3) the per-node state traces both produce.  Third-party arithmetic on the
 * path: ndarray 0.15 element-wise f64 add (IEEE binary64 round-to-nearest) and
 * ndarray-stats 0.5 QuantileExt::argmax / max (first element that compares
 * strictly Greater than the running maximum wins; NaN => Err => panic).
 *
 * Layouts (all log10 probabilities as the reference stores them, -inf = zero):
 *   logA  [K*K]  row-major, logA[from*K + to]            (hmm.rs:12-13, a[[from,to]])
 *   logB  [K*M]  state-major, logB[state*M + obs]        (hmm.rs:14-15; obs = the
 *                D-dimensional observation flattened row-major over bdims)
 *   logPi [K]                                            (hmm.rs:16-17)
 */
#ifndef CV_ORACLE_H
#define CV_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes shared with the product ABI (include/cv_b200.h) */
#define CVO_OK            0
#define CVO_ERR_EMPTY     1   /* reference would panic: empty sequence / N == 0        */
#define CVO_ERR_NAN       2   /* reference would panic: argmax() on NaN                */
#define CVO_ERR_ARG       3   /* bad argument (obs out of range, comp id >= ncomp ...) */
#define CVO_ERR_ASSERT    4   /* reference assert!(obj > best_obj) would fire (cp.rs:87) */

/* ndarray-stats 0.5 QuantileExt::argmax on a 1-D f64 vector (test hook): first element that compares Greater than
 * the running maximum; CVO_ERR_EMPTY for n == 0 (MinMaxError::EmptyInput), CVO_ERR_NAN for NaN (UndefinedOrder). */
int cvo_argmax(const double *v, int64_t n, int64_t *idx_out);

/* R1 -- viterbi::decode, src/viterbi_solver/viterbi.rs:5-32.
 * path_out[T]; *score_out = delta[T-1][end_state]. */
int cvo_decode(int K, int64_t M, const double *logA, const double *logB,
               const uint32_t *obs, int64_t T, uint32_t *path_out, double *score_out);

/* B independent cvo_decode calls; seq_off[B+1] are offsets into obs/path_out.
 * nthreads <= 1: serial (the reference is single-threaded); > 1: OpenMP over
 * sequences (used only for the all-cores CPU baseline). */
int cvo_decode_batch(int K, int64_t M, const double *logA, const double *logB,
                     const uint32_t *obs, const int64_t *seq_off, int64_t B,
                     uint32_t *path_out, double *score_out, int nthreads);

/* Optional dump of the full delta / psi rows of one cvo_decode call (tests). */
int cvo_decode_trace(int K, int64_t M, const double *logA, const double *logB,
                     const uint32_t *obs, int64_t T, double *delta_out /*[T*K]*/,
                     uint32_t *psi_out /*[T*K]*/);

/* R2 -- CPSolver::new + solve + get_*, src/viterbi_solver/cp.rs:20-152 with
 * MetaElements::{arc_p,transitions,is_constrained} (viterbi_solver/utils.rs:24-46)
 * and the HMM accessors (hmm.rs:207-234).
 *   obs[N]          flattened observation of every super-sequence element
 *   is_seq_start[N] 1 iff el.t == 0
 *   comp[N]         constraint component of ACTIVE elements, -1 otherwise
 *   ncomp           SuperSequence::number_constraints()
 *   max_nodes       0 = unlimited (reference behaviour); otherwise stop opening
 *                   new nodes once explored_nodes == max_nodes (deterministic,
 *                   applied identically by the GPU path)
 * Outputs: sol_out[N], *obj_out, *explored_out, *steps_out (forward sweep steps
 * executed = cells / K^2), node_hash_out[max(explored,1)] (optional, may be NULL:
 * FNV-1a hash of the full delta/psi state after every node, cap entries). */
int cvo_cp_solve(int K, int64_t M, const double *logA, const double *logB,
                 const double *logPi, int64_t N, const uint32_t *obs,
                 const uint8_t *is_seq_start, const int32_t *comp, int32_t ncomp,
                 uint64_t max_nodes, uint64_t *sol_out, double *obj_out,
                 uint64_t *explored_out, uint64_t *steps_out,
                 uint64_t *node_hash_out, uint64_t node_hash_cap,
                 double *ub_out /* optional [node_hash_cap] */,
                 double *delta_out /* optional [N*K] final state */,
                 uint64_t *psi_out /* optional [N*K] final state */);

/* HMM::maximum_likelihood_estimation + HMM::log (hmm.rs:30-62,192-205), literal: counts are added with one
 * `+= 1.0` per event on top of the model passed in (the reference starts from HMM::new's random model);
 * a/b/pi are probabilities on entry, ln(x)/ln(10) (-inf for 0) on exit.  tags < 0 = None (unwrap panic). */
int cvo_mle(int K, int64_t M, double *a, double *b, double *pi, const uint32_t *obs, const int32_t *tags,
            const int64_t *seq_off, int64_t B);

/* write_cfn's numeric part (cfn.rs:11-167): cost tables [k][k][K][K], unary costs [k][K] (with -inf replaced by
 * the lower bound), the lower bound, the number of component boundaries.  comp[t] = component of ACTIVE elements,
 * -1 otherwise; k = number_constraints().  No boundary at all => CVO_ERR_EMPTY (the reference unwraps None). */
int cvo_cfn_tables(int K, int64_t M, const double *logA, const double *logB, const double *logPi, int64_t N,
                   const uint32_t *obs, const uint8_t *is_seq_start, const int32_t *comp, int32_t k,
                   double *cost_tables, double *unary, double *lower_bound_out, int64_t *nboundaries_out);

#ifdef __cplusplus
}
#endif
#endif
