/*
 * cv_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see cv_oracle.h).
 * PARITY UNPINNED (no reference tests / fixtures exist; reference is Rust and
 * cannot be built here) -- pinned by brute-force KATs and an independent
 * Python restatement only.
 *
 * Compile WITHOUT fast-math and without FMA contraction:
 *   gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC cv_oracle.c -o libcv_oracle.so
 */
#include "cv_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NEG_INF (-INFINITY)

/* ndarray-stats 0.5 QuantileExt::argmax on a 1-D f64 array: the running maximum
 * starts at element 0 and an element replaces it only when partial_cmp says
 * Greater; any incomparable pair (NaN) is Err(UndefinedOrder) which the
 * reference unwrap()s (viterbi.rs:16,24; cp.rs:39,53,74,86).  Empty => Err. */
static inline int64_t argmax_first(const double *v, int64_t n, int *nan_flag)
{
    int64_t best = 0;
    double cur = v[0];
    for (int64_t j = 0; j < n; j++) {
        double x = v[j];
        if (x != x || cur != cur) { *nan_flag = 1; return 0; }
        if (x > cur) { cur = x; best = j; }
    }
    return best;
}

/* Test hook: the argmax above on a caller-supplied vector, with the reference's error cases
 * (ndarray-stats MinMaxError::{EmptyInput, UndefinedOrder}). */
int cvo_argmax(const double *v, int64_t n, int64_t *idx_out)
{
    if (n <= 0) return CVO_ERR_EMPTY;
    int nan_flag = 0;
    int64_t i = argmax_first(v, n, &nan_flag);
    if (nan_flag) return CVO_ERR_NAN;
    if (idx_out) *idx_out = i;
    return CVO_OK;
}

/* ------------------------------------------------------------------------ */
/* R1: viterbi::decode (viterbi.rs:5-32)                                      */
/* ------------------------------------------------------------------------ */

/* logAT is logA transposed (logAT[to*K + from]) so that hmm.transitions_to(to)
 * (hmm.rs:224-226, a column view) is contiguous; values and add order are
 * unchanged. */
static int decode_core(int K, int64_t M, const double *logAT, const double *logB,
                       const uint32_t *obs, int64_t T, uint32_t *path_out,
                       double *score_out, double *delta_full, uint32_t *psi_full,
                       double *row_scratch /*[3*K]*/, uint32_t *psi_scratch /*[T*K] or NULL if psi_full*/)
{
    if (T <= 0) return CVO_ERR_EMPTY;               /* sequence.len()-1 underflows: panic */
    for (int64_t t = 0; t < T; t++)
        if ((int64_t)obs[t] >= M) return CVO_ERR_ARG; /* ndarray index panic */

    double *prev = row_scratch, *cur = row_scratch + K, *probs = row_scratch + 2 * (int64_t)K;
    uint32_t *psi = psi_full ? psi_full : psi_scratch;
    int nan_flag = 0;

    for (int i = 0; i < K; i++) prev[i] = 0.0;      /* viterbi.rs:6: row 0 stays 0.0 (pi, o_0 unused) */
    if (psi) for (int i = 0; i < K; i++) psi[i] = 0; /* viterbi.rs:7 */
    if (delta_full) memcpy(delta_full, prev, sizeof(double) * K);

    for (int64_t t = 1; t < T; t++) {                /* viterbi.rs:9 */
        uint32_t o = obs[t];
        for (int to = 0; to < K; to++) {             /* viterbi.rs:10 */
            double emit = logB[(int64_t)to * M + o]; /* viterbi.rs:11 */
            if (emit > NEG_INF) {                    /* viterbi.rs:12 */
                const double *tr = logAT + (int64_t)to * K;
                for (int j = 0; j < K; j++) probs[j] = prev[j] + tr[j];  /* viterbi.rs:15 */
                int64_t sf = argmax_first(probs, K, &nan_flag);           /* viterbi.rs:16 */
                cur[to] = probs[sf] + emit;                               /* viterbi.rs:17 */
                psi[t * K + to] = (uint32_t)sf;                           /* viterbi.rs:18 */
            } else {
                cur[to] = NEG_INF;                                        /* viterbi.rs:20 */
                psi[t * K + to] = 0;                                      /* bt keeps its initial 0 */
            }
        }
        if (nan_flag) return CVO_ERR_NAN;
        if (delta_full) memcpy(delta_full + t * K, cur, sizeof(double) * K);
        double *tmp = prev; prev = cur; cur = tmp;
    }
    int64_t end = argmax_first(prev, K, &nan_flag);  /* viterbi.rs:24 */
    if (nan_flag) return CVO_ERR_NAN;
    if (score_out) *score_out = prev[end];
    if (path_out) {
        path_out[T - 1] = (uint32_t)end;             /* viterbi.rs:26 */
        for (int64_t t = T - 2; t >= 0; t--) {       /* viterbi.rs:27-30 */
            end = psi[(t + 1) * K + end];
            path_out[t] = (uint32_t)end;
        }
    }
    return CVO_OK;
}

static double *transpose_a(int K, const double *logA)
{
    double *at = (double *)malloc(sizeof(double) * (size_t)K * K);
    if (!at) return NULL;
    for (int f = 0; f < K; f++)
        for (int t = 0; t < K; t++) at[(size_t)t * K + f] = logA[(size_t)f * K + t];
    return at;
}

int cvo_decode(int K, int64_t M, const double *logA, const double *logB,
               const uint32_t *obs, int64_t T, uint32_t *path_out, double *score_out)
{
    if (K <= 0 || M <= 0) return CVO_ERR_ARG;
    if (T <= 0) return CVO_ERR_EMPTY;
    double *at = transpose_a(K, logA);
    double *rows = (double *)malloc(sizeof(double) * 3 * (size_t)K);
    uint32_t *psi = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)T * K);
    int rc = decode_core(K, M, at, logB, obs, T, path_out, score_out, NULL, NULL, rows, psi);
    free(at); free(rows); free(psi);
    return rc;
}

int cvo_decode_trace(int K, int64_t M, const double *logA, const double *logB,
                     const uint32_t *obs, int64_t T, double *delta_out, uint32_t *psi_out)
{
    if (K <= 0 || M <= 0) return CVO_ERR_ARG;
    if (T <= 0) return CVO_ERR_EMPTY;
    double *at = transpose_a(K, logA);
    double *rows = (double *)malloc(sizeof(double) * 3 * (size_t)K);
    int rc = decode_core(K, M, at, logB, obs, T, NULL, NULL, delta_out, psi_out, rows, NULL);
    free(at); free(rows);
    return rc;
}

int cvo_decode_batch(int K, int64_t M, const double *logA, const double *logB,
                     const uint32_t *obs, const int64_t *seq_off, int64_t B,
                     uint32_t *path_out, double *score_out, int nthreads)
{
    if (K <= 0 || M <= 0 || B < 0) return CVO_ERR_ARG;
    double *at = transpose_a(K, logA);
    int64_t maxT = 0;
    for (int64_t b = 0; b < B; b++) {
        int64_t T = seq_off[b + 1] - seq_off[b];
        if (T > maxT) maxT = T;
    }
    int rc_all = CVO_OK;
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        double *rows = (double *)malloc(sizeof(double) * 3 * (size_t)K);
        uint32_t *psi = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(maxT > 0 ? maxT : 1) * K);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int64_t b = 0; b < B; b++) {
            int64_t off = seq_off[b], T = seq_off[b + 1] - off;
            int rc = decode_core(K, M, at, logB, obs + off, T, path_out + off,
                                 score_out ? score_out + b : NULL, NULL, NULL, rows, psi);
            if (rc != CVO_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
                { if (rc_all == CVO_OK) rc_all = rc; }
            }
        }
        free(rows); free(psi);
    }
    free(at);
    return rc_all;
}

/* ------------------------------------------------------------------------ */
/* R2: CPSolver (cp.rs:8-152)                                                 */
/* ------------------------------------------------------------------------ */

typedef struct {
    int K; int64_t M, N;
    const double *logA, *logAT, *logB, *logPi;
    const uint32_t *obs; const uint8_t *start; const int32_t *comp;
    int32_t ncomp;
    int64_t **cons; int64_t *cons_len;       /* cp.rs:21-27 */
    int64_t *choice;                          /* cstr_choices, -1 = None (cp.rs:28) */
    double best_obj; uint64_t *best_sol;      /* cp.rs:29 */
    uint64_t explored, max_nodes, steps;
    double *delta; uint64_t *psi;             /* array / bt (cp.rs:134-135) */
    double *probs;
    int nan_flag, assert_flag;
    uint64_t *node_hash; double *ub_log; uint64_t node_cap;
} cp_t;

/* MetaElements::arc_p (viterbi_solver/utils.rs:24-30) with HMM::init_prob
 * (hmm.rs:211-213) and HMM::transition_prob (hmm.rs:220-222). */
static inline double arc_p(const cp_t *s, int64_t t, int64_t from, int64_t to)
{
    double b = s->logB[to * s->M + s->obs[t]];
    if (s->start[t]) return s->logPi[to] + b;
    return s->logA[from * s->K + to] + b;
}

/* probs = &array.row(t-1) + &sequence[t].transitions(hmm, state)
 * (viterbi_solver/utils.rs:32-38): constant pi[state] when el.t == 0, column
 * `state` of a otherwise. */
static inline void cand(const cp_t *s, int64_t t, int64_t state, double *probs)
{
    const double *prev = s->delta + (t - 1) * s->K;
    if (s->start[t]) {
        double p = s->logPi[state];
        for (int j = 0; j < s->K; j++) probs[j] = prev[j] + p;
    } else {
        const double *tr = s->logAT + state * s->K;
        for (int j = 0; j < s->K; j++) probs[j] = prev[j] + tr[j];
    }
}

static inline int fixed_at(const cp_t *s, int64_t t)
{   /* sequence[t].is_constrained() && cstr_choices[comp].is_some() (cp.rs:43,48) */
    return s->comp[t] >= 0 && s->choice[s->comp[t]] >= 0;
}

/* one forward step over all target states: cp.rs:49-58 == cp.rs:70-78 */
static void sweep_row(cp_t *s, int64_t t)
{
    int K = s->K;
    for (int st = 0; st < K; st++) {
        cand(s, t, st, s->probs);
        int64_t sf = argmax_first(s->probs, K, &s->nan_flag);
        double arc = arc_p(s, t, sf, st);
        s->delta[t * K + st] = s->delta[(t - 1) * K + sf] + arc;   /* delta + (a + b) */
        s->psi[t * K + st] = (uint64_t)sf;
    }
    s->steps++;
}

/* cp.rs:32-61 */
static void viterbi_from(cp_t *s, int64_t from, int64_t node)
{
    int K = s->K;
    for (int i = 0; i < K; i++) s->delta[from * K + i] = NEG_INF;  /* cp.rs:33 */
    s->delta[from * K + node] = 0.0;                               /* cp.rs:34 */
    if (from != 0) {                                               /* cp.rs:35-41 */
        cand(s, from, node, s->probs);
        s->psi[from * K + node] = (uint64_t)argmax_first(s->probs, K, &s->nan_flag);
    }
    if (from + 1 < s->N && fixed_at(s, from + 1))                  /* cp.rs:43-45 */
        s->psi[(from + 1) * K + s->choice[s->comp[from + 1]]] = (uint64_t)node;
    int64_t t = from + 1;                                          /* cp.rs:47-60 */
    while (t < s->N && !fixed_at(s, t)) { sweep_row(s, t); t++; }
}

/* cp.rs:63-83 */
static void init_viterbi(cp_t *s)
{
    int64_t t = 0;
    while (t < s->N && !(s->comp[t] >= 0)) {
        if (t == 0) {   /* hmm.init_probs(value) = &pi + &b[.][o]  (hmm.rs:215-218) */
            for (int i = 0; i < s->K; i++)
                s->delta[i] = s->logPi[i] + s->logB[(int64_t)i * s->M + s->obs[0]];
        } else {
            sweep_row(s, t);
        }
        t++;
    }
}

/* cp.rs:85-93 */
static void backtrack(cp_t *s, double obj)
{
    int K = s->K;
    int64_t cur = argmax_first(s->delta + (s->N - 1) * K, K, &s->nan_flag);
    if (!(obj > s->best_obj)) { s->assert_flag = 1; return; }
    s->best_obj = obj;
    for (int64_t t = s->N - 1; t >= 0; t--) {
        s->best_sol[t] = (uint64_t)cur;
        cur = (int64_t)s->psi[t * K + cur];
    }
}

static uint64_t state_hash(const cp_t *s)
{
    uint64_t h = 1469598103934665603ULL;
    const unsigned char *p = (const unsigned char *)s->delta;
    size_t n = sizeof(double) * (size_t)s->N * s->K;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
    for (int64_t i = 0; i < s->N * s->K; i++) {
        uint64_t v = s->psi[i];
        for (int k = 0; k < 4; k++) { h ^= (v >> (8 * k)) & 0xff; h *= 1099511628211ULL; }
    }
    return h;
}

/* cp.rs:95-126 */
static void solve_r(cp_t *s, int32_t comp)
{
    for (int state = 0; state < s->K; state++) {
        if (s->max_nodes && s->explored >= s->max_nodes) break;    /* builder-added budget */
        s->explored++;                                             /* cp.rs:97 */
        s->choice[comp] = state;                                   /* cp.rs:98 */
        for (int64_t idx = 0; idx < s->cons_len[comp]; idx++)      /* cp.rs:99-102 */
            viterbi_from(s, s->cons[comp][idx], state);
        double ub = 0.0;                                           /* cp.rs:103 */
        for (int32_t cid = 0; cid <= comp; cid++) {                /* cp.rs:104-116 */
            int64_t st = s->choice[cid];
            for (int64_t idx = 0; idx < s->cons_len[cid]; idx++) {
                int64_t t = s->cons[cid][idx];
                if (t == 0) {
                    ub += arc_p(s, 0, 0, st);
                } else {
                    int64_t sf = (int64_t)s->psi[t * s->K + st];
                    double arc = arc_p(s, t, sf, st);
                    ub += s->delta[(t - 1) * s->K + sf] + arc;
                }
            }
        }
        if (s->node_cap && s->explored - 1 < s->node_cap) {
            if (s->node_hash) s->node_hash[s->explored - 1] = state_hash(s);
            if (s->ub_log) s->ub_log[s->explored - 1] = ub;
        }
        if (s->nan_flag || s->assert_flag) return;
        if (ub != ub) { s->nan_flag = 1; return; }   /* NaN compares false in Rust too, but flag it */
        if (ub > s->best_obj) {                                    /* cp.rs:117 */
            if (comp + 1 < s->ncomp) solve_r(s, comp + 1);         /* cp.rs:118-119 */
            else backtrack(s, ub);                                 /* cp.rs:121 */
        }
        if (s->nan_flag || s->assert_flag) return;
    }
    s->choice[comp] = -1;                                          /* cp.rs:125 */
}

int cvo_cp_solve(int K, int64_t M, const double *logA, const double *logB,
                 const double *logPi, int64_t N, const uint32_t *obs,
                 const uint8_t *is_seq_start, const int32_t *comp, int32_t ncomp,
                 uint64_t max_nodes, uint64_t *sol_out, double *obj_out,
                 uint64_t *explored_out, uint64_t *steps_out,
                 uint64_t *node_hash_out, uint64_t node_hash_cap, double *ub_out,
                 double *delta_out, uint64_t *psi_out)
{
    if (K <= 0 || M <= 0 || ncomp < 0) return CVO_ERR_ARG;
    if (N <= 0) return CVO_ERR_EMPTY;                /* array.row(len-1) panics */
    for (int64_t t = 0; t < N; t++) {
        if ((int64_t)obs[t] >= M) return CVO_ERR_ARG;
        if (comp[t] >= ncomp) return CVO_ERR_ARG;    /* constraints[ucomp] index panic (cp.rs:25) */
    }
    cp_t s; memset(&s, 0, sizeof(s));
    s.K = K; s.M = M; s.N = N; s.logA = logA; s.logB = logB; s.logPi = logPi;
    s.obs = obs; s.start = is_seq_start; s.comp = comp; s.ncomp = ncomp;
    s.max_nodes = max_nodes;
    s.node_hash = node_hash_out; s.ub_log = ub_out; s.node_cap = node_hash_cap;
    double *at = transpose_a(K, logA); s.logAT = at;
    s.cons = (int64_t **)calloc((size_t)(ncomp > 0 ? ncomp : 1), sizeof(int64_t *));
    s.cons_len = (int64_t *)calloc((size_t)(ncomp > 0 ? ncomp : 1), sizeof(int64_t));
    s.choice = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ncomp > 0 ? ncomp : 1));
    for (int32_t c = 0; c < ncomp; c++) s.choice[c] = -1;
    for (int64_t t = 0; t < N; t++) if (comp[t] >= 0) s.cons_len[comp[t]]++;
    for (int32_t c = 0; c < ncomp; c++) {
        s.cons[c] = (int64_t *)malloc(sizeof(int64_t) * (size_t)(s.cons_len[c] > 0 ? s.cons_len[c] : 1));
        s.cons_len[c] = 0;
    }
    for (int64_t t = 0; t < N; t++) if (comp[t] >= 0) s.cons[comp[t]][s.cons_len[comp[t]]++] = t;  /* ascending */
    s.best_obj = NEG_INF;
    s.best_sol = sol_out;
    for (int64_t t = 0; t < N; t++) sol_out[t] = 0;  /* Array1::from_elem(len, 0) (cp.rs:29) */
    s.delta = (double *)calloc((size_t)N * K, sizeof(double));      /* 0.0 (cp.rs:134) */
    s.psi = (uint64_t *)calloc((size_t)N * K, sizeof(uint64_t));    /* 0   (cp.rs:135) */
    s.probs = (double *)malloc(sizeof(double) * (size_t)K);

    init_viterbi(&s);                                               /* cp.rs:136 */
    if (!s.nan_flag) {
        if (ncomp > 0) {
            solve_r(&s, 0);                                         /* cp.rs:137-138 */
        } else {                                                    /* cp.rs:139-142 */
            const double *last = s.delta + (N - 1) * K;
            double obj = last[0];
            for (int i = 0; i < K; i++) {                           /* QuantileExt::max */
                if (last[i] != last[i]) s.nan_flag = 1;
                if (last[i] > obj) obj = last[i];
            }
            if (!s.nan_flag) backtrack(&s, obj);
        }
    }
    if (obj_out) *obj_out = s.best_obj;
    if (explored_out) *explored_out = s.explored;
    if (steps_out) *steps_out = s.steps;
    if (delta_out) memcpy(delta_out, s.delta, sizeof(double) * (size_t)N * K);
    if (psi_out) memcpy(psi_out, s.psi, sizeof(uint64_t) * (size_t)N * K);
    int rc = s.nan_flag ? CVO_ERR_NAN : (s.assert_flag ? CVO_ERR_ASSERT : CVO_OK);
    for (int32_t c = 0; c < ncomp; c++) free(s.cons[c]);
    free(s.cons); free(s.cons_len); free(s.choice); free(s.delta); free(s.psi); free(s.probs); free(at);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * Supervised maximum-likelihood estimation, HMM::maximum_likelihood_estimation (hmm.rs:30-62) followed by
 * HMM::log (hmm.rs:192-205).  Literal: the counts are added one `+= 1.0` at a time ON TOP of whatever the
 * model holds (the reference calls this on HMM::new's random model, hmm.rs:22-28), in the reference's order.
 * tags[t] < 0 stands for None (reference: unwrap() panic).  a [K*K], b [K*M], pi [K] are probabilities on
 * entry and log10-as-the-reference-computes-it on exit: f64::log(10.0) = ln(x) / ln(10).
 * ------------------------------------------------------------------------------------------------ */
int cvo_mle(int K, int64_t M, double *a, double *b, double *pi, const uint32_t *obs, const int32_t *tags,
            const int64_t *seq_off, int64_t B)
{
    if (K <= 0 || M <= 0 || !a || !b || !pi || B < 0) return CVO_ERR_ARG;
    double *seen = (double *)calloc((size_t)K, sizeof(double)), *end = (double *)calloc((size_t)K, sizeof(double));
    if (!seen || !end) { free(seen); free(end); return CVO_ERR_ARG; }
    int rc = CVO_OK;
    for (int64_t i = 0; i < B && !rc; i++) {                               /* hmm.rs:35-48 */
        const int64_t o = seq_off[i], T = seq_off[i + 1] - seq_off[i];
        if (T <= 0) { rc = CVO_ERR_EMPTY; break; }                         /* tag[0] / len()-1 panics */
        for (int64_t t = 0; t < T; t++)
            if (tags[o + t] < 0 || tags[o + t] >= K || (int64_t)obs[o + t] >= M) rc = CVO_ERR_ARG;
        if (rc) break;
        pi[tags[o]] += 1.0;                                                /* :39 */
        for (int64_t t = 0; t < T - 1; t++) {                              /* :40-44 */
            b[(int64_t)tags[o + t] * M + obs[o + t]] += 1.0;
            a[(int64_t)tags[o + t] * K + tags[o + t + 1]] += 1.0;
            seen[tags[o + t]] += 1.0;
        }
        b[(int64_t)tags[o + T - 1] * M + obs[o + T - 1]] += 1.0;           /* :45-47 */
        seen[tags[o + T - 1]] += 1.0;
        end[tags[o + T - 1]] += 1.0;
    }
    if (!rc) {
        for (int s = 0; s < K; s++) {                                      /* hmm.rs:50-59 */
            if (seen[s] != end[s]) { const double d = seen[s] - end[s]; for (int j = 0; j < K; j++) a[(int64_t)s * K + j] /= d; }
            else for (int j = 0; j < K; j++) a[(int64_t)s * K + j] = 0.0;
            pi[s] /= (double)B;
            for (int64_t m = 0; m < M; m++) b[(int64_t)s * M + m] /= seen[s];
        }
        const double ln10 = log(10.0);                                     /* hmm.rs:192-205: x.log(10.0) */
        for (int64_t e = 0; e < (int64_t)K * K; e++) a[e] = a[e] == 0.0 ? -INFINITY : log(a[e]) / ln10;
        for (int64_t e = 0; e < (int64_t)K * M; e++) b[e] = b[e] == 0.0 ? -INFINITY : log(b[e]) / ln10;
        for (int e = 0; e < K; e++) pi[e] = pi[e] == 0.0 ? -INFINITY : log(pi[e]) / ln10;
    }
    free(seen); free(end);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * CFN cost-table compilation, write_cfn's numeric part (reference src/viterbi_solver/cfn.rs:11-167): the
 * boundaries between constraint components, K*K clamped longest_path runs per consecutive boundary pair,
 * their accumulation into the k*k cost tables (with the reference's `== 0.0 => assign, else +=` rule), the
 * unary start/end costs, the lower bound and the -inf -> lower_bound patch of the unary costs.  The file
 * writer (cfn.rs:169-205) is not restated.  Step arithmetic: max_j fl(row[j] + tr_j) (first maximum), + emit.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int K; int64_t M, N;
    const double *A, *B, *Pi; const uint32_t *obs; const uint8_t *start; const int32_t *comp;
} cfn_t;

static double cfn_cell(const cfn_t *s, const double *prev, int64_t t, int n)
{
    /* transitions(hmm, n) (utils.rs:32-38) added to the previous row, max (ndarray-stats), + emit_prob */
    double best = 0.0;
    for (int j = 0; j < s->K; j++) {
        const double tr = s->start[t] ? s->Pi[n] : s->A[(int64_t)j * s->K + n];
        const double v = prev[j] + tr;
        if (j == 0 || v > best) best = v;
    }
    return best + s->B[(int64_t)n * s->M + s->obs[t]];
}

/* cfn.rs:11-35 */
static double cfn_longest_path(const cfn_t *s, double *array, int64_t t_from, int n_from, int64_t t_to, int n_to)
{
    const int K = s->K;
    for (int j = 0; j < K; j++) array[j] = -INFINITY;
    array[n_from] = 0.0;
    int64_t t = 1;
    while (t_from + t <= t_to) {
        double *row = array + t * K; const double *prev = row - K;
        if (s->comp[t_from + t] >= 0) {
            for (int j = 0; j < K; j++) row[j] = -INFINITY;
            const int n = (t_from + t < t_to) ? n_from : n_to;
            row[n] = cfn_cell(s, prev, t_from + t, n);
        } else {
            for (int n = 0; n < K; n++) row[n] = cfn_cell(s, prev, t_from + t, n);
        }
        t++;
    }
    const double *last = array + (t - 1) * K;
    double m = last[0];
    for (int j = 1; j < K; j++) if (last[j] > m) m = last[j];
    return m;
}

int cvo_cfn_tables(int K, int64_t M, const double *logA, const double *logB, const double *logPi, int64_t N,
                   const uint32_t *obs, const uint8_t *is_seq_start, const int32_t *comp, int32_t k,
                   double *cost_tables /* [k][k][K][K] */, double *unary /* [k][K] */, double *lower_bound_out,
                   int64_t *nboundaries_out)
{
    if (K <= 0 || N <= 0 || k <= 0) return CVO_ERR_ARG;
    cfn_t s = {K, M, N, logA, logB, logPi, obs, is_seq_start, comp};
    for (int64_t t = 0; t < N; t++) if ((int64_t)obs[t] >= M || comp[t] >= k) return CVO_ERR_ARG;
    /* cfn.rs:82-112: boundaries (t, cid) where the component changes; longest segment */
    int64_t *bt = (int64_t *)malloc(sizeof(int64_t) * (size_t)N); int32_t *bc = (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
    int64_t nb = 0, longest = 0; int32_t last_cid = -1;
    for (int64_t t = 0; t < N; t++) {
        if (comp[t] >= 0) {
            const int32_t cid = comp[t];
            if (last_cid < 0) { longest = t + 1; bt[nb] = t; bc[nb] = cid; nb++; }
            else if (last_cid != cid) {
                const int64_t seg = t - bt[nb - 1] + 1;
                if (seg > longest) longest = seg;
                bt[nb] = t; bc[nb] = cid; nb++;
            }
            last_cid = cid;
        }
    }
    if (nb == 0) { free(bt); free(bc); return CVO_ERR_EMPTY; }            /* constraint_boundaries.last().unwrap() panics */
    { const int64_t seg = N - bt[nb - 1] + 1; if (seg > longest) longest = seg; }
    double *array = (double *)calloc((size_t)(longest + 1) * K, sizeof(double));
    for (int64_t e = 0; e < (int64_t)k * k * K * K; e++) cost_tables[e] = 0.0;
    #define TAB(c1, c2, n1, n2) cost_tables[(((int64_t)(c1) * k + (c2)) * K + (n1)) * K + (n2)]
    for (int64_t i = 0; i + 1 < nb; i++) {                                  /* cfn.rs:118-137 */
        for (int n1 = 0; n1 < K; n1++)
            for (int n2 = 0; n2 < K; n2++) {
                const double cost = cfn_longest_path(&s, array, bt[i], n1, bt[i + 1], n2);
                if (cost != -INFINITY) {
                    if (TAB(bc[i], bc[i + 1], n1, n2) == 0.0) TAB(bc[i], bc[i + 1], n1, n2) = cost; else TAB(bc[i], bc[i + 1], n1, n2) += cost;
                    if (TAB(bc[i + 1], bc[i], n2, n1) == 0.0) TAB(bc[i + 1], bc[i], n2, n1) = cost; else TAB(bc[i + 1], bc[i], n2, n1) += cost;
                }
            }
    }
    for (int64_t e = 0; e < (int64_t)k * K; e++) unary[e] = 0.0;          /* cfn.rs:139-145 */
    {   /* get_unary_start_cost, cfn.rs:37-55 */
        const int64_t tl = bt[0];
        for (int n = 0; n < K; n++) array[n] = logPi[n] + logB[(int64_t)n * M + obs[0]];
        for (int64_t t = 1; t <= tl; t++)
            for (int n = 0; n < K; n++) array[t * K + n] = cfn_cell(&s, array + (t - 1) * K, t, n);
        for (int n = 0; n < K; n++) unary[(int64_t)bc[0] * K + n] += array[tl * K + n];
    }
    {   /* get_unary_end_cost, cfn.rs:57-80 */
        const int64_t ts = bt[nb - 1];
        for (int n = 0; n < K; n++) {
            double c = 0.0;
            if (ts != N - 1) {
                for (int j = 0; j < K; j++) array[j] = -INFINITY;
                array[n] = 0.0;
                for (int64_t t = ts + 1; t < N; t++) {
                    double *row = array + (t - ts) * K; const double *prev = row - K;
                    if (comp[t] >= 0) { for (int j = 0; j < K; j++) row[j] = -INFINITY; row[n] = cfn_cell(&s, prev, t, n); }
                    else for (int nn = 0; nn < K; nn++) row[nn] = cfn_cell(&s, prev, t, nn);
                }
                const double *last = array + (N - 1 - ts) * K;
                c = last[0];
                for (int j = 1; j < K; j++) if (last[j] > c) c = last[j];
            }
            unary[(int64_t)bc[nb - 1] * K + n] += c;
        }
    }
    double lb = -1.0;                                                        /* cfn.rs:149-157 */
    for (int k1 = 0; k1 < k; k1++)
        for (int k2 = k1 + 1; k2 < k; k2++) {
            const double *tb = &TAB(k1, k2, 0, 0);
            double m = tb[0];
            for (int e = 1; e < K * K; e++) if (tb[e] < m) m = tb[e];
            lb += m;
        }
    for (int64_t e = 0; e < (int64_t)k * K; e++) if (unary[e] == -INFINITY) unary[e] = lb;   /* cfn.rs:160-166 */
    #undef TAB
    if (lower_bound_out) *lower_bound_out = lb;
    if (nboundaries_out) *nboundaries_out = nb;
    free(array); free(bt); free(bc);
    return CVO_OK;
}
