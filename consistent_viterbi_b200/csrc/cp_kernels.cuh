// cp_kernels.cuh -- device kernels of the constrained decode (mode R2).
//
// Replaces the inner loops of CPSolver (reference src/viterbi_solver/cp.rs:32-126):
// viterbi_from / init_viterbi sweeps, the backpointer fix-ups of viterbi_from, the
// upper-bound terms of solve_r and backtrack.  The branch-and-bound control flow
// stays on the host (cp_host.inl) exactly as the reference's recursion.
//
// Persistent device state, as in the reference (cp.rs:134-135): delta f64 [N][K]
// (init 0.0) and psi [N][K] (init 0), never restored between nodes.
//
// One B&B node (comp, state) = the reference's sequential viterbi_from calls for
// every position of the component, executed as parallel phases with bit-identical
// result: (A) reset the clamped rows, (B) sweep every clamp-to-next-
// fixed-position segment concurrently, (C1/C2) backpointer fix-ups, (D) bound terms
// gathered in parallel and summed serially in the reference's order.
//
// Sweep kernel mapping: one warp per segment, lanes = target states (chain_warp.cuh) -- a node's
// segments are many but short and the longest one bounds the node's latency, so the per-step
// latency matters more than lane efficiency.  Arithmetic is the reference's R2 order (cp.rs:49-58):
//   psi = first-argmax_j fl(delta[t-1][j] + tr_j),  tr_j = a[j][s]  (pi[s] if el.t == 0)
//   delta[t][s] = fl(delta[t-1][psi] + fl(tr_psi + b[s][o_t]))
// so the cell keeps (value, index) -- the index IS state here, unlike mode R1.
#pragma once

#include "common.cuh"
#include "chain_warp.cuh"
#include "decode_chain.cuh"

namespace cvb {

typedef uint16_t psi_t;   // backpointer type of the constrained path (K <= 65535)

struct CpParams {
    const double *A;          // [K][Kp] padded -inf
    const double *BT;         // [M][Kp]
    const double *Pi;         // [Kp]
    const uint32_t *obs;      // [N]
    const uint8_t *start;     // [N] el.t == 0
    const int32_t *comp;      // [N] active component or -1
    double *delta;            // [N][K]
    psi_t *psi;               // [N][K]
    int32_t *choice;          // [ncomp] cstr_choices, -1 = None
    int64_t N, M;
    int K, Kp, G;
    int bt_in_smem;           // logB^T staged in shared memory by the warp-per-segment sweep
};

struct CpSweepArgs {
    const int64_t *seg_from;  // [nseg] start row of each sweep, longest sweep first
    const int32_t *seg_len;   // [nseg] number of rows swept after seg_from
    int nseg, ntiles;
    int node;                 // clamped state (ignored when init_mode)
    int init_mode;            // 1: rows start from delta[seg_from] as stored (init_viterbi prefix)
    unsigned int *tile_counter;
    // Leaf batch (cp_host.inl, cp_leaf_group): the K sibling nodes of the last component sweep the same segments
    // and overwrite the same rows, so their bounds can be evaluated together.  leaf_nsib > 0: task r = (sibling
    // r / nseg, segment r % nseg), clamped state = sibling; delta / psi in global memory are NOT touched, only the
    // last delta row of each sweep is kept: leaf_last[(sibling * nseg + segment) * K + i].
    int leaf_nsib;
    double *leaf_last;
};

__host__ __device__ inline size_t cp_sweep_smem_bytes(int K, int Kp)
{
    return (size_t)K * Kp * 8 + (size_t)2 * K * 64 * 8 + (size_t)Kp * 8 + 64 * (8 + 4) + 16;
}

// delta[0][i] = fl(pi[i] + b[i][o_0])   (init_probs, hmm.rs:215-218; cp.rs:66-68)
__global__ void cp_row0_kernel(const CpParams p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p.K) p.delta[i] = p.Pi[i] + p.BT[(size_t)p.obs[0] * p.Kp + i];
}

// ---- sweeps, one warp per segment (latency-oriented; see chain_warp.cuh) -----------------------------
// Phases A (reset row) + B (sweep); segments are listed longest first.
constexpr int CPW_WARPS = 4;

// LPS = lanes per segment: 32, or 16 when K <= 16 (two segments per warp, adjacent in the length-sorted list).
template <int NSL, int KQ, int LPS = 32>
__global__ void __launch_bounds__(32 * CPW_WARPS) cp_sweep_chain_kernel(const CpParams p, const CpSweepArgs a)
{
    static_assert(LPS == 32 || (LPS == 16 && NSL == 1), "half-warp segments need one state per lane");
    constexpr int GP = 32 / LPS;                                          // segments per warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K, Kp = p.Kp;
    double *sA = reinterpret_cast<double *>(smem_raw);
    double *sBT = sA + (size_t)K * Kp;
    const size_t nbt = p.bt_in_smem ? (size_t)p.M * Kp : 0;
    const int lane = threadIdx.x & 31, sub = lane & (LPS - 1), grp = lane / LPS;
    for (int e = threadIdx.x; e < K * Kp; e += blockDim.x) sA[e] = p.A[e];
    for (size_t e = threadIdx.x; e < nbt; e += blockDim.x) sBT[e] = p.BT[e];
    __syncthreads();
    const double *bt = p.bt_in_smem ? sBT : p.BT;
    const bool pf = !p.bt_in_smem;
    double *sdw = sBT + nbt + (size_t)((threadIdx.x >> 5) * GP + grp) * 2 * Kp;   // this segment's delta rows [2][Kp]

    double pi_i[NSL]; int col[NSL];
#pragma unroll
    for (int s = 0; s < NSL; s++) { col[s] = min(sub + 32 * s, Kp - 1); pi_i[s] = p.Pi[col[s]]; }
    double acol[KQ > 0 ? 4 * KQ : 1];
    if (KQ > 0) {
#pragma unroll
        for (int j = 0; j < 4 * KQ; j++) acol[j] = (j < K && sub < Kp) ? sA[(size_t)j * Kp + sub] : neg_inf();
    }

    // Segments are sorted longest first and dealt to the warps round-robin (static: no claim round trip); the next
    // segment's start / length are fetched while the current one is swept.
    const int stride = (int)gridDim.x * CPW_WARPS * GP;
    const bool leaf = a.leaf_nsib > 0;
    const int ntask = leaf ? a.nseg * a.leaf_nsib : a.nseg;                   // host keeps nseg * nsib < 2^31
    int r = ((int)blockIdx.x * CPW_WARPS + (int)(threadIdx.x >> 5)) * GP + grp;
    int64_t from_n = 0; int len_n = -1;
    if (r < ntask) { const int g = leaf ? r % a.nseg : r; from_n = __ldg(a.seg_from + g); len_n = __ldg(a.seg_len + g); }
    for (;; r += stride) {
        if (r - grp >= ntask) break;                                          // warp-uniform
        const bool valid = r < ntask;
        const int64_t from = from_n;
        const int len = len_n;
        const int node = leaf ? r / a.nseg : a.node;
        from_n = 0; len_n = -1;
        if (r + stride < ntask) { const int g = leaf ? (r + stride) % a.nseg : r + stride; from_n = __ldg(a.seg_from + g); len_n = __ldg(a.seg_len + g); }
        const int lenmax = __reduce_max_sync(0xffffffffu, len);

        double d[NSL];
#pragma unroll
        for (int s = 0; s < NSL; s++) {
            const int i = sub + 32 * s;
            d[s] = neg_inf();
            if (valid) {
                if (a.init_mode) {
                    d[s] = (i < K) ? p.delta[(size_t)from * K + i] : neg_inf();
                } else {
                    d[s] = (i == node) ? 0.0 : neg_inf();                        // cp.rs:33-34
                    if (i < K && !leaf) p.delta[(size_t)from * K + i] = d[s];
                }
            }
            if (i < Kp) { sdw[i] = (i < K) ? d[s] : neg_inf(); sdw[Kp + i] = neg_inf(); }
        }
        __syncwarp();
        auto load_meta = [&](int k, uint32_t &o, bool &st) {
            o = 0u; st = false;
            if (k <= len) { o = __ldg(p.obs + from + k); st = __ldg(p.start + from + k) != 0; }
        };
        // rings: oq[q] / sq[q] = observation / sequence-start flag of row from + 2 + q
        uint32_t o1, oq[CHAIN_PF + 1]; bool st1, sq[CHAIN_PF + 1];
        load_meta(1, o1, st1);
#pragma unroll
        for (int q = 0; q <= CHAIN_PF; q++) load_meta(2 + q, oq[q], sq[q]);
        double e_next[NSL];
#pragma unroll
        for (int s = 0; s < NSL; s++) e_next[s] = bt[(size_t)o1 * Kp + col[s]];
        if (pf) {
#pragma unroll
            for (int q = 0; q < CHAIN_PF; q++) prefetch_l1(p.BT + (size_t)oq[q] * Kp + col[0]);
        }

        for (int k = 1; k <= lenmax; k++) {                                       // cp.rs:47-60 / 70-78
            const bool act = k <= len;
            const int64_t t = from + k;
            const bool st = st1;
            double e[NSL];
#pragma unroll
            for (int s = 0; s < NSL; s++) { e[s] = e_next[s]; e_next[s] = bt[(size_t)oq[0] * Kp + col[s]]; }
            if (pf) prefetch_l1(p.BT + (size_t)oq[CHAIN_PF] * Kp + col[0]);
            st1 = sq[0];
#pragma unroll
            for (int q = 0; q < CHAIN_PF; q++) { oq[q] = oq[q + 1]; sq[q] = sq[q + 1]; }
            load_meta(k + 2 + CHAIN_PF, oq[CHAIN_PF], sq[CHAIN_PF]);
            double best[NSL]; int idx[NSL];
            const double *sdo = sdw + ((k - 1) & 1) * Kp;
            if (KQ > 0 && !st) chain_scan_reg<(KQ > 0 ? KQ : 2)>(sdo, reinterpret_cast<const double (&)[4 * (KQ > 0 ? KQ : 2)]>(acol), best[0], idx[0]);
            else chain_scan<NSL>(sdo, sA, Kp, K, sub, st, pi_i, best, idx);       // argmax on delta + tr
            double v[NSL];
#pragma unroll
            for (int s = 0; s < NSL; s++) {
                const double tr = st ? pi_i[s] : sA[(size_t)idx[s] * Kp + col[s]];
                const double arc = tr + e[s];                                     // arc_p (utils.rs:24-30)
                v[s] = sdo[idx[s]] + arc;                                         // delta + (a + b)  cp.rs:55
            }
#pragma unroll
            for (int s = 0; s < NSL; s++) {
                const int i = sub + 32 * s;
                if (act && i < K) {
                    if (!leaf) {
                        p.delta[(size_t)t * K + i] = v[s];
                        p.psi[(size_t)t * K + i] = (psi_t)idx[s];
                    }
                    sdw[(k & 1) * Kp + i] = v[s];
                    d[s] = v[s];
                }
            }
            __syncwarp();
        }
        if (leaf && valid) {                                                      // the sweep's last row (the reset row when len = 0)
            double *out = a.leaf_last + ((size_t)node * a.nseg + (size_t)(r % a.nseg)) * K;
#pragma unroll
            for (int s = 0; s < NSL; s++) { const int i = sub + 32 * s; if (i < K) out[i] = d[s]; }
        }
    }
}

// ---- sweeps for K > 64: one warp per segment, every lane loops over its states i = lane, lane+32, ... ----
// Same phases and arithmetic as cp_sweep_chain_kernel; logA / logB^T are read from global memory (L2), the
// delta rows of the warp live in shared memory.  A correctness-first path: constrained decodes with more than 64
// states are outside BASELINE's configs.
constexpr int CPG_WARPS = 4;

__global__ void __launch_bounds__(32 * CPG_WARPS) cp_sweep_generic_kernel(const CpParams p, const CpSweepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K, Kp = p.Kp;
    const int lane = threadIdx.x & 31;
    double *sdw = reinterpret_cast<double *>(smem_raw) + (size_t)(threadIdx.x >> 5) * 2 * Kp;
    const int nsl = (K + 31) / 32;
    for (;;) {
        unsigned int r = 0;
        if (lane == 0) r = atomicAdd(a.tile_counter, 1u);
        r = __shfl_sync(0xffffffffu, r, 0);
        if ((int)r >= a.nseg) break;
        const int64_t from = a.seg_from[r];
        const int len = a.seg_len[r];
        for (int i = lane; i < Kp; i += 32) {
            double v = neg_inf();
            if (i < K) {
                if (a.init_mode) v = p.delta[(size_t)from * K + i];
                else { v = (i == a.node) ? 0.0 : neg_inf(); p.delta[(size_t)from * K + i] = v; }   // cp.rs:33-34
            }
            sdw[i] = v; sdw[Kp + i] = neg_inf();
        }
        __syncwarp();
        for (int k = 1; k <= len; k++) {                                          // cp.rs:47-60 / 70-78
            const int64_t t = from + k;
            const bool st = p.start[t] != 0;
            const uint32_t o = p.obs[t];
            const double *sdo = sdw + ((k - 1) & 1) * Kp;
            double *sdn = sdw + (k & 1) * Kp;
            for (int s = 0; s < nsl; s++) {
                const int i = lane + 32 * s;                                      // warp-uniform trip count
                const int ic = min(i, Kp - 1);
                const double pi_i[1] = {p.Pi[ic]};
                double best[1]; int idx[1];
                chain_scan<1>(sdo, p.A, Kp, K, i, st, pi_i, best, idx);           // argmax on delta + tr
                const double tr = st ? pi_i[0] : p.A[(size_t)idx[0] * Kp + ic];
                const double arc = tr + p.BT[(size_t)o * Kp + ic];                // arc_p (utils.rs:24-30)
                const double v = sdo[idx[0]] + arc;                               // delta + (a + b)  cp.rs:55
                if (i < K) {
                    p.delta[(size_t)t * K + i] = v;
                    p.psi[(size_t)t * K + i] = (psi_t)idx[0];
                    sdn[i] = v;
                }
            }
            __syncwarp();
        }
    }
}

// Phases C1 + C2 for the positions of the component just assigned (cp.rs:35-45).
//  C2: psi[pos][state] = first-argmax_j fl(delta[pos-1][j] + tr_j(state))          (pos != 0)
//  C1: psi[pos+1][choice[comp[pos+1]]] = state when pos+1 is fixed.  If pos+1 belongs to the same component
//      the reference overwrites that very entry in the next viterbi_from call (C2 of pos+1), so it is skipped.
// A shard (cp_dist.cuh) does C1 for the positions it owns, lo <= pos < hi, and C2 where it owns row pos-1,
// lo < pos <= hi; a single rank passes lo = 0, hi = N.
__global__ void cp_fixup_kernel(const CpParams p, const int64_t *pos_list, int npos, int comp, int state, int64_t lo, int64_t hi)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) p.choice[comp] = state;        // cstr_choices[comp] = Some(state) (cp.rs:98); read by the terms kernel only
    if (k >= npos) return;
    const int64_t pos = pos_list[k];
    const int K = p.K, Kp = p.Kp;
    if (pos != 0 && pos > lo) {
        const double *prev = p.delta + (size_t)(pos - 1) * K;
        const bool st = p.start[pos] != 0;
        double bv = 0.0; int bi = 0;
        for (int j = 0; j < K; j++) {
            const double v = prev[j] + (st ? p.Pi[state] : p.A[(size_t)j * Kp + state]);
            if (j == 0 || v > bv) { bv = v; bi = j; }
        }
        p.psi[(size_t)pos * K + state] = (psi_t)bi;
    }
    if (pos + 1 < p.N && pos < hi) {
        const int c1 = p.comp[pos + 1];
        if (c1 >= 0 && c1 < comp) p.psi[(size_t)(pos + 1) * K + p.choice[c1]] = (psi_t)state;
    }
}

// ---- where a node's bound goes, and whom the sum has to wait for --------------------------------------------
// UbSink: the sum kernels store ub in device memory and, when `host` is set, in pinned host memory followed by
// the node's sequence number -- the host polls that word instead of paying a copy + stream synchronisation
// per node (host[0] = ub, host[1] = peer-timeout flag, host[2] = sequence number as u64).
struct UbSink {
    double *dev; double *host; const int *err; unsigned long long seq;
};
__device__ __forceinline__ void ub_store(const UbSink &o, double ub)
{
    *o.dev = ub;
    if (o.host) {
        o.host[0] = ub;
        o.host[1] = (o.err && *o.err) ? 1.0 : 0.0;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(o.host + 2) = o.seq;
    }
}
// PeerWait: sharded solve (cp_dist.cuh) -- every rank's terms of exchange `epoch` must have arrived (flags[q] >=
// epoch) before the list is read.  R <= 1: nothing to wait for.  Bounded (~4 s), then *err = 1.
struct PeerWait {
    const unsigned int *flags; int R; unsigned int epoch; int *err;
};
__device__ __forceinline__ void peer_wait_block(const PeerWait &w)
{
    if (w.R <= 1) return;
    if ((int)threadIdx.x < w.R) {
        const long long t0 = clock64();
        for (;;) {
            unsigned int v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(w.flags + threadIdx.x) : "memory");
            if ((int)(v - w.epoch) >= 0) break;
            if (clock64() - t0 > (1LL << 33)) { *w.err = 1; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

// Batched form (all sum kernels): blockIdx.y selects list y of a batch -- terms + y * batch.term_stride, block
// workspace + y * batch.blk_stride, result to ub_out.dev[y]; a single list is batch = {0, 0}, gridDim.y = 1.
struct SumBatch { long long term_stride, blk_stride; };
__device__ __forceinline__ UbSink ub_sink_of(const UbSink &o) { UbSink r = o; r.dev = o.dev + blockIdx.y; return r; }

// ---- block statistics for the block-structured exact sum (further down) -------------------------------------
// Every SUM_BLK consecutive terms form a block; a block's plain f64 sum (any order -- it only PREDICTS the
// binade of the running sum) and its special-value flags (1 = positive or NaN term, 2 = -inf term).
constexpr int SUM_BLK = 128;

__device__ __forceinline__ void sum_block_stats(double x, int block, double *bsum, int *bflag)
{
    __shared__ double ws[SUM_BLK / 32]; __shared__ int wf[SUM_BLK / 32];
    int flag = 0;
    if (!(x <= 0.0)) flag = 1; else if (x == neg_inf()) flag = 2;
    double v = x;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    flag = __reduce_or_sync(0xffffffffu, flag);
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = v; wf[threadIdx.x >> 5] = flag; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0; int f = 0;
#pragma unroll
        for (int w = 0; w < SUM_BLK / 32; w++) { t += ws[w]; f |= wf[w]; }
        bsum[block] = t; bflag[block] = f;
    }
}

// Phase D: one bound term per clamped position of components 0..comp, in the reference's order
// (cid ascending, position ascending; cp.rs:104-116).  term_pos[k] / term_comp[k] give position and component.
// Launched with SUM_BLK threads per block; also leaves the block statistics of the exact sum.
__global__ void __launch_bounds__(SUM_BLK) cp_terms_kernel(const CpParams p, const int64_t *term_pos, const int32_t *term_comp,
                                                         int nterms, double *terms, double *bsum, int *bflag)
{
    const int k = blockIdx.x * SUM_BLK + threadIdx.x;
    double term = 0.0;
    if (k < nterms) {
        const int64_t t = term_pos[k];
        const int st = p.choice[term_comp[k]];
        const int K = p.K, Kp = p.Kp;
        const double b = p.BT[(size_t)p.obs[t] * Kp + st];
        if (t == 0) {
            term = p.Pi[st] + b;                               // sequence[0].arc_p(hmm, 0, state), el.t == 0
        } else {
            const int sf = p.psi[(size_t)t * K + st];
            const double arc = (p.start[t] ? p.Pi[st] : p.A[(size_t)sf * Kp + st]) + b;
            term = p.delta[(size_t)(t - 1) * K + sf] + arc;
        }
        terms[k] = term;
    }
    sum_block_stats(term, blockIdx.x, bsum, bflag);
}

// ---- leaf batch: the bound terms of all sibling nodes of the last component in one launch ------------------------
// Sibling s (component `last` clamped to state s; components < last as in p.choice) would, node by node, run its
// sweeps, the C1/C2 fix-ups and cp_terms_kernel.  Here nothing is written to delta / psi: what the sibling's sweeps
// WOULD have left in the rows a term reads comes from leaf_last (cp_sweep_chain_kernel, leaf mode):
//   row t-1 of a clamped position t is the last row of the sweep that starts at the nearest clamped position below
//   t; prev_seg[k] = that sweep's index when it belongs to the last component (sibling-specific), else -1 (the row
//   in global memory is what every sibling sees);
//   psi[t][s_c]: C2 when t itself is a position of the last component (first-argmax over row t-1, cp.rs:35-41),
//   C1 when t-1 is one (psi[t][choice] = s, cp.rs:43-45), else the stored entry.
// The C2 results are kept (c2_out[s * nleaf + leaf_idx[k]]) and written to psi later, sibling by sibling in the
// reference's order (cp_leaf_apply_c2_kernel).  grid = (blocks of SUM_BLK terms, siblings).
struct CpLeafArgs {
    const int64_t *term_pos; const int32_t *term_comp;   // [nterms] as for cp_terms_kernel
    const int32_t *prev_seg;                              // [nterms] leaf sweep that owns row t-1, or -1
    const int32_t *leaf_idx;                              // [nterms] index among the last component's positions, or -1
    const double *leaf_last;                              // [nsib][nseg][K]
    psi_t *c2_out;                                        // [nsib][nleaf]
    double *terms; double *bsum; int *bflag;              // [nsib][term_stride], [nsib][blk_stride] x 2
    int nterms, nseg, nleaf, last;
    long long term_stride, blk_stride;
};

__global__ void __launch_bounds__(SUM_BLK) cp_leaf_terms_kernel(const CpParams p, const CpLeafArgs a)
{
    const int k = blockIdx.x * SUM_BLK + threadIdx.x, sib = blockIdx.y;
    double term = 0.0;
    if (k < a.nterms) {
        const int64_t t = a.term_pos[k];
        const int c = a.term_comp[k];
        const int st = c == a.last ? sib : p.choice[c];
        const int K = p.K, Kp = p.Kp;
        const double b = p.BT[(size_t)p.obs[t] * Kp + st];
        if (t == 0) {
            term = p.Pi[st] + b;
        } else {
            const int g = a.prev_seg[k];
            const double *prev = g >= 0 ? a.leaf_last + ((size_t)sib * a.nseg + (size_t)g) * K : p.delta + (size_t)(t - 1) * K;
            const bool stt = p.start[t] != 0;
            int sf;
            if (c == a.last) {                               // C2 (cp_fixup_kernel): psi[t][sib] = first-argmax_j(prev[j] + tr_j(sib))
                double bv = 0.0; int bi = 0;
                for (int j = 0; j < K; j++) {
                    const double v = prev[j] + (stt ? p.Pi[sib] : p.A[(size_t)j * Kp + sib]);
                    if (j == 0 || v > bv) { bv = v; bi = j; }
                }
                sf = bi;
                a.c2_out[(size_t)sib * a.nleaf + a.leaf_idx[k]] = (psi_t)bi;
            } else if (p.comp[t - 1] == a.last) {            // C1: psi[t][choice[c]] = sib, and st = choice[c]
                sf = sib;
            } else {
                sf = p.psi[(size_t)t * K + st];
            }
            const double arc = (stt ? p.Pi[st] : p.A[(size_t)sf * Kp + st]) + b;
            term = prev[sf] + arc;
        }
        a.terms[(size_t)sib * a.term_stride + k] = term;
    }
    sum_block_stats(term, blockIdx.x, a.bsum + (size_t)sib * a.blk_stride, a.bflag + (size_t)sib * a.blk_stride);
}

// psi[pos][s] = C2 result of sibling s for s in [s0, s1) at every position of the last component (pos != 0):
// the entries the siblings' viterbi_from calls leave behind, one column each (cp.rs:35-41).
__global__ void cp_leaf_apply_c2_kernel(const CpParams p, const int64_t *leaf_pos, int nleaf, const psi_t *c2, int s0, int s1)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nleaf) return;
    const int64_t pos = leaf_pos[i];
    if (pos == 0) return;
    for (int s = s0; s < s1; s++) p.psi[(size_t)pos * p.K + s] = c2[(size_t)s * nleaf + i];
}

// block statistics of an existing term list (debug hook / lists not produced by cp_terms_kernel)
__global__ void __launch_bounds__(SUM_BLK) cp_sum_stats_kernel(const double *terms, int nterms, double *bsum, int *bflag,
                                                             const PeerWait pw, const SumBatch batch)
{
    terms += (size_t)blockIdx.y * batch.term_stride;
    bsum += (size_t)blockIdx.y * batch.blk_stride; bflag += (size_t)blockIdx.y * batch.blk_stride;
    peer_wait_block(pw);
    const int k = blockIdx.x * SUM_BLK + threadIdx.x;
    sum_block_stats(k < nterms ? terms[k] : 0.0, blockIdx.x, bsum, bflag);
}

// ub = ((0.0 + term_0) + term_1) + ...  exactly in order (cp.rs:103,109,114).  The order is part of the result,
// so one thread performs the adds; the rest of the block streams the terms through a double-buffered shared
// memory stage and the adder keeps 16 terms in registers ahead of the dependent DADD chain (8 clk per term).
__global__ void __launch_bounds__(256) cp_sum_kernel(const double *terms, int nterms, const UbSink ub_out_,
                                                     unsigned int *reset_counter, const PeerWait pw, const SumBatch batch)
{
    terms += (size_t)blockIdx.y * batch.term_stride;
    const UbSink ub_out = ub_sink_of(ub_out_);
    peer_wait_block(pw);
    if (threadIdx.x == 0 && reset_counter) *reset_counter = 0u;
    constexpr int CH = 2048;
    __shared__ double buf[2][CH];
    double ub = 0.0;
    const int nchunks = (nterms + CH - 1) / CH;
    for (int e = threadIdx.x; e < min(CH, nterms); e += blockDim.x) buf[0][e] = terms[e];
    __syncthreads();
    for (int c = 0; c < nchunks; c++) {
        const int n = min(CH, nterms - c * CH);
        if (threadIdx.x == 0) {
            const double *b = buf[c & 1];
            int e = 0;
            for (; e + 16 <= n; e += 16) {
                double x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = b[e + k];
#pragma unroll
                for (int k = 0; k < 16; k++) ub += x[k];
            }
            for (; e < n; e++) ub += b[e];
        } else if (c + 1 < nchunks) {
            const int base = (c + 1) * CH, m = min(CH, nterms - base);
            for (int e = threadIdx.x - 1; e < m; e += blockDim.x - 1) buf[(c + 1) & 1][e] = terms[base + e];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) ub_store(ub_out, ub);
}


// ---- exact-order sum, parallel ------------------------------------------------------------------------
// ub = ((0.0 + x_0) + x_1) + ... with one IEEE round-to-nearest-even per add is a serial recurrence, but while
// the running sum s stays inside one binade [2^e, 2^(e+1)) (ulp u = 2^(e-52)) and every term is <= 0, each add
// is an INTEGER update of Q = |s| / u:  Q' = Q + I + [phi > 1/2], where |x| / u = I + phi, and on an exact tie
// (phi = 1/2) the increment is decided by the parity of Q + I (ties-to-even).  A term is therefore a function
// Q -> Q + (Q even ? a0 : a1); such functions are closed under composition, so a parallel scan yields every
// partial sum of the binade exactly.  The scan runs over windows of 8192 terms; at the first add whose result
// leaves the binade (Q' >= 2^53) that single add is done in floating point and the next binade starts.
// Bit-identical to the serial loop (checked against it on the GPU and against a Python model); inputs with a
// positive term, a NaN or a subnormal running sum take the serial loop instead.
struct QFn { unsigned long long a0, a1; };
constexpr unsigned long long QSAT = 1ULL << 60;

__device__ __forceinline__ QFn qfn_compose(const QFn a, const QFn b)    // a first, then b
{
    QFn r;
    r.a0 = a.a0 + ((a.a0 & 1ULL) == 0 ? b.a0 : b.a1);
    r.a1 = a.a1 + (((1ULL + a.a1) & 1ULL) == 0 ? b.a0 : b.a1);
    r.a0 = min(r.a0, QSAT); r.a1 = min(r.a1, QSAT);
    return r;
}
__device__ __forceinline__ unsigned long long qfn_apply(const QFn f, unsigned long long Q)
{
    return min(Q + ((Q & 1ULL) == 0 ? f.a0 : f.a1), QSAT * 2);
}
// term x <= 0 (finite) seen from binade e
__device__ __forceinline__ QFn qfn_elem(double x, int e)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    const int ef = (int)((b >> 52) & 0x7ff);
    unsigned long long M = b & ((1ULL << 52) - 1);
    int E = -1074;
    if (ef) { M |= 1ULL << 52; E = ef - 1075; }
    QFn r; r.a0 = r.a1 = 0;
    if (M == 0) return r;
    const int sh = E - (e - 52);
    if (sh >= 0) { r.a0 = r.a1 = (sh > 6) ? QSAT : (M << sh); return r; }
    const int n = -sh;
    if (n >= 64) return r;
    const unsigned long long I = M >> n, F = M & ((1ULL << n) - 1), half = 1ULL << (n - 1);
    if (F != half) { r.a0 = r.a1 = I + (F > half ? 1ULL : 0ULL); return r; }
    r.a0 = I + (I & 1ULL); r.a1 = I + ((I + 1ULL) & 1ULL);
    return r;
}

constexpr int QS_THREADS = 1024, QS_EPT = 8;
constexpr int QS_SERIAL_HEAD = 2048;                // leading terms summed by the plain loop

__global__ void __launch_bounds__(QS_THREADS) cp_sum_exact_kernel(const double *terms, int nterms, const UbSink ub_out_,
                                                                  unsigned int *reset_counter, const SumBatch batch)
{
    terms += (size_t)blockIdx.y * batch.term_stride;
    const UbSink ub_out = ub_sink_of(ub_out_);
    if (threadIdx.x == 0 && reset_counter) *reset_counter = 0u;
    __shared__ double s_sh; __shared__ int pos_sh, cross_sh, mode_sh;
    __shared__ unsigned long long qbefore_sh, qend_sh;
    __shared__ QFn warp_agg[QS_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    // pre-check: serial fallback for positive / NaN terms; -inf terms make the sum -inf
    if (tid == 0) mode_sh = 0;
    __syncthreads();
    int flag = 0;
    for (int k = tid; k < nterms; k += QS_THREADS) {
        const double x = terms[k];
        if (!(x <= 0.0)) flag |= 1;                 // positive or NaN
        else if (x == neg_inf()) flag |= 2;
    }
    if (flag) atomicOr(&mode_sh, flag);
    __syncthreads();
    const int mode = mode_sh;
    if (mode & 1) {                                 // reference loop, one thread
        if (tid == 0) { double ub = 0.0; for (int k = 0; k < nterms; k++) ub += terms[k]; ub_store(ub_out, ub); }
        return;
    }
    if (mode & 2) { if (tid == 0) ub_store(ub_out, neg_inf()); return; }

    // The first terms go through the plain loop: while the sum is small its binade changes every few adds, and a
    // binade change costs a whole scan window.  After QS_SERIAL_HEAD terms a change needs ~pos more terms.
    {
        double *head = reinterpret_cast<double *>(warp_agg);    // reuse: 32 x 16 B = 64 doubles per round
        double s = 0.0;
        const int nhead = min(nterms, QS_SERIAL_HEAD);
        for (int base = 0; base < nhead; base += 64) {
            if (tid < 64 && base + tid < nhead) head[tid] = terms[base + tid];
            __syncthreads();
            if (tid == 0) {
                const int n = min(64, nhead - base);
#pragma unroll 8
                for (int k = 0; k < n; k++) s = s + head[k];
            }
            __syncthreads();
        }
        if (tid == 0) {
            int pos = nhead;
            while (pos < nterms && s == 0.0) { s = s + terms[pos]; pos++; }  // 0.0 + x is exact
            s_sh = s; pos_sh = pos;
        }
    }
    for (;;) {
        __syncthreads();
        const double s = s_sh; const int pos = pos_sh;
        if (pos >= nterms) break;
        const unsigned long long sb = (unsigned long long)__double_as_longlong(s);
        const int ef = (int)((sb >> 52) & 0x7ff);
        if (ef == 0) {                              // subnormal running sum: finish with the reference loop
            if (tid == 0) { double ub = s; for (int k = pos; k < nterms; k++) ub += terms[k]; s_sh = ub; pos_sh = nterms; }
            continue;
        }
        const int e = ef - 1023;
        const unsigned long long Q0 = (sb & ((1ULL << 52) - 1)) | (1ULL << 52);
        const int W = min(nterms - pos, QS_THREADS * QS_EPT);
        if (tid == 0) cross_sh = 0x7fffffff;
        // local composition of this thread's contiguous terms
        QFn el[QS_EPT]; QFn f; f.a0 = f.a1 = 0;
        const int base = pos + tid * QS_EPT;
#pragma unroll
        for (int k = 0; k < QS_EPT; k++) {
            el[k].a0 = el[k].a1 = 0;
            if (tid * QS_EPT + k < W) el[k] = qfn_elem(terms[base + k], e);
            f = qfn_compose(f, el[k]);
        }
        // block-wide exclusive scan of the composed functions
        QFn inc = f;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            QFn o;
            o.a0 = __shfl_up_sync(0xffffffffu, inc.a0, d);
            o.a1 = __shfl_up_sync(0xffffffffu, inc.a1, d);
            if (lane >= d) inc = qfn_compose(o, inc);
        }
        if (lane == 31) warp_agg[w] = inc;
        __syncthreads();
        if (w == 0) {
            QFn a = warp_agg[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                QFn o;
                o.a0 = __shfl_up_sync(0xffffffffu, a.a0, d);
                o.a1 = __shfl_up_sync(0xffffffffu, a.a1, d);
                if (lane >= d) a = qfn_compose(o, a);
            }
            warp_agg[lane] = a;                     // inclusive over warps
        }
        __syncthreads();
        QFn pre; pre.a0 = pre.a1 = 0;               // everything before this thread
        if (w > 0) pre = warp_agg[w - 1];
        {
            QFn ex;                                  // exclusive within the warp
            ex.a0 = __shfl_up_sync(0xffffffffu, inc.a0, 1);
            ex.a1 = __shfl_up_sync(0xffffffffu, inc.a1, 1);
            if (lane > 0) pre = qfn_compose(pre, ex);
        }
        unsigned long long Q = qfn_apply(pre, Q0);
        // walk own terms, look for the first add that leaves the binade
        int my_cross = 0x7fffffff; unsigned long long q_at_cross = 0;
#pragma unroll
        for (int k = 0; k < QS_EPT; k++) {
            if (tid * QS_EPT + k < W && my_cross == 0x7fffffff) {
                const unsigned long long Qn = qfn_apply(el[k], Q);
                if (Qn >= (1ULL << 53)) { my_cross = base + k; q_at_cross = Q; }
                else Q = Qn;
            }
        }
        if (my_cross != 0x7fffffff) atomicMin(&cross_sh, my_cross);
        if (tid == (W - 1) / QS_EPT) qend_sh = Q;   // the thread owning the last term of the window
        __syncthreads();
        const int cross = cross_sh;
        if (cross != 0x7fffffff && my_cross == cross) qbefore_sh = q_at_cross;
        __syncthreads();
        if (tid == 0) {
            const unsigned long long Qf = (cross == 0x7fffffff) ? qend_sh : qbefore_sh;
            const unsigned long long bits = (1ULL << 63) | ((unsigned long long)(e + 1023) << 52) | (Qf & ((1ULL << 52) - 1));
            double sk = __longlong_as_double((long long)bits);               // -Qf * 2^(e-52), exact
            if (cross == 0x7fffffff) { s_sh = sk; pos_sh = pos + W; }
            else { s_sh = sk + terms[cross]; pos_sh = cross + 1; }          // the one add that changes binade
        }
    }
    if (tid == 0) ub_store(ub_out, s_sh);
}

// ---- exact-order sum, block-structured (many CTAs) ------------------------------------------------------
// The same binade argument as above, arranged so that almost all work is parallel over CTAs:
//  (1) cp_terms_kernel / cp_sum_stats_kernel leave a plain sum per block of SUM_BLK terms;
//  (2) cp_sum_blockfn_kernel, one CTA per block: an approximate prefix (sum of the earlier block sums) predicts
//      the binade e of the running sum while it passes through the block; if the prediction is safe (prefix and
//      prefix + block sum in the same binade with a 2^-30 relative margin) the block's terms are composed into
//      one function F_b : Q -> Q + (Q even ? a0 : a1) valid for that binade, else the block is marked open;
//  (3) cp_sum_chain_kernel, one CTA: warps compose runs of 32 same-binade blocks into group functions, then
//      one warp walks the groups in order.  A function is APPLIED only if the actual running sum is in the
//      binade it was built for and the result stays inside it (Q' < 2^53; the sum is monotone because every
//      term is <= 0), so a wrong prediction costs time, never exactness; open blocks and failed applications
//      are summed term by term in floating point (SUM_BLK dependent adds).
// Special values as in cp_sum_exact_kernel: a positive/NaN term => the plain loop over everything, -inf => -inf.
constexpr int SUM_OPEN = -100000;                     // "no function for this block / group"
constexpr int SUM_MAX_BLOCKS = 1 << 16;               // above this (8M terms) the single-CTA kernel is used
constexpr int SUM_TILE_BLOCKS = 4096;                 // blocks the chain kernel stages in shared memory at a time
constexpr int SUM_PREFIX_MIN = 2048;                  // from here on the approximate prefixes come from their own kernel

// approximate exclusive prefix of the block sums for long lists (short ones sum the earlier blocks inside
// cp_sum_blockfn_kernel): every thread a contiguous stretch, then a scan of the 1024 partial sums
__global__ void __launch_bounds__(1024) cp_sum_prefix_kernel(const double *bsum, int nblk, double *bpre, const SumBatch batch)
{
    bsum += (size_t)blockIdx.y * batch.blk_stride; bpre += (size_t)blockIdx.y * batch.blk_stride;
    __shared__ double part[1024];
    const int tid = threadIdx.x, per = (nblk + 1023) / 1024, lo = min(tid * per, nblk), hi = min(lo + per, nblk);
    double acc = 0.0;
    for (int i = lo; i < hi; i++) acc += bsum[i];
    part[tid] = acc;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const double v = tid >= d ? part[tid - d] : 0.0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    double run = tid ? part[tid - 1] : 0.0;
    for (int i = lo; i < hi; i++) { bpre[i] = run; run += bsum[i]; }
}

__global__ void __launch_bounds__(SUM_BLK) cp_sum_blockfn_kernel(const double *terms, int nterms, const double *bsum,
                                                               const double *bpre, int *bexp, QFn *bfn, const SumBatch batch)
{
    terms += (size_t)blockIdx.y * batch.term_stride;
    bsum += (size_t)blockIdx.y * batch.blk_stride; bexp += (size_t)blockIdx.y * batch.blk_stride; bfn += (size_t)blockIdx.y * batch.blk_stride;
    if (bpre) bpre += (size_t)blockIdx.y * batch.blk_stride;
    __shared__ double ws[SUM_BLK / 32]; __shared__ int e_sh; __shared__ QFn wagg[SUM_BLK / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    double acc = 0.0;
    if (bpre) { if (tid == 0) acc = bpre[b]; }
    else for (int i = tid; i < b; i += SUM_BLK) acc += bsum[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) ws[w] = acc;
    __syncthreads();
    if (tid == 0) {
        double P = 0.0;
#pragma unroll
        for (int k = 0; k < SUM_BLK / 32; k++) P += ws[k];
        const double lo = fabs(P) * (1.0 - 9.313225746154785e-10), hi = fabs(P + bsum[b]) * (1.0 + 9.313225746154785e-10);
        const int elo = (int)(((unsigned long long)__double_as_longlong(lo) >> 52) & 0x7ff);
        const int ehi = (int)(((unsigned long long)__double_as_longlong(hi) >> 52) & 0x7ff);
        e_sh = (lo > 0.0 && elo == ehi && elo != 0 && elo != 0x7ff) ? elo - 1023 : SUM_OPEN;
    }
    __syncthreads();
    const int e = e_sh;
    if (e == SUM_OPEN) { if (tid == 0) bexp[b] = SUM_OPEN; return; }
    const int k = b * SUM_BLK + tid;
    const double x = k < nterms ? terms[k] : 0.0;
    QFn f; f.a0 = f.a1 = 0;
    if (x <= 0.0 && x != neg_inf()) f = qfn_elem(x, e);        // special values: the chain kernel never applies F then
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        QFn o;
        o.a0 = __shfl_up_sync(0xffffffffu, f.a0, d);
        o.a1 = __shfl_up_sync(0xffffffffu, f.a1, d);
        if (lane >= d) f = qfn_compose(o, f);
    }
    if (lane == 31) wagg[w] = f;
    __syncthreads();
    if (tid == 0) {
        QFn t = wagg[0];
#pragma unroll
        for (int q = 1; q < SUM_BLK / 32; q++) t = qfn_compose(t, wagg[q]);
        bfn[b] = t; bexp[b] = e;
    }
}

// s (running sum, < 0, normal) in binade e and F built for e: apply if the result stays in the binade
__device__ __forceinline__ bool qfn_try(double &s, int e, const QFn f)
{
    const unsigned long long sb = (unsigned long long)__double_as_longlong(s);
    if ((sb >> 52) != (unsigned long long)(0x800 | (e + 1023))) return false;       // sign bit set + exponent field
    const unsigned long long Q0 = (sb & ((1ULL << 52) - 1)) | (1ULL << 52);
    const unsigned long long Q1 = Q0 + ((Q0 & 1ULL) == 0 ? f.a0 : f.a1);
    if (Q1 >= (1ULL << 53)) return false;
    s = __longlong_as_double((long long)((1ULL << 63) | ((unsigned long long)(e + 1023) << 52) | (Q1 & ((1ULL << 52) - 1))));
    return true;
}

// workspace of the block-structured sum, carved out of one device buffer
struct SumWs {
    double *bsum = nullptr, *bpre = nullptr; QFn *bfn = nullptr; int *bflag = nullptr, *bexp = nullptr;
    long long blk_stride = 0;                           // entries per list when the workspace holds a batch of lists
    template <class Buf> int bind(Buf &b, size_t nterms, size_t nlists = 1)
    {
        const size_t per = (nterms + SUM_BLK - 1) / SUM_BLK + 1, nblk = per * nlists;
        const int rc = b.ensure(nblk * (2 * sizeof(double) + sizeof(QFn) + 2 * sizeof(int)) + 64);
        if (rc) return rc;
        bfn = (QFn *)b.p; bsum = (double *)(bfn + nblk); bpre = bsum + nblk; bflag = (int *)(bpre + nblk); bexp = bflag + nblk;
        blk_stride = (long long)per;
        return 0;
    }
};

constexpr int SUMC_THREADS = 1024;
constexpr int SUMC_CACHE = 32;                         // open blocks whose terms are staged in shared memory
__host__ __device__ inline size_t sum_chain_smem_bytes(int nblk)
{
    return (size_t)((nblk < SUM_TILE_BLOCKS ? nblk : SUM_TILE_BLOCKS) + 32) * (sizeof(QFn) + 2 * sizeof(int));
}

// Everything the serial walk touches is staged in shared memory first (block functions, the terms of the blocks
// known to be open): the walk is a chain of dependent steps, a global-memory latency per step would dominate it.
// Lists longer than SUM_TILE_BLOCKS blocks go through in tiles, the running sum carried from tile to tile.
__global__ void __launch_bounds__(SUMC_THREADS) cp_sum_chain_kernel(const double *terms, int nterms, int nblk,
                                                                   const int *bflag, const int *bexp, const QFn *bfn,
                                                                   const UbSink ub_out_, unsigned int *reset_counter,
                                                                   const SumBatch batch)
{
    terms += (size_t)blockIdx.y * batch.term_stride;
    bflag += (size_t)blockIdx.y * batch.blk_stride; bexp += (size_t)blockIdx.y * batch.blk_stride; bfn += (size_t)blockIdx.y * batch.blk_stride;
    const UbSink ub_out = ub_sink_of(ub_out_);
    if (threadIdx.x == 0 && reset_counter) *reset_counter = 0u;
    constexpr int MAXG = SUM_TILE_BLOCKS / 32;
    extern __shared__ __align__(16) unsigned char sumc_raw[];
    __shared__ QFn gfn[MAXG]; __shared__ int gexp[MAXG]; __shared__ int mode_sh, ncache_sh;
    __shared__ int cache_blk[SUMC_CACHE];
    __shared__ double cache[SUMC_CACHE][SUM_BLK];
    __shared__ double buf[SUM_BLK];
    __shared__ double s_sh;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int cap = (min(nblk, SUM_TILE_BLOCKS) + 31) & ~31;
    QFn *sfn = reinterpret_cast<QFn *>(sumc_raw);
    int *sexp = reinterpret_cast<int *>(sfn + cap);        // binade or SUM_OPEN; SUM_OPEN - 1 - slot = open + cached
    int *stail = sexp + cap;                               // at the first block of a run: lane of the run's last block
    if (tid == 0) { mode_sh = 0; s_sh = 0.0; }
    __syncthreads();
    int flag = 0;
    for (int b = tid; b < nblk; b += SUMC_THREADS) flag |= bflag[b];
    if (flag) atomicOr(&mode_sh, flag);
    __syncthreads();
    const int mode = mode_sh;
    if (mode & 1) {                                            // reference loop, one thread
        if (tid == 0) { double ub = 0.0; for (int k = 0; k < nterms; k++) ub += terms[k]; ub_store(ub_out, ub); }
        return;
    }
    if (mode & 2) { if (tid == 0) ub_store(ub_out, neg_inf()); return; }

    for (int t0 = 0; t0 < nblk; t0 += SUM_TILE_BLOCKS) {
        const int nb = min(SUM_TILE_BLOCKS, nblk - t0), nb_pad = (nb + 31) & ~31;
        if (tid == 0) ncache_sh = 0;
        __syncthreads();
        for (int b = tid; b < nb_pad; b += SUMC_THREADS) {
            int e = SUM_OPEN; QFn f; f.a0 = f.a1 = 0;
            if (b < nb) {
                e = bexp[t0 + b];
                if (e != SUM_OPEN) f = bfn[t0 + b];
                else { const int slot = atomicAdd(&ncache_sh, 1); if (slot < SUMC_CACHE) { cache_blk[slot] = t0 + b; e = SUM_OPEN - 1 - slot; } }
            }
            sexp[b] = e; sfn[b] = f;
        }
        __syncthreads();
        const int ngrp = nb_pad / 32;
        // Runs: maximal stretches of consecutive blocks of a group with a function for the same binade.  A segmented
        // scan leaves in sfn[b] the composition from the run's first block up to b, so the walk takes a run in one
        // step; a group that is a single run is "clean" and is taken through gfn.
        for (int g = w; g < ngrp; g += SUMC_THREADS / 32) {
            const int b = g * 32 + lane;
            const int e = sexp[b];
            QFn f = sfn[b];
            const int e_prev = __shfl_up_sync(0xffffffffu, e, 1);
            const bool open = e <= SUM_OPEN;
            const bool head = lane == 0 || open || e_prev <= SUM_OPEN || e != e_prev;
            const unsigned int heads = __ballot_sync(0xffffffffu, head);
            const int my_head = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                QFn o;
                o.a0 = __shfl_up_sync(0xffffffffu, f.a0, d);
                o.a1 = __shfl_up_sync(0xffffffffu, f.a1, d);
                if (lane - d >= my_head) f = qfn_compose(o, f);
            }
            const unsigned int later = lane == 31 ? 0u : heads & (0xffffffffu << (lane + 1));
            sfn[b] = f;
            stail[b] = later ? __ffs(later) - 2 : 31;
            // clean = one run over the whole group (blocks past the end count as identity in the same binade)
            const int nreal = min(32, nb - g * 32);
            const bool one_run = (heads & (nreal >= 32 ? 0xffffffffu : ((1u << nreal) - 1u))) == 1u && !open;
            const int src = nreal - 1;
            const QFn fl = QFn{__shfl_sync(0xffffffffu, f.a0, src), __shfl_sync(0xffffffffu, f.a1, src)};
            const int e0 = __shfl_sync(0xffffffffu, e, 0);
            const bool clean = __shfl_sync(0xffffffffu, (int)one_run, 0) != 0;
            if (lane == 0) { gfn[g] = fl; gexp[g] = clean ? e0 : SUM_OPEN; }
        }
        {                                                      // terms of the open blocks, one warp per block
            const int nc = min(ncache_sh, SUMC_CACHE);
            if (w < nc) {
                const int base = cache_blk[w] * SUM_BLK, n = min(SUM_BLK, nterms - base);
#pragma unroll
                for (int q = 0; q < SUM_BLK / 32; q++) {
                    const int k = q * 32 + lane;
                    cache[w][k] = k < n ? terms[base + k] : 0.0;   // s + (+0.0) == s for every s the walk can hold
                }
            }
        }
        __syncthreads();
        if (w == 0) {
            double s = s_sh;                                   // identical in every lane of the warp
            for (int g = 0; g < ngrp; g++) {
                if (gexp[g] != SUM_OPEN && qfn_try(s, gexp[g], gfn[g])) continue;
                const int cnt = min(32, nb - g * 32);
                for (int i = 0; i < cnt;) {                    // mixed group: run by run
                    const int b = g * 32 + i;
                    const int e = sexp[b];
                    if (e > SUM_OPEN) {
                        const int j = min(stail[b], cnt - 1);
                        if (qfn_try(s, e, sfn[g * 32 + j])) { i = j + 1; continue; }
                        // (rare) the run as a whole does not apply -- a prediction was off: block by block
                        for (int ii = i; ii <= j; ii++) {
                            if (qfn_try(s, e, bfn[t0 + g * 32 + ii])) continue;
                            const int base = (t0 + g * 32 + ii) * SUM_BLK, n = min(SUM_BLK, nterms - base);
#pragma unroll
                            for (int q = 0; q < SUM_BLK / 32; q++) {
                                const int k = q * 32 + lane;
                                buf[k] = k < n ? terms[base + k] : 0.0;
                            }
                            __syncwarp();
#pragma unroll 16
                            for (int k = 0; k < SUM_BLK; k++) s = s + buf[k];
                            __syncwarp();
                        }
                        i = j + 1;
                        continue;
                    }
                    const double *src = buf;                   // open block: term by term
                    if (e < SUM_OPEN) {
                        src = cache[SUM_OPEN - 1 - e];
                    } else {                                   // more than SUMC_CACHE open blocks in the tile: not staged
                        const int base = (t0 + b) * SUM_BLK, n = min(SUM_BLK, nterms - base);
#pragma unroll
                        for (int q = 0; q < SUM_BLK / 32; q++) {
                            const int k = q * 32 + lane;
                            buf[k] = k < n ? terms[base + k] : 0.0;
                        }
                        __syncwarp();
                    }
#pragma unroll 16
                    for (int k = 0; k < SUM_BLK; k++) s = s + src[k];
                    __syncwarp();
                    i++;
                }
            }
            if (lane == 0) s_sh = s;
        }
        __syncthreads();
    }
    if (tid == 0) ub_store(ub_out, s_sh);
}

// obj = max(delta[N-1][.]) for the no-constraint case (cp.rs:139-141) and cur = argmax (cp.rs:86)
__global__ void cp_last_row_kernel(const CpParams p, double *obj_out, int *end_out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double *row = p.delta + (size_t)(p.N - 1) * p.K;
    double bv = row[0]; int bi = 0;
    for (int j = 1; j < p.K; j++) if (row[j] > bv) { bv = row[j]; bi = j; }
    *obj_out = bv; *end_out = bi;
}

constexpr int CP_BT_CHUNK = 256;   // rows per chunk of the backtrack's map composition (cp_dist.cuh)

__global__ void cp_set_choice_kernel(int32_t *choice, int comp, int value) { choice[comp] = value; }

// psi widened to u64 for the parity hook (cv_debug_cp_last_state)
__global__ void cp_widen_psi_kernel(const psi_t *psi, int64_t n, uint64_t *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = psi[i];
}

}  // namespace cvb
