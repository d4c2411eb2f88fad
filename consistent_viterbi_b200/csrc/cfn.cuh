// cfn.cuh -- cost-table compilation of the CFN formulation (reference src/viterbi_solver/cfn.rs:11-167).
//
// write_cfn runs longest_path (cfn.rs:11-35) K*K times per consecutive pair of component boundaries: a clamped
// Viterbi sweep from (t_from, n_from) to (t_to, n_to).  Rows before t_to do not depend on n_to, and at t_to the
// only finite entry of the last row is n_to's, so the K*K runs of a pair are K sweeps -- one per n_from -- whose
// last step is evaluated for every n_to at once.  That is the same warp-per-chain max-plus step as the
// constrained decode (chain_warp.cuh), in this file's arithmetic order: max_j fl(row[j] + tr_j), then + emit
// ((max(delta + a)) + b, cfn.rs:18-20; the first maximum is kept like ndarray-stats' max does).  The unary
// start/end costs (cfn.rs:37-80) are chains of the same kind.  Accumulating the costs into the k*k tables has
// an order (boundary pairs ascending) and the reference's `== 0.0 => assign, else +=` rule: one thread per
// (n1, n2) walks the pairs in order.
#pragma once
#include "chain_warp.cuh"
#include "cp_kernels.cuh"

namespace cvb {

struct CfnChain {
    int64_t t0, t1;       // rows t0+1 .. t1 are swept; row t0 is the start row
    int32_t clamp;        // state the constrained rows are clamped to (pair: n_from, end cost: n); -1 = never clamp
    int32_t kind;         // 0 pair (last row unclamped, K values out), 1 unary start (row 0 = init_probs, K values out),
                          // 2 unary end (max of the last row, one value out)
    int64_t out;          // index into `out` (doubles)
};

constexpr int CFN_WARPS = 4;

// one warp per chain; NSL = states per lane (K <= 32: 1, K <= 64: 2)
template <int NSL>
__global__ void __launch_bounds__(32 * CFN_WARPS) cfn_chain_kernel(const CpParams p, const CfnChain *chains, int64_t nchains, double *out)
{
    extern __shared__ __align__(16) unsigned char cfn_raw[];
    const int K = p.K, Kp = p.Kp;
    double *sA = reinterpret_cast<double *>(cfn_raw);
    for (int e = threadIdx.x; e < K * Kp; e += blockDim.x) sA[e] = p.A[e];
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double *sd = sA + (size_t)K * Kp + (size_t)w * 2 * Kp;               // this warp's rows [2][Kp]
    double pi_i[NSL]; int col[NSL];
#pragma unroll
    for (int s = 0; s < NSL; s++) { col[s] = min(lane + 32 * s, Kp - 1); pi_i[s] = p.Pi[col[s]]; }

    for (int64_t c = (int64_t)blockIdx.x * CFN_WARPS + w; c < nchains; c += (int64_t)gridDim.x * CFN_WARPS) {
        const CfnChain ch = chains[c];
        __syncwarp();
#pragma unroll
        for (int s = 0; s < NSL; s++) {
            const int i = lane + 32 * s;
            double d = neg_inf();
            if (i < K) {
                if (ch.kind == 1) d = p.Pi[i] + p.BT[(size_t)p.obs[ch.t0] * Kp + i];     // init_probs (hmm.rs:215-218)
                else if (i == ch.clamp) d = 0.0;                                         // cfn.rs:12-13 / 62-63
            }
            if (i < Kp) { sd[i] = d; sd[Kp + i] = neg_inf(); }
        }
        __syncwarp();
        int cur = 0;
        for (int64_t t = ch.t0 + 1; t <= ch.t1; t++) {
            const uint32_t o = __ldg(p.obs + t);
            const bool st = __ldg(p.start + t) != 0;
            const bool clamped = ch.clamp >= 0 && __ldg(p.comp + t) >= 0 && !(ch.kind == 0 && t == ch.t1);
            double best[NSL];
            chain_scan_val<NSL>(sd + cur * Kp, sA, Kp, K, lane, st, pi_i, best);    // max_j fl(row[j] + tr_j), first max
#pragma unroll
            for (int s = 0; s < NSL; s++) {
                const int i = lane + 32 * s;
                if (i < K) {
                    double v = best[s] + __ldg(p.BT + (size_t)o * Kp + i);          // + emit_prob (cfn.rs:20,27)
                    if (clamped && i != ch.clamp) v = neg_inf();                   // row.fill(-inf), one entry set
                    sd[(cur ^ 1) * Kp + i] = v;
                }
            }
            cur ^= 1;
            __syncwarp();
        }
        if (ch.kind == 2) {                                                        // *row.max().unwrap(): first maximum
            double m = neg_inf(); int mi = 0x7fffffff;
#pragma unroll
            for (int s = 0; s < NSL; s++) {
                const int i = lane + 32 * s;
                if (i < K) { const double v = sd[cur * Kp + i]; if (mi == 0x7fffffff || v > m) { m = v; mi = i; } }
            }
            warp_argmax(m, mi);                                                    // strictly greater, else lower index
            if (lane == 0) out[ch.out] = m;
        } else {
#pragma unroll
            for (int s = 0; s < NSL; s++) {
                const int i = lane + 32 * s;
                if (i < K) out[ch.out + i] = sd[cur * Kp + i];
            }
        }
    }
}

// cfn.rs:118-137: tables[cf][ct][n1][n2] and tables[ct][cf][n2][n1] take the pair's cost when it is not -inf:
// assigned if the entry is still 0.0, added otherwise.  costs[i][n1][n2]; pair i = boundaries (i, i+1).
// The two entries always receive the same updates, and an entry is fed by pairs (cf, ct) through cost[n1][n2] and
// by pairs (ct, cf) through cost[n2][n1].  One CTA per unordered component pair c1 < c2 (the host lists its
// boundary pairs in ascending order, bit 0 = the pair runs c2 -> c1); thread (a, b) owns
// {tables[c1][c2][a][b], tables[c2][c1][b][a]} and adds the contributions in list order -- the reference's order
// per entry.  The costs of 16 pairs are fetched before the dependent add chain touches them.
constexpr int CFN_ACC_U = 16;
__global__ void __launch_bounds__(1024) cfn_accumulate_kernel(const double *costs, const int64_t *plist, const int64_t *toff,
                                                             const int32_t *tc1, const int32_t *tc2, int K, int k, double *tables)
{
    const int tab = blockIdx.x, c1 = tc1[tab], c2 = tc2[tab];
    const int64_t p0 = toff[tab], p1 = toff[tab + 1];
    for (int e = threadIdx.x; e < K * K; e += blockDim.x) {
        const int a = e / K, b = e % K;
        const size_t fwd = (size_t)a * K + b, rev = (size_t)b * K + a;
        double acc = 0.0;                                       // Array2::from_elem(.., 0.0), cfn.rs:116
        for (int64_t q = p0; q < p1; q += CFN_ACC_U) {
            double c[CFN_ACC_U];
#pragma unroll
            for (int u = 0; u < CFN_ACC_U; u++) {
                c[u] = neg_inf();
                if (q + u < p1) {
                    const int64_t pl = plist[q + u];
                    c[u] = costs[(size_t)(pl >> 1) * K * K + ((pl & 1) ? rev : fwd)];
                }
            }
#pragma unroll
            for (int u = 0; u < CFN_ACC_U; u++)
                if (c[u] != neg_inf()) acc = (acc == 0.0) ? c[u] : acc + c[u];
        }
        tables[(((size_t)c1 * k + c2) * K + a) * K + b] = acc;
        tables[(((size_t)c2 * k + c1) * K + b) * K + a] = acc;
    }
}

}  // namespace cvb
