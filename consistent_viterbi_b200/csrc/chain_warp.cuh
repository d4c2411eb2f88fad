// chain_warp.cuh -- latency-oriented max-plus step: ONE WARP PER CHAIN, lanes = target states.
//
// The tile kernels (decode_small / decode_large) need thousands of independent sequences in flight.  Few,
// long chains -- datasets/ar (60 day-sequences), the clamp-to-clamp segments of one B&B node of CPSolver, the
// init_viterbi prefix -- are serial in t, so what matters is the latency of one step.  Here lane i (and i+32
// when K > 32) owns target state i: delta_old sits in a per-warp shared-memory row read with broadcast 128-bit loads, logA is read
// from shared memory (consecutive lanes -> consecutive words), and the K predecessors are scanned as four
// contiguous index ranges with independent running maxima, merged in ascending order with a strict > so the
// result is the reference's first-argmax (ndarray-stats argmax; viterbi.rs:16, cp.rs:53).
#pragma once

#include "common.cuh"

namespace cvb {

// NSL = states per lane (i = lane + 32*s): 1 for K <= 32, 2 for K <= 64

__device__ __forceinline__ double shfl_f64(double v, int src)
{
    return __shfl_sync(0xffffffffu, v, src);
}

// first-argmax over predecessors j of fl(d_old[j] + tr_j(i)) for this lane's states.
//   tr_j(i) = sA[j*Kp + i], or the constant pi_i[s] when use_pi (MetaElements::transitions at el.t == 0).
// sd = this warp's delta_old row in shared memory (Kp doubles, entries >= K are -inf), read as broadcast
// 128-bit loads (two predecessors per load).  Returns best value / index per owned state.
template <int CW_NSL>
__device__ __forceinline__ void chain_scan(const double *__restrict__ sd, const double *__restrict__ sA, int Kp, int K,
                                           int lane, bool use_pi, const double (&pi_i)[CW_NSL],
                                           double (&best)[CW_NSL], int (&idx)[CW_NSL])
{
    // four contiguous ranges [r*Kq, (r+1)*Kq), Kq even so that a range starts on a 16-byte boundary
    const int Kq = (((K + 3) >> 2) + 1) & ~1;
    double b[4][CW_NSL]; int ix[4][CW_NSL];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int s = 0; s < CW_NSL; s++) { b[r][s] = neg_inf(); ix[r][s] = r * Kq; }
    int col[CW_NSL];
#pragma unroll
    for (int s = 0; s < CW_NSL; s++) col[s] = min(lane + 32 * s, Kp - 1);   // lanes beyond Kp read a valid word
    // Branch-free body: predecessors beyond K contribute -inf (never strictly greater), so the loads of the
    // next predecessors overlap the compare-select chains.
#pragma unroll 2
    for (int jj = 0; jj < Kq; jj += 2) {
        double2 dj[4]; double tr[4][2][CW_NSL];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int j = r * Kq + jj;                                         // even
            const int jc = min(j, Kp - 2);
            dj[r] = *reinterpret_cast<const double2 *>(sd + jc);              // delta_old[j], delta_old[j+1]
            if (j >= K) dj[r] = make_double2(neg_inf(), neg_inf());
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int s = 0; s < CW_NSL; s++) {
                    const int ju = min(j + u, K - 1);
                    const double a = use_pi ? pi_i[s] : sA[(size_t)ju * Kp + col[s]];
                    tr[r][u][s] = (j + u < K) ? a : neg_inf();
                }
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int j = r * Kq + jj;
#pragma unroll
            for (int s = 0; s < CW_NSL; s++) {
                const double v0 = dj[r].x + tr[r][0][s];
                if (v0 > b[r][s]) { b[r][s] = v0; ix[r][s] = j; }
                const double v1 = dj[r].y + tr[r][1][s];
                if (v1 > b[r][s]) { b[r][s] = v1; ix[r][s] = j + 1; }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < CW_NSL; s++) {
        // range 0 always contains j = 0; an all -inf range keeps its first index, which can only win the merge
        // below if an earlier range were smaller -- impossible for -inf -- so the result index is the first max
        best[s] = b[0][s]; idx[s] = ix[0][s];
#pragma unroll
        for (int r = 1; r < 4; r++)
            if (b[r][s] > best[s]) { best[s] = b[r][s]; idx[s] = ix[r][s]; }
    }
}


// Value-only variant of chain_scan (the CFN sweeps need max_j fl(d_old[j] + tr_j(i)) but no backpointer):
// DADD + DSETP + 2 selects per predecessor instead of 3.  Ranges ascending, strict >: the first maximum's bits.
template <int CW_NSL>
__device__ __forceinline__ void chain_scan_val(const double *__restrict__ sd, const double *__restrict__ sA, int Kp, int K,
                                               int lane, bool use_pi, const double (&pi_i)[CW_NSL], double (&best)[CW_NSL])
{
    const int Kq = (((K + 3) >> 2) + 1) & ~1;
    double b[4][CW_NSL];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int s = 0; s < CW_NSL; s++) b[r][s] = neg_inf();
    int col[CW_NSL];
#pragma unroll
    for (int s = 0; s < CW_NSL; s++) col[s] = min(lane + 32 * s, Kp - 1);
#pragma unroll 2
    for (int jj = 0; jj < Kq; jj += 2) {
        double2 dj[4]; double tr[4][2][CW_NSL];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int j = r * Kq + jj;
            const int jc = min(j, Kp - 2);
            dj[r] = *reinterpret_cast<const double2 *>(sd + jc);
            if (j >= K) dj[r] = make_double2(neg_inf(), neg_inf());
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int s = 0; s < CW_NSL; s++) {
                    const int ju = min(j + u, K - 1);
                    const double a = use_pi ? pi_i[s] : sA[(size_t)ju * Kp + col[s]];
                    tr[r][u][s] = (j + u < K) ? a : neg_inf();
                }
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int s = 0; s < CW_NSL; s++) {
                const double v0 = dj[r].x + tr[r][0][s];
                b[r][s] = v0 > b[r][s] ? v0 : b[r][s];
                const double v1 = dj[r].y + tr[r][1][s];
                b[r][s] = v1 > b[r][s] ? v1 : b[r][s];
            }
    }
#pragma unroll
    for (int s = 0; s < CW_NSL; s++) {
        best[s] = b[0][s];
#pragma unroll
        for (int r = 1; r < 4; r++) best[s] = b[r][s] > best[s] ? b[r][s] : best[s];
    }
}

// Register-resident variant for K <= 32 (one state per lane): the lane's logA column acol[j] = logA[j][lane]
// (-inf for j >= K) stays in registers for the whole kernel, so a predecessor costs half a broadcast LDS.128
// plus DADD / DSETP / 3 selects.  KQ = 2*ceil(K/8) predecessors per range, fully unrolled; sd[j] = -inf, j >= K.
template <int KQ>
__device__ __forceinline__ void chain_scan_reg(const double *__restrict__ sd, const double (&acol)[4 * KQ],
                                               double &best, int &idx)
{
    double b[4]; int ix[4];
#pragma unroll
    for (int r = 0; r < 4; r++) { b[r] = neg_inf(); ix[r] = r * KQ; }
#pragma unroll
    for (int jj = 0; jj < KQ; jj += 2) {
        double2 dj[4];
#pragma unroll
        for (int r = 0; r < 4; r++) dj[r] = *reinterpret_cast<const double2 *>(sd + r * KQ + jj);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int j = r * KQ + jj;
            const double v0 = dj[r].x + acol[j];
            if (v0 > b[r]) { b[r] = v0; ix[r] = j; }
            const double v1 = dj[r].y + acol[j + 1];
            if (v1 > b[r]) { b[r] = v1; ix[r] = j + 1; }
        }
    }
    best = b[0]; idx = ix[0];
#pragma unroll
    for (int r = 1; r < 4; r++)
        if (b[r] > best) { best = b[r]; idx = ix[r]; }
}

// (value, index) argmax across the warp's owned states: strictly greater value, else lower index.
__device__ __forceinline__ void warp_argmax(double &v, int &ix)
{
#pragma unroll
    for (int dlt = 16; dlt >= 1; dlt >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, dlt);
        const int oi = __shfl_xor_sync(0xffffffffu, ix, dlt);
        if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
    }
}

}  // namespace cvb
