// decode_small.cuh -- batched plain Viterbi (mode R1), K <= 64.
//
// Replaces B calls of viterbi::decode (reference src/viterbi_solver/viterbi.rs:5-32).
//
// Mapping (B200): one persistent CTA per SM pulls tiles of NS = 64*S sequences
// (sequences are pre-sorted by length, longest first, so a tile runs in lock
// step with little idle tail).  logA (K x Kp f64) is staged once per CTA into
// shared memory with a TMA bulk copy; delta lives in shared memory, double
// buffered, laid out [state j][sequence slot] so that
//   - a lane owns TP=2 adjacent sequences  -> one conflict-free LDS.128 per j
//   - a warp owns TQ=8 adjacent target states -> four broadcast LDS.128 per j
// giving 16 max-plus cells per thread per predecessor j from 5 shared loads
// (shared-memory crossbar at ~50 % when the FP64 pipe is saturated).
// Warps = G state groups (G = ceil(K/8)) x S sequence groups.
// Per step a thread adds the emission (viterbi.rs:17: (delta + a) + b), forces
// psi = 0 / delta = -inf where the emission is -inf (viterbi.rs:12,19-21), writes
// 8 backpointers as one 8-byte store and 16 delta values to the other buffer.
// The backtrace (viterbi.rs:24-30) runs fused at the end of the tile, one thread
// per sequence, while the tile's psi rows are still in L2.
#pragma once

#include "common.cuh"

namespace cvb {

struct DecodeSmallParams {
    const double *A;         // [K][Kp]   logA, columns >= K padded with -inf
    const double *BT;        // [M][Kp]   logB transposed (obs-major), padded -inf
    const uint32_t *obs;     // [N]
    const int64_t *seq_off;  // [B+1]
    const uint32_t *order;   // [B] sequence ids, longest first
    uint8_t *psi;            // [N][Kp]   backpointers
    uint32_t *path;          // [N]
    double *score;           // [B] or nullptr
    unsigned int *tile_counter;
    int *status;             // 0 ok, else CV_ERR_*
    int64_t M, B;
    int K, Kp, G, S, ntiles, zero;
};

// dynamic smem: [ A: K*Kp f64 ][ delta: 2*K*NS f64 ][ off: NS i64 ][ len: NS i32 ][ seq: NS u32 ][ mbar ][ tile ]
__host__ __device__ inline size_t decode_small_smem_bytes(int K, int Kp, int NS)
{
    return (size_t)K * Kp * 8 + (size_t)2 * K * NS * 8 + (size_t)NS * (8 + 4 + 4) + 16 + 16;
}

template <int VARIANT>
__global__ void __launch_bounds__(512, 1) decode_small_kernel(const DecodeSmallParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int K = p.K, Kp = p.Kp, NS = 64 * p.S;
    double *sA = reinterpret_cast<double *>(smem_raw);
    double *sD = sA + (size_t)K * Kp;
    int64_t *sOff = reinterpret_cast<int64_t *>(sD + (size_t)2 * K * NS);
    int *sLen = reinterpret_cast<int *>(sOff + NS);
    uint32_t *sSeq = reinterpret_cast<uint32_t *>(sLen + NS);
    uint64_t *sBar = reinterpret_cast<uint64_t *>(sSeq + NS);
    int *sTile = reinterpret_cast<int *>(sBar + 2);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = w % p.G, sg = w / p.G;
    const int i0 = g * TQ;
    const int s0 = sg * SEQ_PER_WARP + lane * TP;

    // ---- stage logA once per CTA (TMA bulk copy, UBLKCP) ----
    if (tid == 0) {
        mbar_init(sBar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t bytes = (uint32_t)((size_t)K * Kp * 8);
        mbar_expect_tx(sBar, bytes);
        tma_bulk_g2s(sA, p.A, bytes, sBar);
    }
    __syncthreads();
    mbar_wait(sBar, 0);

    for (;;) {
        if (tid == 0) *sTile = (int)atomicAdd(p.tile_counter, 1u);
        __syncthreads();
        const int tile = *sTile;
        if (tile >= p.ntiles) break;

        // ---- tile set-up: sequence slots, delta(0) = 0.0 (viterbi.rs:6) ----
        for (int s = tid; s < NS; s += blockDim.x) {
            const int64_t r = (int64_t)tile * NS + s;
            int64_t off = 0; int len = 0; uint32_t b = 0;
            if (r < p.B) {
                b = p.order[r];
                off = p.seq_off[b];
                len = (int)(p.seq_off[b + 1] - off);
            }
            sOff[s] = off; sLen[s] = len; sSeq[s] = b;
        }
        for (int e = tid; e < K * NS; e += blockDim.x) sD[e] = 0.0;
        __syncthreads();

        const int Tmax = sLen[0];   // slot 0 holds the longest sequence of the tile
        int len_p[TP]; int64_t off_p[TP];
#pragma unroll
        for (int q = 0; q < TP; q++) { len_p[q] = sLen[s0 + q]; off_p[q] = sOff[s0 + q]; }
        const bool warp_has_work = __any_sync(0xffffffffu, len_p[0] > 1 || len_p[1] > 1);

        uint32_t o_next[TP];
#pragma unroll
        for (int q = 0; q < TP; q++) o_next[q] = (1 < len_p[q]) ? __ldg(p.obs + off_p[q] + 1) : 0u;

        for (int t = 1; t < Tmax; t++) {
            const bool act0 = t < len_p[0], act1 = t < len_p[1];
            const bool act[TP] = {act0, act1};
            if (warp_has_work && __any_sync(0xffffffffu, act0 || act1)) {
                // emission rows for this step (issued before the j loop, used after it)
                double em[TP][TQ];
#pragma unroll
                for (int q = 0; q < TP; q++) {
                    uint32_t o = o_next[q];
                    if (act[q] && (int64_t)o >= p.M) { *p.status = 3; o = 0; }   // index panic in the reference
                    const double2 *src = reinterpret_cast<const double2 *>(p.BT + (size_t)o * Kp + i0);
#pragma unroll
                    for (int k = 0; k < TQ / 2; k++) {
                        double2 v = act[q] ? __ldg(src + k) : make_double2(0.0, 0.0);
                        em[q][2 * k] = v.x; em[q][2 * k + 1] = v.y;
                    }
                    o_next[q] = (t + 1 < len_p[q]) ? __ldg(p.obs + off_p[q] + t + 1) : 0u;
                }

                double best[TP][TQ]; int idx[TP][TQ];
                const double *dcur = sD + (size_t)((t - 1) & 1) * K * NS + s0;
                maxplus_tile<VARIANT>(dcur, NS, sA + i0, Kp, K, best, idx, p.zero);

                double *dnext = sD + (size_t)(t & 1) * K * NS + s0;
#pragma unroll
                for (int q = 0; q < TP; q++) {
                    if (!act[q]) continue;
                    uint32_t pk[2] = {0u, 0u};
#pragma unroll
                    for (int k = 0; k < TQ; k++) {
                        double v = best[q][k] + em[q][k];            // (delta + a) + b   viterbi.rs:17
                        int ix = idx[q][k];
                        if (!(em[q][k] > neg_inf())) { v = neg_inf(); ix = 0; }   // viterbi.rs:19-21
                        if (i0 + k < K) dnext[(size_t)(i0 + k) * NS + q] = v;
                        pk[k >> 2] |= (uint32_t)ix << (8 * (k & 3));
                    }
                    *reinterpret_cast<uint2 *>(p.psi + (size_t)(off_p[q] + t) * Kp + i0) = make_uint2(pk[0], pk[1]);
                }
            }
            __syncthreads();
        }

        // ---- end state (viterbi.rs:24) + backtrace (viterbi.rs:25-30), one thread per sequence ----
        if (tid < NS) {
            const int len = sLen[tid];
            if (len > 0) {
                const int64_t off = sOff[tid];
                const double *fin = sD + (size_t)((len - 1) & 1) * K * NS + tid;
                double bv = fin[0]; uint32_t cur = 0;
                for (int i = 1; i < K; i++) {
                    const double v = fin[(size_t)i * NS];
                    if (v > bv) { bv = v; cur = (uint32_t)i; }
                }
                if (p.score) p.score[sSeq[tid]] = bv;
                p.path[off + len - 1] = cur;
                for (int t = len - 1; t >= 1; t--) {
                    cur = __ldcg(p.psi + (size_t)(off + t) * Kp + cur);
                    p.path[off + t - 1] = cur;
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace cvb
