// decode_chain.cuh -- plain Viterbi (mode R1) for SMALL batches of long sequences: one warp per sequence.
//
// Replaces viterbi::decode (reference src/viterbi_solver/viterbi.rs:5-32) when there are too few sequences to
// fill the lock-step tile kernel (e.g. datasets/ar: 15-25 day-sequences per model, T up to ~2000).  Lane i
// owns target state i (chain_warp.cuh); backpointers go to HBM as u8 rows and the same warp walks them back
// 32 rows at a time through a shared-memory stage.
#pragma once

#include "chain_warp.cuh"

namespace cvb {

struct DecodeChainParams {
    const double *A;         // [K][Kp]
    const double *BT;        // [M][Kp]
    const uint32_t *obs;     // [N]
    const int64_t *seq_off;  // [B+1]
    const uint32_t *order;   // [B] longest first
    uint8_t *psi;            // [N][Kp], row (off - seq_off[0] + t): the batch (or chunk) is contiguous from seq_off[0]
    uint32_t *path;          // [N]
    double *score;           // [B] or nullptr
    unsigned int *counter;
    int *status;
    int64_t M, B;
    int K, Kp;
    int bt_in_smem;          // logB^T fits shared memory (M * Kp * 8 bytes after logA)
    int obs16, path8;        // narrow host formats (cv_decode_batch_u16u8): u16 observations in, u8 states out
    // long-sequence split of the tile path (cv_api.cu): `order` is the list of long sequences, their number is read on
    // the device (min(*B_dev, B)); backpointer rows of list entry r start at row r * psi_stride; on the streamed host
    // path every finished sequence is counted in chunk_done[chunk of the sequence].  B_dev = nullptr: plain batch.
    const unsigned int *B_dev;
    int64_t psi_stride;
    unsigned int *chunk_done;
    int nch;
    int64_t cb[18];
};

constexpr int DC_WARPS = 4;
constexpr int CHAIN_PF = 4;          // emission rows are prefetched into L1 this many steps ahead
constexpr size_t CHAIN_BT_SMEM_MAX = 96 * 1024;

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// KQ > 0 (K <= 32, NSL = 1): logA column in registers, KQ = 2*ceil(K/8); KQ = 0: logA read from shared memory
template <int NSL, int KQ>
__global__ void __launch_bounds__(32 * DC_WARPS) decode_chain_kernel(const DecodeChainParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K, Kp = p.Kp;
    double *sA = reinterpret_cast<double *>(smem_raw);
    double *sBT = sA + (size_t)K * Kp;
    const size_t nbt = p.bt_in_smem ? (size_t)p.M * Kp : 0;
    uint8_t *stage_all = reinterpret_cast<uint8_t *>(sBT + nbt);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // per warp: delta rows [2][Kp] f64 | psi stage [32][Kp] u8 | path buffer [32] u32
    uint8_t *wbase = stage_all + (size_t)w * (16 * Kp + 32 * Kp + 32 * 4);
    double *sdw = reinterpret_cast<double *>(wbase);
    uint8_t *stage = wbase + 16 * Kp;
    uint32_t *pbuf = reinterpret_cast<uint32_t *>(stage + 32 * Kp);
    for (int e = threadIdx.x; e < K * Kp; e += blockDim.x) sA[e] = p.A[e];
    for (size_t e = threadIdx.x; e < nbt; e += blockDim.x) sBT[e] = p.BT[e];
    __syncthreads();
    const double *bt = p.bt_in_smem ? sBT : p.BT;
    const bool pf = !p.bt_in_smem;
    const int64_t psi_base = p.seq_off[0];      // offsets are absolute into obs / path; the psi buffer is chunk-local
    const int64_t nseq = p.B_dev ? min((int64_t)*p.B_dev, p.B) : p.B;

    const double zero_pi[NSL] = {};
    double acol[KQ > 0 ? 4 * KQ : 1];
    if (KQ > 0) {
#pragma unroll
        for (int j = 0; j < 4 * KQ; j++) acol[j] = (j < K && lane < Kp) ? sA[(size_t)j * Kp + lane] : neg_inf();
    }
    for (;;) {
        unsigned int r = 0;
        if (lane == 0) r = atomicAdd(p.counter, 1u);
        r = __shfl_sync(0xffffffffu, r, 0);
        if ((int64_t)r >= nseq) break;
        const uint32_t b = p.order[r];
        const int64_t off = p.seq_off[b];
        const int len = (int)(p.seq_off[b + 1] - off);
        const int64_t prow = p.B_dev ? (int64_t)r * p.psi_stride : off - psi_base;    // first backpointer row of this sequence

        double d[NSL], e_next[NSL];
        int col[NSL];
#pragma unroll
        for (int s = 0; s < NSL; s++) {                                                      // viterbi.rs:6
            d[s] = 0.0; col[s] = min(lane + 32 * s, Kp - 1);
            if (lane + 32 * s < Kp) { sdw[lane + 32 * s] = (lane + 32 * s < K) ? 0.0 : neg_inf(); sdw[Kp + lane + 32 * s] = neg_inf(); }
        }
        __syncwarp();
        auto load_obs = [&](int t) -> uint32_t {
            if (t >= len) return 0u;
            return p.obs16 ? (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(p.obs) + off + t) : __ldg(p.obs + off + t);
        };
        auto store_path = [&](int64_t idx, uint32_t st) {
            if (p.path8) reinterpret_cast<uint8_t *>(p.path)[idx] = (uint8_t)st; else p.path[idx] = st;
        };
        auto checked = [&](uint32_t o) -> uint32_t {                                             // index panic
            if ((int64_t)o >= p.M) { *p.status = 3; return 0u; }
            return o;
        };
        // observation ring: oq[k] = obs of step t + 1 + k; emission rows go to L1 CHAIN_PF steps ahead
        uint32_t oq[CHAIN_PF + 1];
#pragma unroll
        for (int k = 0; k <= CHAIN_PF; k++) oq[k] = load_obs(2 + k);
        const uint32_t o1 = checked(load_obs(1));
#pragma unroll
        for (int s = 0; s < NSL; s++) e_next[s] = bt[(size_t)o1 * Kp + col[s]];
        if (pf) {
#pragma unroll
            for (int k = 0; k < CHAIN_PF; k++) prefetch_l1(p.BT + (size_t)checked(oq[k]) * Kp + col[0]);
        }

        for (int t = 1; t < len; t++) {
            double e[NSL];
            const uint32_t on = checked(oq[0]);
#pragma unroll
            for (int s = 0; s < NSL; s++) { e[s] = e_next[s]; e_next[s] = bt[(size_t)on * Kp + col[s]]; }
            if (pf) prefetch_l1(p.BT + (size_t)checked(oq[CHAIN_PF]) * Kp + col[0]);
#pragma unroll
            for (int k = 0; k < CHAIN_PF; k++) oq[k] = oq[k + 1];
            oq[CHAIN_PF] = load_obs(t + 2 + CHAIN_PF);
            double best[NSL]; int idx[NSL];
            if (KQ > 0) chain_scan_reg<(KQ > 0 ? KQ : 2)>(sdw + ((t - 1) & 1) * Kp, reinterpret_cast<const double (&)[4 * (KQ > 0 ? KQ : 2)]>(acol), best[0], idx[0]);
            else chain_scan<NSL>(sdw + ((t - 1) & 1) * Kp, sA, Kp, K, lane, false, zero_pi, best, idx);   // viterbi.rs:15-16
#pragma unroll
            for (int s = 0; s < NSL; s++) {
                double v = best[s] + e[s];                                                         // viterbi.rs:17
                int ix = idx[s];
                if (!(e[s] > neg_inf())) { v = neg_inf(); ix = 0; }                              // viterbi.rs:19-21
                const int i = lane + 32 * s;
                if (i < K) p.psi[(size_t)(prow + t) * Kp + i] = (uint8_t)ix;                      // viterbi.rs:18
                d[s] = (i < K) ? v : neg_inf();
                if (i < K) sdw[(t & 1) * Kp + i] = v;
            }
            __syncwarp();
        }
        // end state (viterbi.rs:24)
        double bv = neg_inf(); int cur = 0x7fffffff;
#pragma unroll
        for (int s = 0; s < NSL; s++) {
            const int i = lane + 32 * s;
            if (i < K && (cur == 0x7fffffff || d[s] > bv)) { bv = d[s]; cur = i; }
        }
        warp_argmax(bv, cur);
        if (lane == 0) {
            if (p.score) p.score[b] = bv;
            store_path(off + len - 1, (uint32_t)cur);
        }
        // backtrace (viterbi.rs:27-30): 32 backpointer rows at a time through shared memory
        __syncwarp();
        for (int thi = len - 1; thi >= 1; thi -= 32) {
            const int nrows = min(32, thi);                       // rows thi, thi-1, ..., thi-nrows+1
            if (lane < nrows) {
                const uint2 *src = reinterpret_cast<const uint2 *>(p.psi + (size_t)(prow + thi - lane) * Kp);
                uint2 *dst = reinterpret_cast<uint2 *>(stage + (size_t)lane * Kp);
                for (int k = 0; k < Kp / 8; k++) dst[k] = __ldcg(src + k);
            }
            __syncwarp();
            if (lane == 0) {
                for (int k = 0; k < nrows; k++) { cur = stage[(size_t)k * Kp + cur]; pbuf[k] = (uint32_t)cur; }
            }
            __syncwarp();
            cur = __shfl_sync(0xffffffffu, cur, 0);
            if (lane < nrows) store_path(off + thi - lane - 1, pbuf[lane]);
            __syncwarp();
        }
        if (p.chunk_done) {                                          // streamed host path: this sequence's results may leave
            __threadfence_system();
            __syncwarp();
            if (lane == 0) {
                int c = 0;
                while (c + 1 < p.nch && (int64_t)b >= p.cb[c + 1]) c++;
                atomicAdd(p.chunk_done + c, 1u);
            }
        }
    }
}

}  // namespace cvb
