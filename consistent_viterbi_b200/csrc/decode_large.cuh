// decode_large.cuh -- batched plain Viterbi (mode R1) for K > 64: logA tiled
// through shared memory (reference: viterbi::decode, src/viterbi_solver/viterbi.rs:5-32).
//
// One time step of 64 sequences is a max-plus "GEMM"
//     delta_t[b][i] = max_j ( delta_{t-1}[b][j] + logA[j][i] )   (+ logB[i][o_t(b)])
// that no longer fits one SM's shared memory (logA alone is 8 MB at K = 1024), so
// the step is cut into work items (t, rb, cb): row block rb = 64 sequences (sorted
// by length), column block cb = 128 target states.  A persistent grid (one CTA
// per SM) claims items from a global counter in (t, rb, cb) order; item (t, rb, *)
// may start once all NCB items of (t-1, rb, *) have published their delta slice
// (per-row-block completion counter, release/acquire).  Items are claimed only by
// running CTAs and depend only on lower-numbered items, so the scheme cannot
// deadlock and keeps all 148 SMs busy without clusters or grid-wide barriers.
//
// Inside an item, a producer warp streams 16-row chunks of logA[:, cb] (16 KB) and
// of delta_{t-1}[rb] (8 KB) from L2 into a 6-stage shared-memory ring with TMA
// bulk copies (cp.async.bulk + mbarrier full/empty pairs); 16 warps each
// own 8 target states x 64 sequences and run the same value-only 2x8 register
// micro-tile as the small-K kernel.  Every delta row is kept: the history is
// [t][Kl][Bpad] f64 (state-major, sequence fastest), so the chunk loads of step t-1
// and the slice stores of step t are contiguous, and the backtrace recomputes the
// backpointer of the one path state per step from the stored row (lazy psi, see
// decode_small.cuh).  When the history of the whole batch does not fit in HBM the
// host runs the batch in groups of row blocks.
#pragma once

#include "common.cuh"

namespace cvb {

constexpr int LG_BM = 64;        // sequences per row block
constexpr int LG_BK = 16;        // predecessor rows per pipeline stage
constexpr int LG_STAGES = 6;
constexpr int LG_CONSUMER_WARPS = LARGE_BN / TQ;               // 16
constexpr int LG_THREADS = 32 * (LG_CONSUMER_WARPS + 1);       // + one producer warp (warp 16)
constexpr int LG_STAGE_A = LG_BK * LARGE_BN;                   // doubles
constexpr int LG_STAGE_D = LG_BK * LG_BM;                      // doubles
constexpr size_t LG_SMEM_BYTES = (size_t)LG_STAGES * (LG_STAGE_A + LG_STAGE_D) * 8 + 256;

struct DecodeLargeParams {
    const double *A;          // [Kl][Kl] logA padded with -inf
    const double *BT;         // [M][Kl]
    const uint32_t *obs;      // [N]
    const int64_t *seq_off;   // [B+1]
    const uint32_t *order;    // [B] longest first
    const uint32_t *sorted_len;  // [B] lengths in that order
    uint32_t *path;           // [N]
    double *score;            // [B] or nullptr
    double *hist;             // [Tmax][Kl][Bpad] delta history of this group of row blocks
    const double *AT;         // [Kl][Kl] logA transposed (backtrace)
    const long long *step_start; // [Tmax+1] first item id of step t (index t, t >= 1); [Tmax] = total
    unsigned long long *item_counter;
    unsigned int *done;       // [NRB] completed items per row block
    int *status;
    int64_t M, B, Bpad;
    int64_t rank0;            // sorted rank of the group's first sequence
    int K, Kl, NCB, NRB, Tmax;
};

// NJ > 0: the chunk's predecessor count at compile time (full chunks: the loop is unrolled completely, no loop
// overhead and no loop-carried register moves inside a chunk); NJ = 0: run-time count (the last, partial chunk)
template <int NJ = 0>
__device__ __forceinline__ void maxplus_accum_val(const double *__restrict__ dcol, int ldd,
                                                  const double *__restrict__ arow, int lda, int nj,
                                                  double (&best)[TP][TQ])
{
#pragma unroll (NJ > 0 ? NJ : 2)
    for (int jj = 0; jj < (NJ > 0 ? NJ : nj); jj++) {
        const double2 d = *reinterpret_cast<const double2 *>(dcol + jj * ldd);
        const double2 a01 = *reinterpret_cast<const double2 *>(arow + jj * lda);
        const double2 a23 = *reinterpret_cast<const double2 *>(arow + jj * lda + 2);
        const double2 a45 = *reinterpret_cast<const double2 *>(arow + jj * lda + 4);
        const double2 a67 = *reinterpret_cast<const double2 *>(arow + jj * lda + 6);
        const double a[TQ] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y, a67.x, a67.y};
        const double dd[TP] = {d.x, d.y};
#pragma unroll
        for (int p = 0; p < TP; p++)
#pragma unroll
            for (int q = 0; q < TQ; q++) {
                const double v = dd[p] + a[q];
                best[p][q] = v > best[p][q] ? v : best[p][q];
            }
    }
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __maxnreg__(96) decode_large_kernel(const DecodeLargeParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sA = reinterpret_cast<double *>(smem_raw);                       // [STAGES][BK][128]
    double *sD = sA + (size_t)LG_STAGES * LG_STAGE_A;                        // [STAGES][BK][64]
    uint64_t *full = reinterpret_cast<uint64_t *>(sD + (size_t)LG_STAGES * LG_STAGE_D);
    uint64_t *empty = full + LG_STAGES;
    long long *sItem = reinterpret_cast<long long *>(empty + LG_STAGES);    // [0]=item, [1]=t, [2]=rb, [3]=cb

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int Kl = p.Kl, K = p.K;
    const int nchunks = (K + LG_BK - 1) / LG_BK;

    if (tid == 0) {
        for (int s = 0; s < LG_STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, LG_CONSUMER_WARPS); }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const long long total_items = p.step_start[p.Tmax];
    uint32_t chunk_ctr = 0;   // chunks pushed (producer) / consumed (consumers) so far; same sequence in every warp

    for (;;) {
        // ---- claim the next work item ----
        if (tid == 0) {
            const long long n = (long long)atomicAdd(p.item_counter, 1ULL);
            long long t = 0, rb = 0, cb = 0;
            if (n < total_items) {
                int lo = 1, hi = p.Tmax - 1;          // largest t with step_start[t] <= n
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (p.step_start[mid] <= n) lo = mid; else hi = mid - 1;
                }
                t = lo;
                const long long r = n - p.step_start[t];
                rb = r / p.NCB; cb = r % p.NCB;
            }
            sItem[0] = n; sItem[1] = t; sItem[2] = rb; sItem[3] = cb;
        }
        __syncthreads();
        const long long item = sItem[0];
        const int t = (int)sItem[1], rb = (int)sItem[2], cb = (int)sItem[3];
        if (item >= total_items) break;

        const size_t slab = (size_t)Kl * p.Bpad;
        const double *dsrc = p.hist + (size_t)(t - 1) * slab + (size_t)rb * LG_BM;
        const double *asrc = p.A + (size_t)cb * LARGE_BN;

        if (w == LG_CONSUMER_WARPS) {
            // =================== producer warp: one lane feeds the TMA ring ===================
            if (lane == 0) {
                // wait until step t-1 of this row block is fully published (step 1 reads the zeroed slab)
                const unsigned int need = (unsigned int)(t - 1) * (unsigned int)p.NCB;
                while (ld_acquire_u32(p.done + rb) < need) { __nanosleep(64); }
                asm volatile("fence.proxy.async;" ::: "memory");
                for (int c = 0; c < nchunks; c++) {
                    const uint32_t g = chunk_ctr + c;
                    const int s = g % LG_STAGES;
                    if (g >= LG_STAGES) mbar_wait(empty + s, ((g / LG_STAGES) - 1) & 1);
                    mbar_expect_tx(full + s, (uint32_t)((LG_STAGE_A + LG_STAGE_D) * 8));
                    double *dstA = sA + (size_t)s * LG_STAGE_A;
                    double *dstD = sD + (size_t)s * LG_STAGE_D;
                    const int j0 = c * LG_BK;
#pragma unroll 4
                    for (int jj = 0; jj < LG_BK; jj++) {
                        tma_bulk_g2s(dstA + jj * LARGE_BN, asrc + (size_t)(j0 + jj) * Kl, LARGE_BN * 8, full + s);
                        tma_bulk_g2s(dstD + jj * LG_BM, dsrc + (size_t)(j0 + jj) * p.Bpad, LG_BM * 8, full + s);
                    }
                }
            }
            chunk_ctr += nchunks;
        } else {
            // =================== consumer warps ===================
            const int i0 = cb * LARGE_BN + w * TQ;        // global target state of q = 0
            const int r0 = rb * LG_BM + lane * TP;        // column (rank inside the group) of p = 0
            double best[TP][TQ];
#pragma unroll
            for (int q = 0; q < TP; q++)
#pragma unroll
                for (int k = 0; k < TQ; k++) best[q][k] = neg_inf();

            for (int c = 0; c < nchunks; c++) {
                const uint32_t g = chunk_ctr + c;
                const int s = g % LG_STAGES;
                mbar_wait(full + s, (g / LG_STAGES) & 1);
                const int nj = min(LG_BK, K - c * LG_BK);
                if (nj == LG_BK)
                    maxplus_accum_val<LG_BK>(sD + (size_t)s * LG_STAGE_D + lane * TP, LG_BM,
                                             sA + (size_t)s * LG_STAGE_A + w * TQ, LARGE_BN, nj, best);
                else
                    maxplus_accum_val<0>(sD + (size_t)s * LG_STAGE_D + lane * TP, LG_BM,
                                         sA + (size_t)s * LG_STAGE_A + w * TQ, LARGE_BN, nj, best);
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s);
            }
            chunk_ctr += nchunks;

            // ---- epilogue: emission ((delta + a) + b, viterbi.rs:17; -inf emission => -inf) and the slice of slab t ----
            double *dnext = p.hist + (size_t)t * slab;
#pragma unroll
            for (int q = 0; q < TP; q++) {
                const int r = r0 + q;
                const int64_t rg = p.rank0 + r;
                if (rg >= p.B) continue;
                const int len = (int)p.sorted_len[rg];
                if (t >= len) continue;
                const uint32_t b = p.order[rg];
                const int64_t pos = p.seq_off[b] + t;
                uint32_t o = p.obs[pos];
                if ((int64_t)o >= p.M) { *p.status = 3; o = 0; }
                const double *em = p.BT + (size_t)o * Kl + i0;
#pragma unroll
                for (int k = 0; k < TQ; k++)
                    if (i0 + k < K) dnext[(size_t)(i0 + k) * p.Bpad + r] = best[q][k] + __ldg(em + k);
            }
            __threadfence();   // publish this thread's delta slice before the item is counted done
        }
        __syncthreads();
        if (tid == 0) atomicAdd(p.done + rb, 1u);   // release: all consumers fenced before the barrier
    }
}

// End state (viterbi.rs:24) and backtrace (viterbi.rs:25-30) with lazy backpointers, 8 lanes per sequence:
// lane `sub` scans predecessors j = sub, sub+8, ... of the current path state (delta row from the history,
// logA column from the transposed copy, both 64-byte runs per sequence), then the 8 lanes reduce
// (value, index) -- strictly greater value, else lower index.  psi = 0 whenever delta[t][cur] = -inf
// (emission -inf or all candidates -inf; see decode_small.cuh).
__global__ void __launch_bounds__(256) backtrace_large_kernel(const DecodeLargeParams p)
{
    constexpr int L = 8;
    const int sub = threadIdx.x % L;
    const int64_t gi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;      // column inside the group
    const int64_t rg = p.rank0 + gi;
    const bool valid = gi < (int64_t)p.NRB * LG_BM && rg < p.B;
    const int len = valid ? (int)p.sorted_len[rg] : 0;
    const uint32_t b = valid ? p.order[rg] : 0u;
    const int64_t off = valid ? p.seq_off[b] : 0;
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    const size_t slab = (size_t)p.Kl * p.Bpad;
    const double *col = p.hist + gi;
    auto reduce = [&](double &v, int &ix) {
#pragma unroll
        for (int d = L / 2; d >= 1; d >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, d, L);
            const int oi = __shfl_xor_sync(0xffffffffu, ix, d, L);
            if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
        }
    };
    // end state
    double bv = neg_inf(); int cur = 0x7fffffff;
    if (valid) {
        const double *row = col + (size_t)(len - 1) * slab;
        for (int j = sub; j < p.K; j += L) {
            const double v = __ldcs(row + (size_t)j * p.Bpad);
            if (cur == 0x7fffffff || v > bv) { bv = v; cur = j; }
        }
    }
    reduce(bv, cur);
    if (valid && sub == 0) {
        if (p.score) p.score[b] = bv;
        p.path[off + len - 1] = (uint32_t)cur;
    }
    double dcur = bv;
    for (int tt = maxlen - 1; tt >= 1; tt--) {
        const bool act = valid && tt < len;
        double mv = neg_inf(); int mi = 0x7fffffff;
        if (act && dcur > neg_inf()) {
            const double *row = col + (size_t)(tt - 1) * slab;
            const double *at = p.AT + (size_t)cur * p.Kl;
#pragma unroll 4
            for (int j = sub; j < p.K; j += L) {
                const double v = __ldcs(row + (size_t)j * p.Bpad) + __ldg(at + j);
                if (mi == 0x7fffffff || v > mv) { mv = v; mi = j; }
            }
        }
        reduce(mv, mi);
        if (act) {
            cur = (dcur > neg_inf()) ? mi : 0;
            dcur = __ldcg(col + (size_t)(tt - 1) * slab + (size_t)cur * p.Bpad);
            if (sub == 0) p.path[off + tt - 1] = (uint32_t)cur;
        }
    }
}

}  // namespace cvb
