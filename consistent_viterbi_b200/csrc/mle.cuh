// mle.cuh -- event counts of the supervised maximum-likelihood estimate.
//
// Replaces the counting loop of HMM::maximum_likelihood_estimation (reference src/hmm/hmm.rs:35-48): for every
// tagged sequence  pi[tag_0] += 1;  b[tag_t][obs_t] += 1, seen[tag_t] += 1 for every t;  a[tag_t][tag_t+1] += 1
// for t < T-1;  end[tag_T-1] += 1.  The reference adds 1.0 in f64 on top of a random initial model; WHICH value
// an entry ends with depends only on how many times it was incremented, so the device counts events in u64
// (exact, order-free) and the host replays "+= 1.0 count times" on the initial value (mle_host.inl).
// HBM-bound byte work: ~10 B per element (obs u32 + tag i32 + sequence-start flags); four elements per thread, a / seen / end / pi counters
// privatised in shared memory per CTA when K*K fits, b counters straight to L2 atomics (K*M is MBs).
#pragma once
#include "common.cuh"

namespace cvb {

struct MleParams {
    const uint32_t *obs;      // [N]
    const int32_t *tags;      // [N], < 0 = None
    const int64_t *seq_off;   // [B+1]
    const uint8_t *first;     // [N+1] 1 at the first element of every sequence and at N (mle_mark_kernel)
    int64_t N, B, M;
    int K;
    unsigned long long *a, *b, *pi, *seen, *end;   // [K*K] [K*M] [K] [K] [K]
    int *status;              // bit 0: tag None / >= K, bit 1: observation >= M
};

// first[seq_off[b]] = 1 for every sequence, first[N] = 1 (the array is zeroed before): an element is the last of
// its sequence iff first[e + 1], so the count kernel needs no per-element sequence lookup
__global__ void mle_mark_kernel(const int64_t *seq_off, int64_t B, uint8_t *first)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b <= B) first[seq_off[b]] = 1;
}

constexpr int MLE_THREADS = 256;
constexpr int MLE_SMEM_K = 64;        // K <= 64: a (K*K), seen, end, pi counters privatised per CTA (u32)

__global__ void __launch_bounds__(MLE_THREADS) mle_count_kernel(const MleParams p)
{
    extern __shared__ unsigned int mle_sh[];
    const int K = p.K;
    const bool priv = K <= MLE_SMEM_K;
    unsigned int *sa = mle_sh, *sseen = sa + K * K, *send = sseen + K, *spi = send + K;
    if (priv) {
        for (int i = threadIdx.x; i < K * K + 3 * K; i += blockDim.x) mle_sh[i] = 0u;
        __syncthreads();
    }
    int bad = 0;
    // 4 elements per thread and pass: every load below is independent of the others
    constexpr int U = 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * U;
    for (int64_t e0 = ((int64_t)blockIdx.x * blockDim.x) * U + threadIdx.x; e0 < p.N; e0 += stride) {
        int tg[U], tn[U]; uint32_t o[U]; bool first[U], last[U], live[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t e = e0 + (int64_t)u * blockDim.x;
            live[u] = e < p.N;
            const int64_t ec = live[u] ? e : 0;
            tg[u] = p.tags[ec]; o[u] = p.obs[ec];
            first[u] = p.first[ec] != 0; last[u] = p.first[ec + 1] != 0;
            tn[u] = (live[u] && e + 1 < p.N) ? p.tags[e + 1] : -1;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!live[u]) continue;
            if (tg[u] < 0 || tg[u] >= K) { bad |= 1; continue; }
            if ((int64_t)o[u] >= p.M) { bad |= 2; continue; }
            atomicAdd(p.b + (size_t)tg[u] * p.M + o[u], 1ULL);                  // hmm.rs:41,45
            int nx = -1;
            if (!last[u]) { nx = tn[u]; if (nx < 0 || nx >= K) { bad |= 1; nx = -1; } }
            if (priv) {
                atomicAdd(sseen + tg[u], 1u);                                    // :43,46
                if (first[u]) atomicAdd(spi + tg[u], 1u);                        // :39
                if (last[u]) atomicAdd(send + tg[u], 1u);                        // :47
                else if (nx >= 0) atomicAdd(sa + tg[u] * K + nx, 1u);            // :42
            } else {
                atomicAdd(p.seen + tg[u], 1ULL);
                if (first[u]) atomicAdd(p.pi + tg[u], 1ULL);
                if (last[u]) atomicAdd(p.end + tg[u], 1ULL);
                else if (nx >= 0) atomicAdd(p.a + (size_t)tg[u] * K + nx, 1ULL);
            }
        }
    }
    if (bad) atomicOr(p.status, bad);
    if (priv) {
        __syncthreads();
        for (int i = threadIdx.x; i < K * K; i += blockDim.x) if (sa[i]) atomicAdd(p.a + i, (unsigned long long)sa[i]);
        for (int i = threadIdx.x; i < K; i += blockDim.x) {
            if (sseen[i]) atomicAdd(p.seen + i, (unsigned long long)sseen[i]);
            if (send[i]) atomicAdd(p.end + i, (unsigned long long)send[i]);
            if (spi[i]) atomicAdd(p.pi + i, (unsigned long long)spi[i]);
        }
    }
}

}  // namespace cvb
