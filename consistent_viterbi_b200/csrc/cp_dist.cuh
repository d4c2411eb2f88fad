// cp_dist.cuh -- constrained decode (mode R2) sharded over the GPUs of one NVSwitch box.
//
// One B&B node's sweeps are independent, so the super-sequence is cut into one contiguous row range
// per rank.  The cuts are positions of component 0: component 0 is assigned in every sweep of solve_r
// (cp.rs:95-126), and a sweep stops at the next position whose component is assigned (cp.rs:48), so NO sweep
// ever crosses a cut -- the ranks never exchange delta rows.  Ownership (lo/hi = the rank's cut pair):
//   delta rows [lo, hi)     sweeps that start in [lo, hi), the clamp resets, the C1 fix-up of pos in [lo, hi)
//   psi rows   (lo, hi]     the C2 fix-up and the bound term of a clamped position t read delta[t-1], so they
//                           belong to the owner of row t-1; backtrack steps t -> t-1 likewise
// What IS exchanged per node is the list of bound terms (cp.rs:103-116): the kernel that computes a rank's
// terms stores them straight into every peer's term buffer (NVLink peer stores through CUDA-IPC mappings, at
// the term's index in the reference's summation order), then raises a flag in every peer; each rank waits for
// all flags and runs the same exact-order sum, so every rank takes the same branch-and-bound decision without
// any other message.  A backtrack exchanges one K-entry map (state at the rank's upper cut -> state at its
// lower cut) per rank the same way; the final solution is published range by range.
// No NCCL on this path: the payloads are a few KB to ~1 MB and latency-bound, a fused store+flag costs one
// NVLink round trip.  Buffers that a fast rank could overwrite while a slow one still reads them are double
// buffered (a rank can be at most one exchange ahead of any peer).
#pragma once
#include "cp_kernels.cuh"

namespace cvb {

constexpr int CP_MAX_RANKS = 8;
constexpr int CP_FLAGS_BYTES = 256;

// the same layout on every rank; base[q] = this process's mapping of rank q's exchange buffer
struct PeerTab {
    unsigned char *base[CP_MAX_RANKS];
    int R, rank;
};

__device__ __forceinline__ void peer_signal_all(const PeerTab &pt, unsigned int epoch)
{
    for (int q = 0; q < pt.R; q++) {
        unsigned int *f = reinterpret_cast<unsigned int *>(pt.base[q]) + pt.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
    }
}

// Block-collective epilogue of a publishing kernel: every thread has issued its peer stores; the last block to
// arrive raises this rank's flag (value = exchange number) on every peer.
__device__ __forceinline__ void peer_publish_done(const PeerTab &pt, unsigned int epoch, unsigned int *done_counter)
{
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_counter, 1u);
        if (prev == gridDim.x - 1) {
            *done_counter = 0u;
            __threadfence_system();
            peer_signal_all(pt, epoch);
        }
    }
}

// Wait until every rank has published exchange `epoch` into THIS rank's buffer (peer_wait_block, cp_kernels.cuh;
// the bound sum waits inside its first kernel, the backtrack and the final gather use this stand-alone form).
__global__ void cp_peer_wait_kernel(const PeerWait pw) { peer_wait_block(pw); }

// Phase D on a shard: this rank's bound terms (clamped positions t in (lo, hi], t = 0 on the first rank), each
// stored at its index in the reference's summation order in EVERY rank's term buffer, then the flag.
__global__ void __launch_bounds__(SUM_BLK) cp_terms_peer_kernel(const CpParams p, const int64_t *term_pos, const int32_t *term_comp,
                                                              const int32_t *term_gidx, int nlocal, const PeerTab pt,
                                                              size_t terms_off, unsigned int epoch, unsigned int *done_counter)
{
    const int k = blockIdx.x * SUM_BLK + threadIdx.x;
    if (k < nlocal) {
        const int64_t t = term_pos[k];
        const int st = p.choice[term_comp[k]];
        const int K = p.K, Kp = p.Kp;
        const double b = p.BT[(size_t)p.obs[t] * Kp + st];
        double term;
        if (t == 0) {
            term = p.Pi[st] + b;
        } else {
            const int sf = p.psi[(size_t)t * K + st];
            const double arc = (p.start[t] ? p.Pi[st] : p.A[(size_t)sf * Kp + st]) + b;
            term = p.delta[(size_t)(t - 1) * K + sf] + arc;
        }
        const int g = term_gidx[k];
        for (int q = 0; q < pt.R; q++) reinterpret_cast<double *>(pt.base[q] + terms_off)[g] = term;
    }
    peer_publish_done(pt, epoch, done_counter);
}

// ---- backtrack (cp.rs:85-93) over a row range ------------------------------------------------------------
// A rank walks the backpointer rows t = rtop .. rlo+1 (cur = psi[t][cur] is the state at t-1).  rtop is the
// rank's upper cut (a row of the next rank whose STATE is the walk's entry) or N-1 on the last rank; sol[t] is
// written for the rows the rank owns (t < own_hi) and for rlo.  The walk is a chain of N dependent loads, so it
// is cut into chunks of CP_BT_CHUNK rows and done as a composition of maps:
//   (1) maps   per chunk and entry state the exit state        F[c][e]      one CTA per chunk, rows staged in smem
//   (2) total  the rank's map M = F[0] o .. o F[n-1], published to every rank (sharded solve only)
//   (3) chain  entry[c] = state at the last row of chunk c, from the end state pushed through the maps of the
//              ranks above and the chunks above; F staged in smem, groups of BTR_GROUP chunks chained in parallel
//   (4) fill   every chunk replays its walk from entry[c] and writes sol
// Integer-exact; the dependent steps run against shared memory (~30 clk) instead of L2 (~500 clk).
constexpr int BTR_GROUP = 64;
constexpr size_t BTR_SMEM_MAX = 200 * 1024;

__host__ __device__ inline size_t btr_rows_smem_bytes(int K) { return ((size_t)CP_BT_CHUNK * K * sizeof(psi_t) + 15) / 16 * 16 + 16; }
__host__ __device__ inline size_t btr_chain_smem_bytes(int nchunks, int K)
{
    const int ngroups = (nchunks + BTR_GROUP - 1) / BTR_GROUP;
    return ((size_t)nchunks * K * sizeof(psi_t) + 15) / 16 * 16 + ((size_t)ngroups * K * sizeof(psi_t) + 15) / 16 * 16 +
           (size_t)ngroups * sizeof(int) + 16;
}

// rows (lo, hi] of psi into shared memory: srow[(t - lo - 1) * K + s]
__device__ __forceinline__ void btr_stage_rows(const CpParams &p, int64_t lo, int64_t hi, psi_t *srow)
{
    const size_t n = (size_t)(hi - lo) * p.K;
    const psi_t *src = p.psi + (size_t)(lo + 1) * p.K;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const size_t n8 = n / 8;
        for (size_t i = threadIdx.x; i < n8; i += blockDim.x) reinterpret_cast<uint4 *>(srow)[i] = reinterpret_cast<const uint4 *>(src)[i];
        for (size_t i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) srow[i] = src[i];
    } else {
        for (size_t i = threadIdx.x; i < n; i += blockDim.x) srow[i] = src[i];
    }
    __syncthreads();
}

// one CTA per chunk (64 threads); rows_in_smem = 0: K too large to stage, walk global memory
__global__ void __launch_bounds__(64) cp_btr_maps_kernel(const CpParams p, int64_t rlo, int64_t rtop, int nchunks, psi_t *F,
                                                        int rows_in_smem)
{
    extern __shared__ __align__(16) unsigned char btr_raw[];
    psi_t *srow = reinterpret_cast<psi_t *>(btr_raw);
    const int c = blockIdx.x;
    const int64_t hi = min(rlo + (int64_t)(c + 1) * CP_BT_CHUNK, rtop), lo = rlo + (int64_t)c * CP_BT_CHUNK;
    if (rows_in_smem) btr_stage_rows(p, lo, hi, srow);
    for (int e = threadIdx.x; e < p.K; e += blockDim.x) {
        int cur = e;
        if (rows_in_smem) for (int64_t t = hi; t > lo; t--) cur = srow[(size_t)(t - lo - 1) * p.K + cur];
        else for (int64_t t = hi; t > lo; t--) cur = p.psi[(size_t)t * p.K + cur];
        F[(size_t)c * p.K + e] = (psi_t)cur;
    }
}

// F into shared memory and the group maps GF[g][e] (exit state of group g for entry state e); all threads
__device__ __forceinline__ void btr_stage_groups(const psi_t *F, int nchunks, int K, psi_t *sF, psi_t *sGF)
{
    const size_t n = (size_t)nchunks * K;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) sF[i] = F[i];
    __syncthreads();
    const int ngroups = (nchunks + BTR_GROUP - 1) / BTR_GROUP;
    for (int i = threadIdx.x; i < ngroups * K; i += blockDim.x) {
        const int g = i / K; int cur = i % K;
        for (int c = min((g + 1) * BTR_GROUP, nchunks) - 1; c >= g * BTR_GROUP; c--) cur = sF[(size_t)c * K + cur];
        sGF[i] = (psi_t)cur;
    }
    __syncthreads();
}

// maps region of the exchange buffer: [2][CP_MAX_RANKS][mapw] psi_t, entry K of a rank's map = its end state
// (argmax of delta[N-1], meaningful on the last rank only).  in_smem = 0: F too large, chained from global.
__global__ void __launch_bounds__(1024) cp_btr_total_kernel(const CpParams p, int nchunks, const psi_t *F, const int *end_state,
                                                           const PeerTab pt, size_t maps_off, int mapw, unsigned int epoch, int in_smem)
{
    extern __shared__ __align__(16) unsigned char btr_raw[];
    const int K = p.K, ngroups = (nchunks + BTR_GROUP - 1) / BTR_GROUP;
    psi_t *sF = reinterpret_cast<psi_t *>(btr_raw);
    psi_t *sGF = reinterpret_cast<psi_t *>(btr_raw + ((size_t)nchunks * K * sizeof(psi_t) + 15) / 16 * 16);
    if (in_smem) btr_stage_groups(F, nchunks, K, sF, sGF);
    for (int e = threadIdx.x; e <= K; e += blockDim.x) {
        int cur = e;
        if (e == K) cur = *end_state;
        else if (in_smem) for (int g = ngroups - 1; g >= 0; g--) cur = sGF[(size_t)g * K + cur];
        else for (int c = nchunks - 1; c >= 0; c--) cur = F[(size_t)c * K + cur];
        for (int q = 0; q < pt.R; q++)
            reinterpret_cast<psi_t *>(pt.base[q] + maps_off)[(size_t)pt.rank * mapw + e] = (psi_t)cur;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) peer_signal_all(pt, epoch);
}

// maps == nullptr: single rank, the entry is *end_state; pw: the ranks' maps must have arrived
__global__ void __launch_bounds__(1024) cp_btr_chain_kernel(const CpParams p, int nchunks, const psi_t *F, const int *end_state,
                                                           const psi_t *maps, int mapw, int rank, int R, int *entry, int in_smem,
                                                           const PeerWait pw)
{
    extern __shared__ __align__(16) unsigned char btr_raw[];
    peer_wait_block(pw);
    const int K = p.K, ngroups = (nchunks + BTR_GROUP - 1) / BTR_GROUP;
    int cur = 0;
    if (threadIdx.x == 0) {
        if (maps) {
            cur = maps[(size_t)(R - 1) * mapw + K];
            for (int q = R - 1; q > rank; q--) cur = maps[(size_t)q * mapw + cur];
        } else {
            cur = *end_state;
        }
    }
    if (!in_smem) {
        if (threadIdx.x == 0) for (int c = nchunks - 1; c >= 0; c--) { entry[c] = cur; cur = F[(size_t)c * K + cur]; }
        return;
    }
    psi_t *sF = reinterpret_cast<psi_t *>(btr_raw);
    psi_t *sGF = reinterpret_cast<psi_t *>(btr_raw + ((size_t)nchunks * K * sizeof(psi_t) + 15) / 16 * 16);
    int *gentry = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(sGF) + ((size_t)ngroups * K * sizeof(psi_t) + 15) / 16 * 16);
    btr_stage_groups(F, nchunks, K, sF, sGF);
    if (threadIdx.x == 0) for (int g = ngroups - 1; g >= 0; g--) { gentry[g] = cur; cur = sGF[(size_t)g * K + cur]; }
    __syncthreads();
    for (int g = threadIdx.x; g < ngroups; g += blockDim.x) {
        int c2 = gentry[g];
        for (int c = min((g + 1) * BTR_GROUP, nchunks) - 1; c >= g * BTR_GROUP; c--) { entry[c] = c2; c2 = sF[(size_t)c * K + c2]; }
    }
}

// one CTA per chunk (64 threads): thread 0 replays the walk against the staged rows, all threads write sol
__global__ void __launch_bounds__(64) cp_btr_fill_kernel(const CpParams p, int64_t rlo, int64_t rtop, int64_t own_hi, int nchunks,
                                                        const int *entry, uint64_t *sol, int rows_in_smem)
{
    extern __shared__ __align__(16) unsigned char btr_raw[];
    __shared__ psi_t st[CP_BT_CHUNK + 1];
    psi_t *srow = reinterpret_cast<psi_t *>(btr_raw);
    const int c = blockIdx.x;
    const int64_t hi = min(rlo + (int64_t)(c + 1) * CP_BT_CHUNK, rtop), lo = rlo + (int64_t)c * CP_BT_CHUNK;
    if (rows_in_smem) btr_stage_rows(p, lo, hi, srow);
    if (threadIdx.x == 0) {
        int cur = entry[c];
        for (int64_t t = hi; t > lo; t--) {                    // st[t - lo] = state at row t
            st[t - lo] = (psi_t)cur;
            cur = rows_in_smem ? srow[(size_t)(t - lo - 1) * p.K + cur] : p.psi[(size_t)t * p.K + cur];
        }
        st[0] = (psi_t)cur;
    }
    __syncthreads();
    for (int64_t t = lo + 1 + threadIdx.x; t <= hi; t += blockDim.x)
        if (t < own_hi) sol[t] = (uint64_t)st[t - lo];
    if (c == 0 && threadIdx.x == 0) sol[rlo] = (uint64_t)st[0];
}

// the rank's rows of the solution, stored into every rank's solution buffer, then the flag
__global__ void cp_sol_publish_kernel(const uint64_t *sol, int64_t lo, int64_t hi, const PeerTab pt, size_t sol_off,
                                      unsigned int epoch, unsigned int *done_counter)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < hi; t += stride) {
        const uint64_t v = sol[t];
        for (int q = 0; q < pt.R; q++)
            if (q != pt.rank) reinterpret_cast<uint64_t *>(pt.base[q] + sol_off)[t] = v;
    }
    peer_publish_done(pt, epoch, done_counter);
}

}  // namespace cvb
