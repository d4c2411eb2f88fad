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

// Wait until every rank has published exchange `epoch` into THIS rank's buffer.  Bounded: after ~4 s the error
// word is set and the kernel returns (the host then fails the solve instead of hanging the GPU).
__global__ void cp_peer_wait_kernel(const unsigned int *flags, int R, unsigned int epoch, int *err)
{
    const int q = threadIdx.x;
    if (q >= R) return;
    const long long t0 = clock64();
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + q) : "memory");
        if ((int)(v - epoch) >= 0) break;
        if (clock64() - t0 > (1LL << 33)) { *err = 1; break; }
        __nanosleep(64);
    }
}

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
// written for the rows the rank owns (t < own_hi) and for rlo.  Chunks of CP_BT_CHUNK rows, map composition as
// in the single-GPU path: (1) per chunk and entry state the exit state, (2) K threads chain the chunks into the
// rank's total map and publish it, (3) with every rank's map known, the rank's entry state, (4) replay.
__global__ void cp_btr_maps_kernel(const CpParams p, int64_t rlo, int64_t rtop, int nchunks, psi_t *F)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)nchunks * p.K) return;
    const int c = (int)(gid / p.K); int cur = (int)(gid % p.K);
    const int64_t hi = min(rlo + (int64_t)(c + 1) * CP_BT_CHUNK, rtop), lo = rlo + (int64_t)c * CP_BT_CHUNK;
    for (int64_t t = hi; t > lo; t--) cur = p.psi[(size_t)t * p.K + cur];
    F[gid] = (psi_t)cur;
}

// maps region of the exchange buffer: [2][CP_MAX_RANKS][mapw] psi_t, entry K of a rank's map = its end state
// (argmax of delta[N-1], meaningful on the last rank only)
__global__ void cp_btr_total_kernel(const CpParams p, int nchunks, const psi_t *F, const int *end_state, const PeerTab pt,
                                    size_t maps_off, int mapw, unsigned int epoch)
{
    const int e = threadIdx.x;
    if (e < p.K) {
        int cur = e;
        for (int c = nchunks - 1; c >= 0; c--) cur = F[(size_t)c * p.K + cur];
        for (int q = 0; q < pt.R; q++)
            reinterpret_cast<psi_t *>(pt.base[q] + maps_off)[(size_t)pt.rank * mapw + e] = (psi_t)cur;
    } else if (e == p.K) {
        for (int q = 0; q < pt.R; q++)
            reinterpret_cast<psi_t *>(pt.base[q] + maps_off)[(size_t)pt.rank * mapw + p.K] = (psi_t)*end_state;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) peer_signal_all(pt, epoch);
}

// entry[c] = state at the last row of chunk c; the rank's own entry comes from the end state pushed through the
// maps of the ranks above it (maps == nullptr: single rank, entry = *end_state)
__global__ void cp_btr_chain_kernel(const CpParams p, int nchunks, const psi_t *F, const int *end_state, const psi_t *maps,
                                    int mapw, int rank, int R, int *entry)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int cur;
    if (maps) {
        cur = maps[(size_t)(R - 1) * mapw + p.K];
        for (int q = R - 1; q > rank; q--) cur = maps[(size_t)q * mapw + cur];
    } else {
        cur = *end_state;
    }
    for (int c = nchunks - 1; c >= 0; c--) { entry[c] = cur; cur = F[(size_t)c * p.K + cur]; }
}

__global__ void cp_btr_fill_kernel(const CpParams p, int64_t rlo, int64_t rtop, int64_t own_hi, int nchunks, const int *entry,
                                   uint64_t *sol)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    int cur = entry[c];
    const int64_t hi = min(rlo + (int64_t)(c + 1) * CP_BT_CHUNK, rtop), lo = rlo + (int64_t)c * CP_BT_CHUNK;
    for (int64_t t = hi; t > lo; t--) {
        if (t < own_hi) sol[t] = (uint64_t)cur;
        cur = p.psi[(size_t)t * p.K + cur];
    }
    if (c == 0) sol[rlo] = (uint64_t)cur;
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
