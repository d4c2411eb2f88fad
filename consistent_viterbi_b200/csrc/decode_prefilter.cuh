// decode_prefilter.cuh -- batched plain Viterbi (mode R1), K <= 64: forward kernel with an f32 PRE-FILTER.
//
// Replaces B calls of viterbi::decode (reference src/viterbi_solver/viterbi.rs:5-32), like decode_small.cuh, and
// produces the same bits.  Why a second forward kernel: a max-plus cell in f64 is DADD + DSETP + 2 x FSEL on
// sm_100a (no 64-bit select, no DMNMX) -- four issue slots, two on the half-rate FP64 pipe and two on the half-rate
// ALU pipe -- and decode_small_fwd_kernel sits at ~0.50 of the FP64 issue peak with every pipe and the issue port
// equally loaded.  Nothing at four instructions per cell gets past that.  This kernel spends ~2 instructions per
// cell on a 32-bit pre-filter and f64 work only where the maximum can be:
//
//   pass 1 (f32, all K predecessors):  y_j = fl32(d32_j + a32_ji) with packed adds (FADD2, two cells per
//       instruction) and 3-input maxima (FMNMX3, two cells per instruction), per block of 8 predecessors; the block
//       maximum gets the block id in its 3 low mantissa bits and is merged into the running best r and runner-up s
//       (3 FMNMX);
//   pass 2 (f64, exact):  for every (sequence, target state) the 8 predecessors of the winning block are scanned in
//       f64 in ascending order with a strict > (first maximum), exactly as the plain kernel scans all of them.
//
// Exactness.  All finite model entries are <= 0 (checked on the host; a model with a positive entry takes the plain
// kernel), so every delta and every candidate x_j = delta_j + a_ji is <= 0 or -inf.  With d32 = rn32(delta_j),
// a32 = rn32(a_ji): |d32 + a32 - x_j| <= 2^-24 |x_j| (same signs: no cancellation), the f32 add adds 2^-24, clearing
// the three low mantissa bits 7 * 2^-23: key_j in x_j * (1 +- 2^-19.8).  Let W be the block whose key is r and s the
// best key of any other block.  If the true maximum x* were outside W then s >= x* (1 + eta) and r <= x* (1 - eta),
// eta = 2^-19.8, i.e. s >= r (1 + 2^-18).  So whenever s < thr = r (1 + 2^-18) - 2^-120 the maximum over all
// predecessors is the maximum over block W, and -- f64 rounding being monotone -- pass 2 returns the bits the full
// scan returns (an equal candidate elsewhere would have s ~ r).  Otherwise (probability ~1e-4 per pair; also when r
// is below -1e29) pass 2 scans ALL predecessors in f64: never wrong, occasionally slow.  -inf never enters the f32
// pass (the key's low bits would turn it into a NaN): rn32 values are clamped to PF_NEG = -1e30, sums stay finite,
// and a winner below -1e29 -- a clamped or absurdly small value -- sends the pair to the full scan.
//
// Layouts (shared memory, per CTA = one tile of 64 sequences, G = ceil(K/8) warps, a lane owns sequences lane and
// lane + 32, a warp owns one group of <= 8 target states as in the balanced split of decode_small.cuh):
//   sA32  [Kp][Kp]      f32  rn32(logA[j][state of slot c]), slot-permuted columns, -inf padding
//   sA64T [Kp][Kp]      f64  logA transposed: row = slot c, column = predecessor j (8 consecutive j = 4 LDS.128)
//   sD64  [2][64][PD]   f64  delta, [sequence][state], PD = Kp + 2 (conflict-free 16-byte row reads per lane);
//                            the emission row of the next observation is TMA-copied INTO the row the step will
//                            write (delta = best + b in place): no separate emission stage
//   sD32  [2][Kp][64]   f32  rn32(delta), [state][column], column 2*lane + p = sequence lane + 32 p
// History: the sD64 buffer of every step leaves as one TMA bulk store, slab (tile, t) = [64][PD] doubles; the
// backtrace (backtrace_small_kernel, LAYOUT = 1) reads a sequence's row as 16-byte vectors.
#pragma once

#include "decode_small.cuh"

namespace cvb {

constexpr float PF_NEG = -1e30f;      // stand-in for -inf (and anything below) in the f32 pass
__host__ __device__ inline int pf_pitch(int Kp) { return Kp + 2; }
__host__ __device__ inline size_t decode_pf_smem_bytes(int Kp)
{
    return (size_t)Kp * Kp * 4 + (size_t)Kp * Kp * 8 + (size_t)2 * 64 * pf_pitch(Kp) * 8 + (size_t)2 * Kp * 64 * 4 +
           64 * (8 + 4) + 32 + 16;
}

__device__ __forceinline__ float2 fadd2(float2 a, float2 b)
{
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}

__global__ void __launch_bounds__(256, 2) decode_pf_fwd_kernel(const DecodeSmallParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NS = 64;
    const int K = p.K, Kp = p.Kp, PD = pf_pitch(Kp), NB = Kp >> 3;
    float *sA32 = reinterpret_cast<float *>(smem_raw);
    double *sA64T = reinterpret_cast<double *>(sA32 + (size_t)Kp * Kp);
    double *sD64 = sA64T + (size_t)Kp * Kp;
    float *sD32 = reinterpret_cast<float *>(sD64 + (size_t)2 * NS * PD);
    int64_t *sOff = reinterpret_cast<int64_t *>(sD32 + (size_t)2 * Kp * NS);
    int *sLen = reinterpret_cast<int *>(sOff + NS);
    uint64_t *sBar = reinterpret_cast<uint64_t *>(sLen + NS);   // [0] logA copies, [1] emissions
    int *sTile = reinterpret_cast<int *>(sBar + 4);

    const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
    const int nw = (int)blockDim.x >> 5;
    const int c0 = g * 8;                                                        // first slot of the group
    const int srow0 = p.nq_base ? g * p.nq_base + min(g, p.nq_rem) : c0;         // its first state
    const int nreal = p.nq_base ? p.nq_base + (g < p.nq_rem ? 1 : 0) : max(0, min(8, K - c0));
    const uint32_t slab_bytes = (uint32_t)((size_t)NS * PD * 8);
    const uint32_t row_bytes = (uint32_t)(Kp * 8);
    constexpr int EMK = 4;

    if (tid == 0) {
        if (p.started) atomicAdd(p.started, 1u);
        mbar_init(sBar, 1);
        mbar_init(sBar + 1, nw);
        fence_proxy_async_smem();
        const uint32_t b32 = (uint32_t)((size_t)Kp * Kp * 4), b64 = (uint32_t)((size_t)Kp * Kp * 8);
        mbar_expect_tx(sBar, b32 + b64);
        tma_bulk_g2s(sA32, p.A32s, b32, sBar);
        tma_bulk_g2s(sA64T, p.A64Ts, b64, sBar);
    }
    __syncthreads();
    mbar_wait(sBar, 0);
    uint32_t em_phase = 0;

    const bool obs_streamed = p.arrived != nullptr;
    auto ld_obs = [&](int64_t idx) -> uint32_t { return load_obs_at(p, idx, obs_streamed); };
    // emission rows of step t go straight into the delta buffer the step writes (row of the sequence, columns 0..Kp)
    auto issue_emissions = [&](int t, const uint32_t (&o_cur)[EMK]) {
        double *dst = sD64 + (size_t)(t & 1) * NS * PD;
        int nact = 0;
#pragma unroll
        for (int k = 0; k < EMK; k++) {
            const int s = g + nw * (lane + 32 * k);
            if (s < NS && t < sLen[s]) nact++;
        }
        const int total = __reduce_add_sync(0xffffffffu, nact);
        if (lane == 0) mbar_expect_tx(sBar + 1, (uint32_t)total * row_bytes);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < EMK; k++) {
            const int s = g + nw * (lane + 32 * k);
            if (s < NS && t < sLen[s]) {
                uint32_t o = o_cur[k];
                if ((int64_t)o >= p.M) { *p.status = 3; o = 0; }   // index panic in the reference
                tma_bulk_g2s(dst + (size_t)s * PD, p.BT + (size_t)o * Kp, row_bytes, sBar + 1);
            }
        }
    };

    for (;;) {
        if (tid == 0) *sTile = (int)atomicAdd(p.tile_counter, 1u);
        __syncthreads();
        const int tile = *sTile;
        if (tile >= p.ntiles) break;

        for (int s = tid; s < NS; s += blockDim.x) {
            const int64_t r = (int64_t)tile * NS + s;
            int64_t off = 0; int len = 0;
            if (r < p.B) {
                const uint32_t b = p.order[r];
                off = p.seq_off[b];
                len = (int)(p.seq_off[b + 1] - off);
                if (p.is_long && p.is_long[b]) len = 0;
            }
            sOff[s] = off; sLen[s] = len;
        }
        // delta(0) = 0.0 (viterbi.rs:6); padding states -inf in both buffers (the emission copies keep them -inf)
        for (int e = tid; e < 2 * NS * PD; e += blockDim.x) {
            const int j = e % PD;
            sD64[e] = (e < NS * PD && j < K) ? 0.0 : neg_inf();
        }
        for (int e = tid; e < 2 * Kp * NS; e += blockDim.x) {
            const int j = (e / NS) % Kp;
            sD32[e] = (e < Kp * NS && j < K) ? 0.0f : PF_NEG;
        }
        if (p.arrived && tid == 0) {
            const unsigned int need = (unsigned int)p.tile_chunk[tile] + 1u;
            const long long t0 = clock64();
            for (;;) {
                unsigned int v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.arrived) : "memory");
                if (v >= need) break;
                if (clock64() - t0 > (1LL << 32)) { *p.status = 5; break; }
                __nanosleep(128);
            }
        }
        fence_proxy_async_smem();
        __syncthreads();

        int Tmax = p.tile_tmax ? (int)p.tile_tmax[tile] : sLen[0];
        if (p.tile_base[tile] + (long long)Tmax > p.hist_cap_slabs) {
            if (tid == 0) atomicMax(p.status, 7);
            Tmax = 0;
        }
        double *slab = p.hist + (size_t)p.tile_base[tile] * NS * PD;
        if (tid == 0 && Tmax > 0) tma_bulk_s2g(slab, sD64, slab_bytes);   // history slab 0 = delta(0)

        uint32_t o_nxt[EMK];
        {
            uint32_t o1[EMK];
#pragma unroll
            for (int k = 0; k < EMK; k++) {
                const int s = g + nw * (lane + 32 * k);
                const bool in = s < NS;
                o1[k] = (in && 1 < sLen[s]) ? ld_obs(sOff[s] + 1) : 0u;
                o_nxt[k] = (in && 2 < sLen[s]) ? ld_obs(sOff[s] + 2) : 0u;
            }
            // slab 0 has to be read out of buffer 0 ... it is buffer 1 that step 1 writes: no conflict
            if (Tmax > 1) issue_emissions(1, o1);
        }

        for (int t = 1; t < Tmax; t++) {
            const int cur = (t - 1) & 1, nxt = t & 1;
            // ---------------- pass 1: f32 pre-filter over all predecessors ----------------
            float r[2][8], s2[2][8];
#pragma unroll
            for (int pp = 0; pp < 2; pp++)
#pragma unroll
                for (int q = 0; q < 8; q++) { r[pp][q] = __int_as_float(0xff800000); s2[pp][q] = __int_as_float(0xff800000); }
            {
                const float *d32 = sD32 + (size_t)cur * Kp * NS + 2 * lane;
                const float *a32 = sA32 + c0;
                for (int blk = 0; blk < NB; blk++) {
                    float m[2][8];
#pragma unroll
                    for (int pp = 0; pp < 2; pp++)
#pragma unroll
                        for (int q = 0; q < 8; q++) m[pp][q] = __int_as_float(0xff800000);
#pragma unroll
                    for (int jj = 0; jj < 8; jj += 2) {
                        float y[2][2][8];
#pragma unroll
                        for (int u = 0; u < 2; u++) {
                            const int j = blk * 8 + jj + u;
                            const float2 d = *reinterpret_cast<const float2 *>(d32 + (size_t)j * NS);
                            const float4 a0 = *reinterpret_cast<const float4 *>(a32 + (size_t)j * Kp);
                            const float4 a1 = *reinterpret_cast<const float4 *>(a32 + (size_t)j * Kp + 4);
                            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                            for (int pp = 0; pp < 2; pp++) {
                                const float dp = pp ? d.y : d.x;
#pragma unroll
                                for (int q = 0; q < 8; q += 2) {
                                    const float2 v = fadd2(make_float2(dp, dp), make_float2(a[q], a[q + 1]));
                                    y[u][pp][q] = v.x; y[u][pp][q + 1] = v.y;
                                }
                            }
                        }
#pragma unroll
                        for (int pp = 0; pp < 2; pp++)
#pragma unroll
                            for (int q = 0; q < 8; q++) m[pp][q] = fmaxf(fmaxf(m[pp][q], y[0][pp][q]), y[1][pp][q]);
                    }
#pragma unroll
                    for (int pp = 0; pp < 2; pp++)
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            const float key = __uint_as_float((__float_as_uint(m[pp][q]) & ~7u) | (unsigned)blk);
                            s2[pp][q] = fmaxf(s2[pp][q], fminf(r[pp][q], key));
                            r[pp][q] = fmaxf(r[pp][q], key);
                        }
                }
            }
            // ---------------- pass 2: exact f64 scan of the winning block (or of everything) ----------------
            double best[2][8];
            {
                const double *d64 = sD64 + (size_t)cur * NS * PD;
#pragma unroll
                for (int pp = 0; pp < 2; pp++) {
                    const double *drow = d64 + (size_t)(lane + 32 * pp) * PD;
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        double bv = neg_inf();
                        if (q < nreal) {
                            const float rr = r[pp][q];
                            const float thr = fmaf(rr, 1.0f + 3.814697265625e-06f, -7.5231638452626401e-37f);   // r (1 + 2^-18) - 2^-120
                            const bool amb = !(s2[pp][q] < thr) || rr < -1e29f;     // runner-up too close, or a clamped winner
                            const double *arow = sA64T + (size_t)(c0 + q) * Kp;
                            if (!amb) {
                                const int j0 = (int)(__float_as_uint(rr) & 7u) << 3;
#pragma unroll
                                for (int jj = 0; jj < 8; jj += 2) {
                                    const double2 d = *reinterpret_cast<const double2 *>(drow + j0 + jj);
                                    const double2 a = *reinterpret_cast<const double2 *>(arow + j0 + jj);
                                    const double v0 = d.x + a.x, v1 = d.y + a.y;
                                    bv = v0 > bv ? v0 : bv;
                                    bv = v1 > bv ? v1 : bv;
                                }
                            } else {
                                for (int j = 0; j < Kp; j += 2) {
                                    const double2 d = *reinterpret_cast<const double2 *>(drow + j);
                                    const double2 a = *reinterpret_cast<const double2 *>(arow + j);
                                    const double v0 = d.x + a.x, v1 = d.y + a.y;
                                    bv = v0 > bv ? v0 : bv;
                                    bv = v1 > bv ? v1 : bv;
                                }
                            }
                        }
                        best[pp][q] = bv;
                    }
                }
            }
            // ---------------- epilogue: (delta + a) + b in place (viterbi.rs:17), f64 and f32 copies ----------------
            mbar_wait(sBar + 1, em_phase);
            em_phase ^= 1;
            {
                double *d64n = sD64 + (size_t)nxt * NS * PD;
                float *d32n = sD32 + (size_t)nxt * Kp * NS + 2 * lane;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    if (q < nreal) {
                        const int st = srow0 + q;
                        double *e0 = d64n + (size_t)lane * PD + st, *e1 = d64n + (size_t)(lane + 32) * PD + st;
                        const double v0 = best[0][q] + *e0, v1 = best[1][q] + *e1;
                        *e0 = v0; *e1 = v1;
                        *reinterpret_cast<float2 *>(d32n + (size_t)st * NS) =
                            make_float2(fmaxf(__double2float_rn(v0), PF_NEG), fmaxf(__double2float_rn(v1), PF_NEG));
                    }
                }
            }
            fence_proxy_async_smem();
            if (tid == 0) tma_store_wait_read_all();
            __syncthreads();
            if (tid == 0) tma_bulk_s2g(slab + (size_t)t * NS * PD, sD64 + (size_t)nxt * NS * PD, slab_bytes);
            if (t + 1 < Tmax) {
                issue_emissions(t + 1, o_nxt);
#pragma unroll
                for (int k = 0; k < EMK; k++) {
                    const int s = g + nw * (lane + 32 * k);
                    o_nxt[k] = (s < NS && t + 2 < sLen[s]) ? ld_obs(sOff[s] + t + 2) : 0u;
                }
            }
        }
        if (tid == 0) {
            if (p.tile_done) {
                tma_store_wait_all();
                asm volatile("fence.proxy.async;" ::: "memory");
                __threadfence();
                const unsigned int slot = atomicAdd(p.started + 1, 1u);
                asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.tile_done + slot), "r"(tile + 1) : "memory");
            } else {
                tma_store_wait_read_all();
            }
        }
        __syncthreads();
    }
}

}  // namespace cvb
