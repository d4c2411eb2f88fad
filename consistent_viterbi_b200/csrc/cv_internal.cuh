// cv_internal.cuh -- what the translation units of libcv_b200.so share: the model handle, device buffers, error
// reporting and the tuning block.  Not part of the ABI (include/cv_b200.h is; include/cv_b200_debug.h declares the
// test / bench hooks).
#pragma once

#include "../../include/cv_b200.h"
#include "../../include/cv_b200_debug.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace cvb {

// ---- errors, counters --------------------------------------------------------------------------------------------
int fail(int code, const char *fmt, ...);          // records the calling thread's error text, returns `code`
extern std::atomic<uint64_t> g_launches;           // kernels launched by this library (cv_launch_count)
extern std::atomic<int> g_timing;                  // cv_set_timing

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            int _c = (_e == cudaErrorMemoryAllocation) ? CV_ERR_OOM : CV_ERR_CUDA;             \
            return ::cvb::fail(_c, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                      \
    } while (0)

int check_device(int device);

// ---- tuning block ---------------------------------------------------------------------------------------------------
// Defaults are what the library ships with.  The environment is read ONCE, when the library is loaded (never on a
// launch path); tests and bench.py change single fields through include/cv_b200_debug.h.
struct Tuning {
    int small_cfg = -1;            // small-K launch shape 10*S + MINB, -1 = automatic          (cv_debug_set_small_config)
    int chunks = -1;               // chunks a host batch is cut into, -1 = automatic             (cv_debug_set_chunks, CV_CHUNKS)
    long long chain_max_b = -1;    // batches up to this size use the warp-per-sequence kernel, -1 = 8192
    int bt_concurrent = 1;         // backtrace next to the forward kernel                        (CV_BT_CONCURRENT)
    int streamed = 1;              // host copies streamed past one launch                        (CV_STREAMED)
    int tq = 8;                    // target states per warp of the forward tile kernel: 6, 8, 12, 0 = auto (CV_TQ)
    int tp = 2;                    // sequences per lane: 2 or 4                                  (CV_TP)
    int balanced_split = 1;        // forward tile kernel: state groups of near-equal size, no padded states (CV_BALANCED)
    int fwd_ldc = 1;               // forward tile kernel with compile-time row pitches (tiles of 64, logA pitch 64) when it costs no occupancy (CV_FWD_LDC)
    int em_light = 1;              // uneven balanced split: only the lighter state groups fetch the emission rows (CV_EM_LIGHT)
    int bt_split = 1;              // backtrace with four lanes per sequence (backtrace_split_kernel) where it applies (CV_BT_SPLIT)
    int uneven_chunks = 1;         // streamed host path, automatic chunking: 10/25/25/20/12/8 % with the last chunk's long sequences promoted (CV_UNEVEN_CHUNKS)
    int long_pct = 0;              // long-sequence split: limit in % of the tile-steps per resident CTA, 0 = automatic (70 / 45) (CV_LONG_PCT)
    int prefilter = 0;             // forward tile kernel with the f32 pre-filter (decode_prefilter.cuh) when the model allows it (CV_PREFILTER)
    int long_split = 1;            // very long sequences of a short batch go to the warp-per-sequence kernel (CV_LONG_SPLIT)
    int debug = 0;                 // print launch shapes                                         (CV_DEBUG)
    int bt_prof = 0, e2e_prof = 0; // print pipeline timelines                                    (CV_BT_PROF, CV_E2E_PROF)
    long long large_group_rb = 0;  // row blocks per group of the large-K kernel, 0 = automatic    (CV_LARGE_GROUP_RB)
    int cp_fullwarp = 0, cp_prof = 0, cp_sum = -1, cp_hostpoll = 1, cp_leaf_batch = 1;   // constrained solver (CV_CP_*)
    int probe_threads = 384;       // CV_PROBE_THREADS
};
extern Tuning g_tune;

// ---- device buffer that only grows ---------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need)
    {
        if (need <= bytes) return CV_OK;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        size_t want = need + need / 8;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, need);
            want = need;
        }
        if (e != cudaSuccess) { p = nullptr; return fail(CV_ERR_OOM, "cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e)); }
        bytes = want;
        return CV_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

}  // namespace cvb

// ---- model handle (opaque in the ABI) ------------------------------------------------------------------------------
struct cv_hmm {
    int device = 0, K = 0, Kp = 0, G = 0, D = 0, num_sms = 0, TQT = 8;
    int64_t M = 1;
    bool nonpositive = false;   // every finite model entry is <= 0 (true for log-probabilities)
    // device model
    double *dA = nullptr;    // [K][Kp]   small-K layout (Kp = 8*ceil(K/8)), pad = -inf
    double *dBT = nullptr;   // [M][Kp]
    double *dPi = nullptr;   // [Kp]
    // slot-permuted copies for the balanced state split of the forward tile kernel (TQT = 8, K % 8 != 0), else null
    double *dAb = nullptr, *dBTb = nullptr;
    // pre-filter forward kernel (decode_prefilter.cuh): f32 logA with slot-permuted columns, f64 logA transposed per slot
    float *dA32 = nullptr; double *dA64T = nullptr;
    // f32 mode (decode_f32.cuh): rn32(logA) natural [K][Kp] and slot-permuted [K][Kp], rn32(logB^T) slot-permuted [M][Kp]
    float *dA32n = nullptr, *dA32f = nullptr, *dBT32 = nullptr;
    int f32mode = 0;                          // the call in progress is cv_decode_batch_f32 / cv_decode_batch_dev_f32
    // large-K layout
    int Kl = 0;              // K padded to a multiple of LARGE_BN
    double *dAl = nullptr;   // [Kl][Kl]
    double *dATl = nullptr;  // [Kl][Kl] transposed (lazy-psi backtrace)
    double *dBTl = nullptr;  // [M][Kl]
    // host copy (control logic of the CP solver)
    std::vector<double> hA, hB, hPi;
    // workspaces
    cvb::DevBuf obs, seq_off, path, score;   // device copies of the host-API buffers
    // decode workspaces: two sets so that consecutive chunks of a batch overlap (forward of chunk k+1 with the
    // backtrace / copies of chunk k) on two internal streams
    struct DecodeWs {
        cvb::DevBuf order, keys_in, keys_out, vals_in, cub_tmp, hist, tmax, base, misc, lg_arr, lg_start, lg_done, delta_g;
        cvb::DevBuf is_long, long_list, long_psi;          // long-sequence split of the tile path (cv_api.cu)
        cudaStream_t st = nullptr;
        cudaEvent_t done = nullptr;
        cudaStream_t st_bt = nullptr;                      // concurrent backtrace (launch_decode_small)
        cudaEvent_t ev_pre = nullptr, ev_bt = nullptr;
    } ws[2];
    cudaEvent_t ev_fork = nullptr;
    cudaStream_t st_long = nullptr;                        // warp-per-sequence kernel for the long sequences of a tile launch
    cudaEvent_t ev_long_dep = nullptr, ev_long_done = nullptr, ev_all_in = nullptr;
    cvb::DevBuf cpb[24];     // constrained-solver state (kept after cv_cp_solve for the parity hooks)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    double last_ms = 0.0, last_bt_ms = 0.0;   // forward kernel / backtrace kernel
    int obs16 = 0, path8 = 0;                 // host formats of the call in progress (cv_decode_batch_u16u8), else 0
    // CP debug state
    int64_t cp_N = 0;
    std::vector<double> cp_ub;
    void *pinned_status = nullptr;
};
