// cv_api.cu -- C ABI (include/cv_b200.h) over the sm_100a kernels.
//
// Host side of the drop-in boundary: model upload, workspace management, launch
// logic and the branch-and-bound control loop of the constrained solver.  No
// CPU compute fallback exists anywhere in this file: if CUDA is unavailable the
// entry points return CV_ERR_CUDA.
#include "cv_internal.cuh"

#include <chrono>

#include <cub/cub.cuh>

#include "common.cuh"
#include "decode_small.cuh"
#include "decode_prefilter.cuh"
#include "decode_f32.cuh"
#include "decode_chain.cuh"
#include "decode_large.cuh"

using namespace cvb;

// ---------------------------------------------------------------------------
// errors, counters, tuning block
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
std::atomic<uint64_t> cvb::g_launches{0};
std::atomic<int> cvb::g_timing{0};

int cvb::fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// The environment is read once, here, when the library is loaded.
static Tuning tuning_from_env()
{
    Tuning t;
    auto geti = [](const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; };
    t.chunks = geti("CV_CHUNKS", t.chunks);
    t.small_cfg = geti("CV_SMALL_CFG", t.small_cfg);
    t.bt_concurrent = geti("CV_BT_CONCURRENT", t.bt_concurrent);
    t.streamed = geti("CV_STREAMED", t.streamed);
    if (const char *e = getenv("CV_TQ")) t.tq = !strcmp(e, "auto") ? 0 : (atoi(e) == 12 ? 12 : atoi(e) == 6 ? 6 : 8);
    t.tp = geti("CV_TP", 2) == 4 ? 4 : 2;
    t.balanced_split = geti("CV_BALANCED", t.balanced_split);
    t.fwd_ldc = geti("CV_FWD_LDC", t.fwd_ldc);
    t.em_light = geti("CV_EM_LIGHT", t.em_light);
    t.bt_split = geti("CV_BT_SPLIT", t.bt_split);
    t.uneven_chunks = geti("CV_UNEVEN_CHUNKS", t.uneven_chunks);
    t.long_split = geti("CV_LONG_SPLIT", t.long_split);
    t.long_pct = geti("CV_LONG_PCT", t.long_pct);
    t.prefilter = geti("CV_PREFILTER", t.prefilter);
    t.debug = getenv("CV_DEBUG") != nullptr;
    t.bt_prof = getenv("CV_BT_PROF") != nullptr;
    t.e2e_prof = getenv("CV_E2E_PROF") != nullptr;
    if (const char *e = getenv("CV_LARGE_GROUP_RB")) t.large_group_rb = atoll(e);
    t.cp_fullwarp = getenv("CV_CP_FULLWARP") != nullptr;
    t.cp_prof = getenv("CV_CP_PROF") != nullptr;
    t.cp_sum = geti("CV_CP_SUM", t.cp_sum);
    t.cp_hostpoll = geti("CV_CP_HOSTPOLL", t.cp_hostpoll);
    t.cp_leaf_batch = geti("CV_CP_LEAF_BATCH", t.cp_leaf_batch);
    t.probe_threads = geti("CV_PROBE_THREADS", t.probe_threads);
    return t;
}
Tuning cvb::g_tune = tuning_from_env();

extern "C" const char *cv_last_error(void) { return g_err.c_str(); }
extern "C" uint64_t cv_launch_count(void) { return g_launches.load(); }
extern "C" void cv_set_timing(int on) { g_timing.store(on); }

int cvb::check_device(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(CV_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device >= n) return fail(CV_ERR_ARG, "device %d out of range (%d devices)", device, n);
    return CV_OK;
}

__global__ void build_layouts_kernel(const double *A, const double *Bm, const double *pi, int K, int64_t M, int Kp,
                                     int rowsA, double *Ap, double *BT, double *Pip)
{
    const int64_t n1 = (int64_t)rowsA * Kp, n2 = M * Kp;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n1 + n2 + Kp; e += stride) {
        if (e < n1) {
            int j = (int)(e / Kp), i = (int)(e % Kp);
            Ap[e] = (i < K && j < K) ? A[(int64_t)j * K + i] : neg_inf();
        } else if (e < n1 + n2) {
            int64_t r = e - n1;
            int64_t o = r / Kp; int i = (int)(r % Kp);
            BT[r] = i < K ? Bm[(int64_t)i * M + o] : neg_inf();
        } else {
            int i = (int)(e - n1 - n2);
            Pip[i] = i < K ? pi[i] : neg_inf();
        }
    }
}

extern "C" void cv_hmm_destroy(cv_hmm *h);

// Slot-permuted copies for the balanced state split: column slot 8g + q of row r <- column (state) first(g) + q of the
// natural layout when q < count(g), else -inf; first(g) = g * base + min(g, rem), count(g) = base + (g < rem).
__global__ void permute_slots_kernel(const double *in, double *out, int64_t rows, int Kp, int base, int rem)
{
    const int64_t n = rows * Kp, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const int64_t r = e / Kp; const int c = (int)(e % Kp), g = c >> 3, q = c & 7;
        const int first = g * base + min(g, rem), cnt = base + (g < rem ? 1 : 0);
        out[e] = q < cnt ? in[r * Kp + first + q] : neg_inf();
    }
}

__global__ void transpose_kernel(const double *in, double *out, int n)
{
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8)
        if (by + r < n && bx + threadIdx.x < n) tile[r][threadIdx.x] = in[(size_t)(by + r) * n + bx + threadIdx.x];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8)
        if (bx + r < n && by + threadIdx.x < n) out[(size_t)(bx + r) * n + by + threadIdx.x] = tile[threadIdx.x][r];
}

// f32 copies for the optional f32 mode (decode_f32.cuh): out[r][c] = rn32(in[r][state(c)]) with the slot permutation of
// the balanced split (base = 0: identity), -inf padding.
__global__ void build_f32_kernel(const double *in, float *out, int64_t rows, int K, int Kp, int base, int rem)
{
    const int64_t n = rows * Kp, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const int64_t r = e / Kp; const int c = (int)(e % Kp), g = c >> 3, q = c & 7;
        int st = c;
        if (base) { const int first = g * base + min(g, rem), cnt = base + (g < rem ? 1 : 0); st = q < cnt ? first + q : -1; }
        out[e] = (st >= 0 && st < K) ? __double2float_rn(in[r * Kp + st]) : __int_as_float(0xff800000);
    }
}

// Operands of the pre-filter forward kernel (decode_prefilter.cuh) from the natural layout A [K][Kp]:
//   A32 [Kp][Kp] f32: row = predecessor j, column = slot c -> rn32(logA[j][state(c)]), clamped to PF_NEG; padding PF_NEG
//   A64T [Kp][Kp] f64: row = slot c, column = predecessor j -> logA[j][state(c)]; padding -inf
// state(c) = first(g) + q for slot c = 8 g + q of the balanced split (base, rem), or c itself when base = 0.
__global__ void build_prefilter_kernel(const double *A, int K, int Kp, int base, int rem, float *A32, double *A64T)
{
    const int n = Kp * Kp;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int j = e / Kp, c = e % Kp, g = c >> 3, q = c & 7;
        int st = c;
        if (base) { const int first = g * base + min(g, rem), cnt = base + (g < rem ? 1 : 0); st = q < cnt ? first + q : -1; }
        const bool real = j < K && st >= 0 && st < K;
        const double v = real ? A[(size_t)j * Kp + st] : neg_inf();
        A32[(size_t)j * Kp + c] = fmaxf(__double2float_rn(v), PF_NEG);
        A64T[(size_t)c * Kp + j] = v;
    }
}

extern "C" int cv_hmm_create(int K, int D, const uint64_t *bdims, const double *logA, const double *logB,
                             const double *logPi, int device, cv_hmm **out)
{
    if (!out) return fail(CV_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (K <= 0 || K > 65535) return fail(CV_ERR_ARG, "K=%d out of range [1,65535]", K);
    if (D <= 0 || !bdims || !logA || !logB || !logPi) return fail(CV_ERR_ARG, "NULL/empty model argument");
    int64_t M = 1;
    for (int d = 0; d < D; d++) {
        if (bdims[d] == 0) return fail(CV_ERR_ARG, "bdims[%d] == 0", d);
        M *= (int64_t)bdims[d];
    }
    // NaN / +inf would make the reference's argmax().unwrap() panic (or poison every sum): reject.
    auto bad = [](const double *v, int64_t n) {
        for (int64_t i = 0; i < n; i++) if (std::isnan(v[i]) || v[i] == std::numeric_limits<double>::infinity()) return true;
        return false;
    };
    if (bad(logA, (int64_t)K * K) || bad(logB, (int64_t)K * M) || bad(logPi, K))
        return fail(CV_ERR_NAN, "model contains NaN or +inf");
    if (device < 0) {
        int rc = check_device(0);
        if (rc) return rc;
        CUDA_TRY(cudaGetDevice(&device));
    }
    int rc = check_device(device);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));

    cv_hmm *h = new cv_hmm();
    struct Guard { cv_hmm *h; ~Guard() { if (h) cv_hmm_destroy(h); } } guard{h};   // released on every early return
    h->device = device; h->K = K; h->D = D; h->M = M;
    // tile shape of the forward kernel: TQT target states per warp, G = ceil(K/TQT) state groups.  Pick the
    // shape with the least padded work among those whose warps per CTA (G*S, S in {1,2,4}) divide by 4.
    {
        double best_eff = -1.0; int best_tq = 8;
        for (int tq : {8, 12}) {
            const int g = (K + tq - 1) / tq;
            double eff = 0.0;
            for (int s : {1, 2, 4}) {
                if (g * s * 32 > 512) continue;
                const int w = g * s;
                const double bal = (double)w / (4.0 * ((w + 3) / 4));
                // wide CTAs (one per SM) lose to the per-step barrier: mild penalty so that S = 1 wins ties
                eff = std::max(eff, bal * (s == 1 ? 1.0 : 0.93) * (double)K / (g * tq));
            }
            if (eff > best_eff + 1e-9) { best_eff = eff; best_tq = tq; }
        }
        // measured on B200 (POS shape): TQ = 8 with two 6-warp CTAs per SM beats the balanced TQ = 12 shape
        // (12 vs 8 resident warps), so 8 stays the default; CV_TQ=12 / 6 / auto select the alternatives
        h->TQT = g_tune.tq == 0 ? best_tq : g_tune.tq;
    }
    h->G = (K + h->TQT - 1) / h->TQT;
    h->Kp = ((std::max(h->G * h->TQT, K) + 7) / 8) * 8;
    h->hA.assign(logA, logA + (size_t)K * K);
    h->hB.assign(logB, logB + (size_t)K * M);
    h->hPi.assign(logPi, logPi + K);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&h->ev0));
    CUDA_TRY(cudaEventCreate(&h->ev1));
    CUDA_TRY(cudaEventCreate(&h->ev2));
    CUDA_TRY(cudaMallocHost(&h->pinned_status, 2048));     // [0, 64): status / bound words, [1024, 2048): leaf-batch bounds
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->st_long, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_long_dep, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_long_done, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_all_in, cudaEventDisableTiming));
    for (auto &w : h->ws) {
        CUDA_TRY(cudaStreamCreateWithFlags(&w.st, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
        CUDA_TRY(cudaStreamCreateWithFlags(&w.st_bt, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&w.ev_pre, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&w.ev_bt, cudaEventDisableTiming));
    }

    const int Kp = (K <= SMALL_K_MAX) ? h->Kp : ((K + LARGE_BN - 1) / LARGE_BN) * LARGE_BN;
    if (K > SMALL_K_MAX) { h->Kl = Kp; }
    double *tA = nullptr, *tB = nullptr, *tPi = nullptr;
    double *pA = nullptr, *pBT = nullptr, *pPi = nullptr;
    // staging + layout buffers are freed on every early return (the layout buffers are handed to the handle at the end)
    struct Staging { double **p[6]; ~Staging() { for (double **q : p) if (*q) { cudaFree(*q); *q = nullptr; } } } staging{{&tA, &tB, &tPi, &pA, &pBT, &pPi}};
    CUDA_TRY(cudaMalloc(&tA, sizeof(double) * (size_t)K * K));
    CUDA_TRY(cudaMalloc(&tB, sizeof(double) * (size_t)K * M));
    CUDA_TRY(cudaMalloc(&tPi, sizeof(double) * (size_t)K));
    CUDA_TRY(cudaMemcpy(tA, logA, sizeof(double) * (size_t)K * K, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(tB, logB, sizeof(double) * (size_t)K * M, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(tPi, logPi, sizeof(double) * (size_t)K, cudaMemcpyHostToDevice));
    const int rowsA = (K <= SMALL_K_MAX) ? K : Kp;   // the large-K TMA ring reads whole 16-row chunks
    CUDA_TRY(cudaMalloc(&pA, sizeof(double) * (size_t)rowsA * Kp));
    CUDA_TRY(cudaMalloc(&pBT, sizeof(double) * (size_t)M * Kp));
    CUDA_TRY(cudaMalloc(&pPi, sizeof(double) * (size_t)Kp));
    build_layouts_kernel<<<std::max(1, h->num_sms) * 4, 256>>>(tA, tB, tPi, K, M, Kp, rowsA, pA, pBT, pPi);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    if (K <= SMALL_K_MAX) { h->dA = pA; h->dBT = pBT; h->dPi = pPi; }
    else { h->dAl = pA; h->dBTl = pBT; h->dPi = pPi; }
    pA = pBT = pPi = nullptr;                      // owned by the handle now (cv_hmm_destroy)
    if (K <= SMALL_K_MAX && h->TQT == 8 && K % 8 != 0 && K / h->G >= 5) {
        // balanced state split of the forward tile kernel: no padded target states (K = 45: 8,8,8,7,7,7 instead of 5 x 8 + 5)
        CUDA_TRY(cudaMalloc(&h->dAb, sizeof(double) * (size_t)K * Kp));
        CUDA_TRY(cudaMalloc(&h->dBTb, sizeof(double) * (size_t)M * Kp));
        permute_slots_kernel<<<std::max(1, h->num_sms) * 4, 256>>>(h->dA, h->dAb, K, Kp, K / h->G, K % h->G);
        permute_slots_kernel<<<std::max(1, h->num_sms) * 4, 256>>>(h->dBT, h->dBTb, M, Kp, K / h->G, K % h->G);
        g_launches += 2;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
    }
    if (K <= SMALL_K_MAX && h->TQT == 8 && Kp % 8 == 0) {
        // optional f32 mode (cv_decode_batch_f32): f32 copies of the model
        const bool bal = h->dAb != nullptr;
        const int base = bal ? K / h->G : 0, rem = bal ? K % h->G : 0;
        CUDA_TRY(cudaMalloc(&h->dA32n, sizeof(float) * (size_t)K * Kp));
        CUDA_TRY(cudaMalloc(&h->dA32f, sizeof(float) * (size_t)K * Kp));
        CUDA_TRY(cudaMalloc(&h->dBT32, sizeof(float) * (size_t)M * Kp));
        build_f32_kernel<<<32, 256>>>(h->dA, h->dA32n, K, K, Kp, 0, 0);
        build_f32_kernel<<<32, 256>>>(h->dA, h->dA32f, K, K, Kp, base, rem);
        build_f32_kernel<<<std::max(1, h->num_sms) * 4, 256>>>(h->dBT, h->dBT32, M, K, Kp, base, rem);
        g_launches += 3;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
    }
    // every finite entry of logA / logB <= 0 (log-probabilities): the f32 pre-filter's error bound needs it
    h->nonpositive = true;
    for (int64_t i = 0; i < (int64_t)K * K && h->nonpositive; i++) if (logA[i] > 0.0) h->nonpositive = false;
    for (int64_t i = 0; i < (int64_t)K * M && h->nonpositive; i++) if (logB[i] > 0.0) h->nonpositive = false;
    if (K <= SMALL_K_MAX && h->TQT == 8 && h->nonpositive && K >= 17 && Kp % 8 == 0) {
        CUDA_TRY(cudaMalloc(&h->dA32, sizeof(float) * (size_t)Kp * Kp));
        CUDA_TRY(cudaMalloc(&h->dA64T, sizeof(double) * (size_t)Kp * Kp));
        const bool bal = h->dAb != nullptr;
        build_prefilter_kernel<<<32, 256>>>(h->dA, K, Kp, bal ? K / h->G : 0, bal ? K % h->G : 0, h->dA32, h->dA64T);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
    }
    if (K > SMALL_K_MAX) {
        CUDA_TRY(cudaMalloc(&h->dATl, sizeof(double) * (size_t)Kp * Kp));
        transpose_kernel<<<dim3((Kp + 31) / 32, (Kp + 31) / 32), dim3(32, 8)>>>(h->dAl, h->dATl, Kp);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
    }
    guard.h = nullptr;
    *out = h;
    return CV_OK;
}

extern "C" void cv_hmm_destroy(cv_hmm *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (double *p : {h->dA, h->dBT, h->dPi, h->dAl, h->dBTl, h->dATl, h->dAb, h->dBTb, h->dA64T}) if (p) cudaFree(p);
    for (float *p : {h->dA32, h->dA32n, h->dA32f, h->dBT32}) if (p) cudaFree(p);
    for (DevBuf *b : {&h->obs, &h->seq_off, &h->path, &h->score}) b->release();
    for (auto &w : h->ws) {
        for (DevBuf *b : {&w.order, &w.keys_in, &w.keys_out, &w.vals_in, &w.cub_tmp, &w.hist, &w.tmax, &w.base, &w.misc,
                          &w.lg_arr, &w.lg_start, &w.lg_done, &w.delta_g, &w.is_long, &w.long_list, &w.long_psi})
            b->release();
        if (w.st) cudaStreamDestroy(w.st);
        if (w.done) cudaEventDestroy(w.done);
        if (w.st_bt) cudaStreamDestroy(w.st_bt);
        if (w.ev_pre) cudaEventDestroy(w.ev_pre);
        if (w.ev_bt) cudaEventDestroy(w.ev_bt);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->st_long) cudaStreamDestroy(h->st_long);
    for (cudaEvent_t e : {h->ev_long_dep, h->ev_long_done, h->ev_all_in}) if (e) cudaEventDestroy(e);
    for (auto &b : h->cpb) b.release();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev2) cudaEventDestroy(h->ev2);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->pinned_status) cudaFreeHost(h->pinned_status);
    delete h;
}

extern "C" int cv_hmm_nstates(const cv_hmm *h) { return h ? h->K : 0; }
extern "C" int64_t cv_hmm_nobs(const cv_hmm *h) { return h ? h->M : 0; }
extern "C" double cv_last_kernel_ms(const cv_hmm *h) { return h ? h->last_ms : 0.0; }
extern "C" double cv_last_backtrace_ms(const cv_hmm *h) { return h ? h->last_bt_ms : 0.0; }

extern "C" void *cv_host_alloc(uint64_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void cv_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---------------------------------------------------------------------------
// batched plain Viterbi
// ---------------------------------------------------------------------------
// Long-sequence split (launch_decode_small): the tile kernel runs a tile in lock step for as many steps as its longest
// sequence has, ~8 us per step, so in a SHORT batch (a slice of a sharded batch) a handful of very long sequences
// are the critical path of the whole launch.  Sequences longer than `lstar` are taken out of the tiles: they get
// sort key 0 (they sit in a tile of the shortest sequences as inactive slots), are flagged in is_long[] and
// appended to long_list, and the warp-per-sequence kernel decodes them next to the tile kernel (0.3-0.5 us per
// step).  lstar = 0: no split.  At most cap_long sequences are moved (the rest stay in the tiles: still correct).
struct LongSplit {
    uint32_t lstar, cap_long;
    uint8_t *is_long;            // [B], zeroed by the host
    uint32_t *long_list;         // [cap_long]
    unsigned int *n_long;        // appended so far (may exceed cap_long)
};
__device__ __forceinline__ bool take_long(const LongSplit &ls, int64_t b, int64_t len)
{
    if (ls.lstar == 0 || len <= (int64_t)ls.lstar) return false;
    const unsigned int idx = atomicAdd(ls.n_long, 1u);
    if (idx >= ls.cap_long) return false;
    ls.long_list[idx] = (uint32_t)b;
    ls.is_long[b] = 1;
    return true;
}

// max_len > 0: the caller's bound on the lengths -- the sort then looks only at the bits a length can have, so a longer
// sequence must be an error (CV_ERR_ARG), not a mis-sorted batch
__global__ void seq_len_keys_kernel(const int64_t *seq_off, int64_t B, uint32_t *keys, uint32_t *vals, int *status, const LongSplit ls,
                                    int64_t max_len)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int64_t len = seq_off[b + 1] - seq_off[b];
    if (len <= 0) *status = CV_ERR_EMPTY;          // reference: sequence.len()-1 underflow panic
    else if (max_len > 0 && len > max_len) *status = CV_ERR_ARG;
    const bool lng = take_long(ls, b, len);
    keys[b] = lng ? 0u : (uint32_t)(len < 0 ? 0 : (len > 0xffffffffLL ? 0xffffffffLL : len));
    vals[b] = (uint32_t)b;
}

// bits of the sort key that a length <= max_len can set
static int key_bits(int64_t max_len)
{
    int b = 1;
    while (b < 32 && (max_len >> b) != 0) b++;
    return b;
}

// longest length of every tile of NS sequences (lengths are sorted descending)
__global__ void tile_tmax_kernel(const uint32_t *sorted_len, int ntiles, int NS, long long *tmax)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntiles) tmax[t] = (long long)sorted_len[(size_t)t * NS];
}

// ---- streamed host path: keys = (chunk descending-coded, length) so that one descending sort orders the batch by
// chunk (ascending) and, inside a chunk, by length (longest first); lengths must fit 24 bits
constexpr int STREAM_MAX_CHUNKS = 16;
// key = ordering group << shift | length.  promote_len > 0: the sequences of the chunks behind promote_to that are longer
// than this are ordered with chunk promote_to (all observations have long arrived when that group is reached): every
// group starts with its longest tiles, and a late, small group must not start tiles that outlast everything else.
struct ChunkBounds { int nch; int shift; int promote_len, promote_to; int64_t cb[STREAM_MAX_CHUNKS + 2]; };

__global__ void seq_chunk_keys_kernel(const int64_t *seq_off, int64_t B, const ChunkBounds cbs, uint32_t *keys, uint32_t *vals,
                                      int *status, unsigned int *max_len, const LongSplit ls)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int64_t len = seq_off[b + 1] - seq_off[b];
    if (len == 0) atomicMax(status, CV_ERR_EMPTY);            // reference: sequence.len()-1 underflow panic
    if (len < 0) atomicMax(status, CV_ERR_ARG);               // offsets not monotone
    const uint32_t l = (uint32_t)(len < 0 ? 0 : (len > 0xffffffLL ? 0xffffffLL : len));
    if (len > 0xffffffLL) atomicMax(status, CV_ERR_UNSUPPORTED);
    atomicMax(max_len, l);
    int c = 0;
    while (c + 1 < cbs.nch && b >= cbs.cb[c + 1]) c++;
    if (cbs.promote_len > 0 && c > cbs.promote_to && len > (int64_t)cbs.promote_len) c = cbs.promote_to;
    keys[b] = ((uint32_t)(cbs.nch - 1 - c) << cbs.shift) | (take_long(ls, b, len) ? 0u : l);
    vals[b] = (uint32_t)b;
}

// per tile of NS sorted sequences: the longest length and the last chunk it takes sequences from
// promote_to >= 0: tiles of that ordering group may hold sequences of any later chunk (ChunkBounds::promote_len)
__global__ void tile_meta_kernel(const uint32_t *sorted_keys, int ntiles, int NS, int64_t B, int nch, int shift, int promote_to, long long *tmax, int *tchunk)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    uint32_t lmax = 0; int cmax = 0;
    for (int s = 0; s < NS; s++) {
        const int64_t r = (int64_t)t * NS + s;
        if (r >= B) break;
        const uint32_t k = sorted_keys[r];
        lmax = max(lmax, k & ((1u << shift) - 1u));
        cmax = max(cmax, nch - 1 - (int)(k >> shift));
    }
    if (promote_to >= 0 && cmax == promote_to) cmax = nch - 1;
    tmax[t] = (long long)lmax; tchunk[t] = cmax;
}

// what launch_decode_small needs to know about the streamed host path
struct StreamedIO {
    ChunkBounds cbs;
    unsigned int *d_arrived;        // chunks whose observations have arrived
    unsigned int *d_chunk_done;     // [nch] finished sequences per chunk
    cudaEvent_t all_in;             // recorded on the copy stream behind the last chunk of observations
};

// Threshold of the long-sequence split (see LongSplit): a tile should not run longer than ~70 % of the tile-steps an
// average resident CTA executes in this launch; never below 64 steps; 0 = no split (nothing is that long).
// Batches above the four-lane backtrace's range (launch_decode_small) are backtraced by one thread per sequence at
// 6-9 us per step -- about as long again as the sequence's forward pass in a tile -- so there the limit is 45 %: at
// 250 k sentences (a rank of a 4-GPU run) the slice holding a 191-step sentence ended 0.18 ms after the others.
static uint32_t long_threshold(const cv_hmm *h, int64_t B, int64_t N, int64_t max_len)
{
    if (!g_tune.long_split || max_len <= 0 || h->K > SMALL_K_MAX || h->f32mode) return 0;   // (the warp-per-sequence kernel is f64)
    const int64_t per_cta = (N / 64) / std::max<int64_t>(1, (int64_t)h->num_sms * 2);
    const int pct = g_tune.long_pct > 0 ? g_tune.long_pct : (g_tune.bt_split && B <= (int64_t)1300 * h->num_sms) ? 70 : 45;
    const int64_t l = std::max<int64_t>(64, per_cta * pct / 100);
    return l >= max_len ? 0u : (uint32_t)std::min<int64_t>(l, 0xffffff);
}
constexpr uint32_t LONG_CAP = 4096;

// binds the split's buffers for one launch (workspace set w) and zeroes the flags; ls.lstar = 0 when there is no split
static int long_split_setup(cv_hmm *h, cv_hmm::DecodeWs &w, int64_t B, int64_t N, int64_t max_len, cudaStream_t st, LongSplit &ls)
{
    ls = LongSplit{0u, 0u, nullptr, nullptr, nullptr};
    const uint32_t lstar = long_threshold(h, B, N, max_len);
    if (!lstar) return CV_OK;
    int rc;
    const size_t row = (size_t)max_len * h->Kp;
    const uint32_t cap = (uint32_t)std::max<size_t>(1, std::min<size_t>(LONG_CAP, ((size_t)512 << 20) / std::max<size_t>(row, 1)));
    if ((rc = w.is_long.ensure((size_t)B))) return rc;
    if ((rc = w.long_list.ensure(sizeof(uint32_t) * cap))) return rc;
    if ((rc = w.long_psi.ensure(row * cap + 64))) return rc;
    CUDA_TRY(cudaMemsetAsync(w.is_long.p, 0, (size_t)B, st));
    ls.lstar = lstar; ls.cap_long = cap;
    ls.is_long = (uint8_t *)w.is_long.p; ls.long_list = (uint32_t *)w.long_list.p;
    ls.n_long = (unsigned int *)w.misc.p + 9;                // misc is zeroed by the caller before the keys kernel
    return CV_OK;
}

typedef cv_hmm::DecodeWs DecodeWs;
#include "decode_large_host.inl"

// cuStreamWaitValue32 through the runtime's driver entry-point lookup (no link-time dependency on libcuda):
// lets a stream wait until a word in device memory reaches a value.  nullptr when unavailable.
typedef int (*StreamWaitValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
static StreamWaitValue32Fn stream_wait_value32()
{
    static StreamWaitValue32Fn fn = []() -> StreamWaitValue32Fn {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (StreamWaitValue32Fn)f;
    }();
    return fn;
}


// the warp-per-sequence kernel with the instantiation that fits K (register-resident logA column for K <= 32)
static int launch_chain_kernel(const cv_hmm *h, const DecodeChainParams &p, int grid, cudaStream_t st)
{
    const size_t smem = (size_t)h->K * h->Kp * 8 + (p.bt_in_smem ? (size_t)h->M * h->Kp * 8 : 0) +
                        (size_t)DC_WARPS * (16 * h->Kp + 32 * h->Kp + 128);
    auto go = [&](auto kern) -> int {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 32 * DC_WARPS, smem, st>>>(p);
        return CV_OK;
    };
    const bool regs = h->Kp == 8 * ((h->K + 7) / 8);     // register-resident logA column needs Kp = 4 * KQ
    int rc;
    switch (regs ? (h->K + 7) / 8 : 9) {
        case 1: rc = go(decode_chain_kernel<1, 2>); break;
        case 2: rc = go(decode_chain_kernel<1, 4>); break;
        case 3: rc = go(decode_chain_kernel<1, 6>); break;
        case 4: rc = go(decode_chain_kernel<1, 8>); break;
        default: rc = h->K <= 32 ? go(decode_chain_kernel<1, 0>) : go(decode_chain_kernel<2, 0>);
    }
    if (rc) return rc;
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return CV_OK;
}

static int launch_decode_small(cv_hmm *h, DecodeWs &w, const uint32_t *d_obs, const int64_t *d_off, int64_t B,
                               int64_t N, uint32_t *d_path, double *d_score, unsigned int *d_counter, int *d_status,
                               int64_t max_len, cudaStream_t st, bool timing, const StreamedIO *sio = nullptr,
                               const LongSplit *ls = nullptr)
{
    const int G = h->G;
    const uint32_t *d_order = (const uint32_t *)w.order.p, *d_sorted_len = (const uint32_t *)w.keys_out.p;
    // Launch shape: S sequence groups of 64 per CTA, and which register budget (kernel instantiation) to use.
    int S = 1, variant = 1;
    if (g_tune.small_cfg >= 0) { S = std::max(1, std::min(4, g_tune.small_cfg / 10)); variant = std::max(1, g_tune.small_cfg % 10); }
    int tpt = g_tune.tp;
    if (h->TQT != 8) tpt = 2;
    // forward kernel with the f32 pre-filter (decode_prefilter.cuh): non-positive models, 64 sequences per tile
    const bool pf = !h->f32mode && g_tune.prefilter && h->dA32 && h->TQT == 8 && g_tune.small_cfg < 0 && tpt == 2 && 32 * G <= 256 &&
                    decode_pf_smem_bytes(h->Kp) <= 113 * 1024;
    // optional f32 mode (decode_f32.cuh): its own forward kernel, f32 history, f32 backtrace
    const bool f32 = h->f32mode && h->dA32f && h->TQT == 8 && g_tune.small_cfg < 0 && tpt == 2 && 32 * G <= 256;
    if (h->f32mode && !f32) return fail(CV_ERR_UNSUPPORTED, "the f32 mode needs the default tile shape (K <= 64)");
    if (pf || f32) S = 1;
    while (S > 1 && (B + 32 * tpt * S - 1) / (32 * tpt * S) < 4 * (int64_t)h->num_sms) S--;
    size_t smem = f32 ? decode_f32_smem_bytes(h->K, h->Kp) : pf ? decode_pf_smem_bytes(h->Kp) : decode_small_smem_bytes(h->K, h->Kp, 32 * tpt * S);
    while (smem > 220 * 1024 && S > 1) { S--; smem = decode_small_smem_bytes(h->K, h->Kp, 32 * tpt * S); }
    if (smem > 220 * 1024 && tpt == 4) { tpt = 2; smem = decode_small_smem_bytes(h->K, h->Kp, 32 * tpt * S); }
    const int NS = 32 * tpt * S;
    const int threads = 32 * G * S;
    if (variant == 3 && threads > 256) variant = 2;
    if (variant == 2 && threads > 384) variant = 1;
    if (threads > 512) return fail(CV_ERR_UNSUPPORTED, "small-K launch shape needs %d threads", threads);
    const int64_t ntiles64 = (B + NS - 1) / NS;
    if (ntiles64 > 0x7fffffffLL) return fail(CV_ERR_UNSUPPORTED, "batch too large");
    const int ntiles = (int)ntiles64;
    int rc;
    if (max_len <= 0) {   // not supplied: read the longest length back (one sync)
        uint32_t L = 0;
        CUDA_TRY(cudaMemcpyAsync(&L, d_sorted_len, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        max_len = L;
    }
    // history slabs: sum over tiles of NS * Tmax(tile) <= N + NS * max_len because lengths are sorted (per chunk, plus
    // the tiles that straddle two chunks, on the streamed path)
    // Streamed path: tiles are ordered by (chunk, length).  Inside a chunk the staircase bound above holds (one
    // NS * max_len term per chunk); a tile that straddles two chunks takes its length from either and adds up to
    // NS * max_len of its own, and there are at most nch - 1 of those: 2 * nch - 1 terms in all.  The forward kernel
    // also checks every tile against hist_cap_rows and reports CV_ERR_UNSUPPORTED instead of storing past the end.
    const size_t stairs = sio ? 2 * (size_t)sio->cbs.nch - 1 : 1;
    const size_t hist_rows = (size_t)N + (size_t)NS * (size_t)max_len * stairs;     // in units of K doubles
    const size_t hist_elems = hist_rows * (size_t)(pf ? pf_pitch(h->Kp) : h->K);    // pre-filter kernel: rows of Kp + 2 doubles
    if ((rc = w.hist.ensure(hist_elems * (f32 ? sizeof(float) : sizeof(double))))) return rc;
    if ((rc = w.tmax.ensure(sizeof(long long) * (size_t)ntiles))) return rc;
    if ((rc = w.base.ensure(sizeof(long long) * (size_t)ntiles))) return rc;
    if (sio) {
        if ((rc = w.delta_g.ensure(sizeof(int) * (size_t)ntiles + 64))) return rc;     // tile -> chunk (buffer unused for K <= 64)
        tile_meta_kernel<<<(ntiles + 127) / 128, 128, 0, st>>>(d_sorted_len, ntiles, NS, B, sio->cbs.nch, sio->cbs.shift, sio->cbs.promote_len > 0 ? sio->cbs.promote_to : -1, (long long *)w.tmax.p, (int *)w.delta_g.p);
    } else {
        tile_tmax_kernel<<<(ntiles + 255) / 256, 256, 0, st>>>(d_sorted_len, ntiles, NS, (long long *)w.tmax.p);
    }
    g_launches++;
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (long long *)w.tmax.p, (long long *)w.base.p, ntiles, st));
    if ((rc = w.cub_tmp.ensure(tmp_bytes))) return rc;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(w.cub_tmp.p, tmp_bytes, (long long *)w.tmax.p, (long long *)w.base.p, ntiles, st));

    DecodeSmallParams p;
    p.A = h->dA; p.BT = h->dBT; p.obs = d_obs; p.seq_off = d_off; p.order = d_order;
    p.At = h->dA; p.BTt = h->dBT; p.nq_base = p.nq_rem = 0;
    p.obs16 = h->obs16; p.path8 = h->path8; p.em_light = g_tune.em_light;
    p.is_long = (ls && ls->lstar) ? ls->is_long : nullptr;
    p.A32s = f32 ? h->dA32f : h->dA32; p.A64Ts = h->dA64T; p.A32n = h->dA32n; p.BT32 = h->dBT32;
    if ((pf || f32 || g_tune.balanced_split) && h->dAb && h->TQT == 8 && tpt == 2) {     // (the pre-filter operands are built for the balanced split)
        p.At = h->dAb; p.BTt = h->dBTb; p.nq_base = h->K / G; p.nq_rem = h->K % G;
    }
    p.tile_base = (const long long *)w.base.p;
    p.hist = (double *)w.hist.p; p.hist_cap_slabs = (long long)(hist_rows / (size_t)NS); p.path = d_path; p.score = d_score; p.tile_counter = d_counter; p.status = d_status;
    p.M = h->M; p.B = B; p.K = h->K; p.Kp = h->Kp; p.G = G; p.S = S; p.NS = NS; p.ntiles = ntiles;
    p.tile_tmax = nullptr; p.tile_chunk = nullptr; p.arrived = nullptr; p.chunk_done = nullptr; p.nch = 0;
    if (sio) {
        p.tile_tmax = (const long long *)w.tmax.p; p.tile_chunk = (const int *)w.delta_g.p;
        p.arrived = sio->d_arrived; p.chunk_done = sio->d_chunk_done; p.nch = sio->cbs.nch;
        for (int c = 0; c <= sio->cbs.nch; c++) p.cb[c] = sio->cbs.cb[c];
    }
    void (*kern)(DecodeSmallParams);
    if (f32) kern = decode_f32_fwd_kernel;
    else if (pf) kern = decode_pf_fwd_kernel;
    else if (h->TQT == 12) kern = variant == 1 ? decode_small_fwd_kernel<12, 512, 1> : decode_small_fwd_kernel<12, 256, 2>;
    else if (h->TQT == 6) kern = variant == 1 ? decode_small_fwd_kernel<6, 512, 1> : decode_small_fwd_kernel<6, 256, 2>;
    else if (tpt == 4 && threads <= 256) kern = decode_small_fwd_kernel<8, 256, 1, 4>;
    else if (tpt == 4) kern = decode_small_fwd_kernel<8, 512, 1, 4>;
    else kern = variant == 1 ? decode_small_fwd_kernel<8, 512, 1>
              : variant == 2 ? decode_small_fwd_kernel<8, 384, 2> : decode_small_fwd_kernel<8, 256, 3>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int occ = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    occ = std::max(1, occ);
    if (!f32 && !pf && h->TQT == 8 && tpt == 2 && S == 1 && variant == 1 && g_tune.fwd_ldc) {
        // Default shape: the same kernel with compile-time row pitches (tiles of 64 sequences, logA rows at a pitch of
        // 64 doubles in shared memory) -- unless the wider logA costs a resident CTA per SM.
        void (*kc)(DecodeSmallParams) = decode_small_fwd_kernel<8, 512, 1, 2, FWD_LDC>;
        const size_t smem_c = decode_small_smem_bytes(h->K, h->Kp, NS, FWD_LDC);
        int occ_c = 0;
        if (smem_c <= 220 * 1024) {
            CUDA_TRY(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
            CUDA_TRY(cudaFuncSetAttribute(kc, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, kc, threads, smem_c));
        }
        if (occ_c >= occ) { kern = kc; smem = smem_c; }
    }
    const int grid = (int)std::min<int64_t>((int64_t)h->num_sms * occ, ntiles);
    if (g_tune.debug)
        fprintf(stderr, "[cv] decode_small: K=%d TQ=%d TP=%d G=%d S=%d variant=%d prefilter=%d threads=%d smem=%zu occ=%d grid=%d tiles=%d\n",
                h->K, h->TQT, tpt, G, S, variant, (int)pf, threads, smem, occ, grid, ntiles);
    // Concurrent backtrace: the backtrace kernel runs next to the forward kernel on a second stream and follows
    // it tile by tile (tile_done flags).  It is released only when every forward CTA is resident (stream wait on
    // the `started` counter), so its spinning CTAs can never keep a forward CTA off an SM; it then lives on the
    // registers / shared memory the forward CTAs leave free.
    StreamWaitValue32Fn wait32 = stream_wait_value32();
    const bool concurrent = !timing && g_tune.bt_concurrent && wait32 != nullptr && w.st_bt != nullptr;
    if (sio && !concurrent) return fail(CV_ERR_UNSUPPORTED, "the streamed path needs the concurrent backtrace");
    p.tile_done = nullptr; p.started = nullptr;
    if (concurrent) {
        if ((rc = w.lg_done.ensure(sizeof(int) * ((size_t)ntiles + 8)))) return rc;
        CUDA_TRY(cudaMemsetAsync(w.lg_done.p, 0, sizeof(int) * ((size_t)ntiles + 8), st));
        p.tile_done = (int *)w.lg_done.p;
        p.started = (unsigned int *)w.lg_done.p + ntiles;
        CUDA_TRY(cudaEventRecord(w.ev_pre, st));
    }
    // Long-sequence split: the flagged sequences run on the warp-per-sequence kernel next to the tile kernels, on a
    // stream of its own; it is launched first so that its few CTAs are resident before the persistent forward CTAs.
    bool long_launched = false;
    if (ls && ls->lstar) {
        DecodeChainParams q;
        q.A = h->dA; q.BT = h->dBT; q.obs = d_obs; q.seq_off = d_off; q.order = ls->long_list;
        q.psi = (uint8_t *)w.long_psi.p; q.path = d_path; q.score = d_score;
        q.counter = (unsigned int *)w.misc.p + 8; q.status = d_status;
        q.M = h->M; q.B = (int64_t)ls->cap_long; q.K = h->K; q.Kp = h->Kp;
        q.bt_in_smem = ((size_t)h->M * h->Kp * 8 <= CHAIN_BT_SMEM_MAX) ? 1 : 0;
        q.obs16 = h->obs16; q.path8 = h->path8;
        q.B_dev = ls->n_long; q.psi_stride = max_len;
        q.chunk_done = sio ? sio->d_chunk_done : nullptr; q.nch = sio ? sio->cbs.nch : 0;
        if (sio) for (int c = 0; c <= sio->cbs.nch; c++) q.cb[c] = sio->cbs.cb[c];
        CUDA_TRY(cudaEventRecord(h->ev_long_dep, st));                        // keys kernel (list, count) and the offsets are done
        CUDA_TRY(cudaStreamWaitEvent(h->st_long, h->ev_long_dep, 0));
        if (sio) CUDA_TRY(cudaStreamWaitEvent(h->st_long, sio->all_in, 0));   // streamed: all observations on the device
        const int grid_l = (int)std::min<int64_t>((ls->cap_long + DC_WARPS - 1) / DC_WARPS, (int64_t)h->num_sms);
        if ((rc = launch_chain_kernel(h, q, grid_l, h->st_long))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_long_done, h->st_long));
        long_launched = true;
    }
    const bool bt_prof = g_tune.bt_prof != 0;   // prints forward / forward+backtrace times of the concurrent mode
    if (timing || bt_prof) CUDA_TRY(cudaEventRecord(h->ev0, st));
    kern<<<std::max(1, grid), threads, smem, st>>>(p);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    if (timing || bt_prof) CUDA_TRY(cudaEventRecord(h->ev1, st));
    // end state + backtrace with lazy backpointers: one thread per sequence
    const size_t smem_bt = (size_t)h->K * (h->K | 1) * 8 + 16 * 8;      // + one chunk of padding behind the last row
    void (*bt_seq)(DecodeSmallParams) = f32 ? backtrace_small_kernel<16, 4, 64, 0, float> : pf ? backtrace_small_kernel<16, 4, 64, 1> : NS == 64 ? backtrace_small_kernel<16, 4, 64> : backtrace_small_kernel<16, 4, 0>;
    void (*bt_con)(DecodeSmallParams) = f32 ? backtrace_small_kernel<8, 8, 64, 0, float> : pf ? backtrace_small_kernel<8, 8, 64, 1> : NS == 64 ? backtrace_small_kernel<8, 8, 64> : backtrace_small_kernel<8, 8, 0>;
    // Four lanes per sequence (backtrace_split_kernel): f64, history layout 0, tiles of 64 sequences -- for batches that do
    // not fill the GPU for long (a rank's slice of a sharded batch): there the one-thread-per-sequence kernel is latency
    // bound and most of its work is still open when the forward kernel ends.  Measured, device step with the concurrent
    // backtrace, split vs. one thread per sequence: 125 k sentences 1.51 vs 1.74 ms, 250 k 2.83 vs 2.79 ms, 500 k 5.59 vs
    // 5.29 ms, 1 M 11.20 vs 10.56 ms (four times the threads re-stage logA four times as often and execute ~15 % more
    // instructions next to an issue-bound forward kernel).  bt_split = 2 forces it for every size.
    bool bt_split = g_tune.bt_split && !f32 && !pf && NS == 64 && (g_tune.bt_split >= 2 || B <= (int64_t)1300 * h->num_sms);
    size_t smem_bt_use = smem_bt;
    if (bt_split) {
        const int JP = h->K <= 16 ? 4 : h->K <= 32 ? 8 : h->K <= 48 ? 12 : 16;
        bt_seq = bt_con = JP == 4 ? backtrace_split_kernel<4, 8> : JP == 8 ? backtrace_split_kernel<8, 8>
                        : JP == 12 ? backtrace_split_kernel<12, 8> : backtrace_split_kernel<16, 8>;
        smem_bt_use = ((size_t)(h->K + 1) * (h->K | 1) + 4 * JP) * 8;
    }
    CUDA_TRY(cudaFuncSetAttribute(bt_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bt_use));
    CUDA_TRY(cudaFuncSetAttribute(bt_con, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bt_use));
    // same shared-memory carve-out as the forward kernel, or the two kernels cannot share an SM
    CUDA_TRY(cudaFuncSetAttribute(bt_con, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    const int64_t nblk = ((int64_t)ntiles * NS * (bt_split ? 4 : 1) + 127) / 128;
    if (concurrent) {
        CUDA_TRY(cudaStreamWaitEvent(w.st_bt, w.ev_pre, 0));
        if (wait32(w.st_bt, (unsigned long long)(uintptr_t)p.started, (unsigned int)std::max(1, grid), 0x0 /* GEQ */) != 0)
            return fail(CV_ERR_CUDA, "cuStreamWaitValue32 failed");
        bt_con<<<(unsigned)std::max<int64_t>(1, nblk), 128, smem_bt_use, w.st_bt>>>(p);   // blocks in tile order
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(w.ev_bt, w.st_bt));
        CUDA_TRY(cudaStreamWaitEvent(st, w.ev_bt, 0));
        if (long_launched) CUDA_TRY(cudaStreamWaitEvent(st, h->ev_long_done, 0));
        if (bt_prof) {
            CUDA_TRY(cudaEventRecord(h->ev2, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, h->ev0, h->ev1); cudaEventElapsedTime(&b, h->ev0, h->ev2);
            fprintf(stderr, "[cv] concurrent: forward %.3f ms, forward + backtrace %.3f ms (B = %lld)\n", a, b, (long long)B);
        }
        return CV_OK;
    }
    const int grid_bt = (int)std::max<int64_t>(1, std::min<int64_t>(nblk, (int64_t)h->num_sms * 12));
    bt_seq<<<grid_bt, 128, smem_bt_use, st>>>(p);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    if (timing) CUDA_TRY(cudaEventRecord(h->ev2, st));
    if (long_launched) CUDA_TRY(cudaStreamWaitEvent(st, h->ev_long_done, 0));
    return CV_OK;
}

// One chunk of sequences [0, B) of d_off (offsets are absolute into d_obs / d_path): order by length,
// forward, backtrace; everything enqueued on `st` with workspace set `w`.
static int decode_chunk(cv_hmm *h, DecodeWs &w, const uint32_t *d_obs, const int64_t *d_off, int64_t B, int64_t N,
                        int64_t max_len, uint32_t *d_path, double *d_score, cudaStream_t st, bool timing)
{
    int rc;
    if (B > 0x7fffffffLL) return fail(CV_ERR_UNSUPPORTED, "more than 2^31 sequences in one chunk");
    if ((rc = w.order.ensure(sizeof(uint32_t) * (size_t)B))) return rc;
    if ((rc = w.keys_in.ensure(sizeof(uint32_t) * (size_t)B))) return rc;
    if ((rc = w.keys_out.ensure(sizeof(uint32_t) * (size_t)B))) return rc;
    if ((rc = w.vals_in.ensure(sizeof(uint32_t) * (size_t)B))) return rc;
    if ((rc = w.misc.ensure(256))) return rc;
    unsigned int *d_counter = (unsigned int *)w.misc.p;
    int *d_status = (int *)w.misc.p + 16;
    CUDA_TRY(cudaMemsetAsync(w.misc.p, 0, 256, st));
    const bool chain_all = h->K <= SMALL_K_MAX && (g_tune.chain_max_b < 0 ? B <= 8192 : B <= g_tune.chain_max_b);
    LongSplit ls{0u, 0u, nullptr, nullptr, nullptr};
    if (!chain_all && (rc = long_split_setup(h, w, B, N, max_len, st, ls))) return rc;
    // order sequences by length, longest first (stable radix sort => deterministic)
    seq_len_keys_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(d_off, B, (uint32_t *)w.keys_in.p,
                                                                    (uint32_t *)w.vals_in.p, d_status, ls, max_len);
    g_launches++;
    // the radix sort only visits the bits a length can have (one 8-bit pass for lengths up to 255 instead of four)
    const int end_bit = max_len > 0 ? key_bits(max_len) : 32;
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, (uint32_t *)w.keys_in.p,
                                                       (uint32_t *)w.keys_out.p, (uint32_t *)w.vals_in.p,
                                                       (uint32_t *)w.order.p, (int)B, 0, end_bit, st));
    if ((rc = w.cub_tmp.ensure(tmp_bytes))) return rc;
    CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(w.cub_tmp.p, tmp_bytes, (uint32_t *)w.keys_in.p,
                                                       (uint32_t *)w.keys_out.p, (uint32_t *)w.vals_in.p,
                                                       (uint32_t *)w.order.p, (int)B, 0, end_bit, st));
    if (chain_all) {
        // few sequences: one warp per sequence (latency-oriented), backpointers as u8 rows
        if ((rc = w.hist.ensure((size_t)N * h->Kp + 64))) return rc;
        DecodeChainParams p;
        p.A = h->dA; p.BT = h->dBT; p.obs = d_obs; p.seq_off = d_off; p.order = (const uint32_t *)w.order.p;
        p.psi = (uint8_t *)w.hist.p; p.path = d_path; p.score = d_score; p.counter = d_counter; p.status = d_status;
        p.M = h->M; p.B = B; p.K = h->K; p.Kp = h->Kp;
        p.bt_in_smem = ((size_t)h->M * h->Kp * 8 <= CHAIN_BT_SMEM_MAX) ? 1 : 0;
        p.obs16 = h->obs16; p.path8 = h->path8;
        p.B_dev = nullptr; p.psi_stride = 0; p.chunk_done = nullptr; p.nch = 0;
        const int grid = (int)std::min<int64_t>((B + DC_WARPS - 1) / DC_WARPS, (int64_t)h->num_sms * 8);
        if (timing) CUDA_TRY(cudaEventRecord(h->ev0, st));
        if ((rc = launch_chain_kernel(h, p, grid, st))) return rc;
        if (timing) { CUDA_TRY(cudaEventRecord(h->ev1, st)); CUDA_TRY(cudaEventRecord(h->ev2, st)); }
        return CV_OK;
    }
    if (h->K <= SMALL_K_MAX)
        return launch_decode_small(h, w, d_obs, d_off, B, N, d_path, d_score, d_counter, d_status, max_len, st, timing, nullptr, &ls);
    if (timing) CUDA_TRY(cudaEventRecord(h->ev0, st));
    if ((rc = launch_decode_large(h, w, d_obs, d_off, B, N, d_path, d_score, d_counter, d_status, max_len, st))) return rc;
    if (timing) { CUDA_TRY(cudaEventRecord(h->ev1, st)); CUDA_TRY(cudaEventRecord(h->ev2, st)); }
    return CV_OK;
}

static int chunk_count(const cv_hmm *h, int64_t B, bool timing, bool host_buffers)
{
    if (timing || h->K > SMALL_K_MAX) return 1;          // kernel timing wants one launch for the whole batch
    if (g_tune.chunks > 0) return (int)std::min<int64_t>(g_tune.chunks, std::max<int64_t>(B, 1));
    // >= ~6 tiles per resident CTA per chunk keeps the dynamic tile scheduler balanced.  Host buffers: six
    // chunks so that H2D / D2H copies hide behind the kernels (measured at the POS shape: 2 chunks 16.5 ms,
    // 4: 15.4, 6: 15.3, 8: 15.7, 12: 19.2).
    // Device buffers: one launch when the backtrace can run next to the forward kernel (launch_decode_small), else two
    // chunks so that the backtrace of one overlaps the forward pass of the other.
    const int64_t per_chunk = (int64_t)h->num_sms * 2 * 6 * 64;
    const int dev_chunks = (g_tune.bt_concurrent && stream_wait_value32() != nullptr) ? 1 : 2;
    // Host buffers, streamed past one launch (decode_streamed): 4 chunks measured best (2: 14.8 ms, 3: 14.3, 4: 14.2-14.3,
    // 6: 14.4, 10: 14.8, 16: 15.0 -- every chunk restarts the longest-first tile order); one launch per chunk: 6.
    const bool can_stream = g_tune.streamed && g_tune.bt_concurrent && stream_wait_value32() != nullptr;
    // Streaming pays from ~450 k sequences: below, every chunk's restart of the longest-first tile order costs more
    // than the copies it hides (measured: 250 k sequences 4.2 ms streamed in two chunks vs 3.0 ms on the device; one
    // H2D -> decode -> D2H pass is 3.7 ms), so shorter batches go through in one piece.
    // With the tapering chunks of decode_streamed the crossover is lower: 350 k sequences 5.05 ms streamed vs 5.62 ms in
    // one piece, 250 k 4.27 vs 4.03 ms.
    if (host_buffers && can_stream) return (g_tune.uneven_chunks ? B >= (int64_t)2000 * h->num_sms : B / per_chunk >= 4) ? 4 : 1;
    return (int)std::max<int64_t>(1, std::min<int64_t>(host_buffers ? 6 : dev_chunks, B / per_chunk));
}

static int report_status(const int *status_words, int n)
{
    for (int i = 0; i < n; i++) {
        const int s = status_words[i];
        if (s == CV_ERR_EMPTY) return fail(CV_ERR_EMPTY, "empty sequence in batch (reference: usize underflow panic)");
        if (s == CV_ERR_ARG) return fail(CV_ERR_ARG, "observation index >= M (reference: ndarray index panic)");
        if (s) return fail(s, "device status %d", s);
    }
    return CV_OK;
}

static int decode_batch_dev_impl(cv_hmm *h, const uint32_t *d_obs, const int64_t *d_off, int64_t B, int64_t N,
                                 int64_t max_len, uint32_t *d_path, double *d_score, void *stream, int sync_status);

extern "C" int cv_decode_batch_dev(cv_hmm *h, const uint32_t *d_obs, const int64_t *d_off, int64_t B, int64_t N,
                                   int64_t max_len, uint32_t *d_path, double *d_score, void *stream, int sync_status)
{
    return decode_batch_dev_impl(h, d_obs, d_off, B, N, max_len, d_path, d_score, stream, sync_status);
}

extern "C" int cv_decode_batch_dev_u8(cv_hmm *h, const uint32_t *d_obs, const int64_t *d_off, int64_t B, int64_t N,
                                      int64_t max_len, uint8_t *d_path, double *d_score, void *stream, int sync_status)
{
    if (!h) return fail(CV_ERR_ARG, "NULL model");
    if (h->K > SMALL_K_MAX) return fail(CV_ERR_UNSUPPORTED, "u8 paths are implemented for K <= %d", SMALL_K_MAX);
    h->path8 = 1;
    const int rc = decode_batch_dev_impl(h, d_obs, d_off, B, N, max_len, reinterpret_cast<uint32_t *>(d_path), d_score, stream, sync_status);
    h->path8 = 0;
    return rc;
}

extern "C" int cv_decode_batch_dev_f32(cv_hmm *h, const uint32_t *d_obs, const int64_t *d_off, int64_t B, int64_t N,
                                       int64_t max_len, uint32_t *d_path, double *d_score, void *stream, int sync_status)
{
    if (!h) return fail(CV_ERR_ARG, "NULL model");
    if (!h->dA32f) return fail(CV_ERR_UNSUPPORTED, "the f32 mode is implemented for K <= %d", SMALL_K_MAX);
    h->f32mode = 1;
    const int rc = decode_batch_dev_impl(h, d_obs, d_off, B, N, max_len, d_path, d_score, stream, sync_status);
    h->f32mode = 0;
    return rc;
}

static int decode_batch_dev_impl(cv_hmm *h, const uint32_t *d_obs, const int64_t *d_off, int64_t B, int64_t N,
                                 int64_t max_len, uint32_t *d_path, double *d_score, void *stream, int sync_status)
{
    if (!h) return fail(CV_ERR_ARG, "NULL model");
    if (B < 0 || N < 0) return fail(CV_ERR_ARG, "negative size");
    if (B == 0) return CV_OK;
    if (!d_obs || !d_off || !d_path) return fail(CV_ERR_ARG, "NULL buffer");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t user = (cudaStream_t)stream;
    const bool timing = g_timing.load() != 0;
    const int nch = chunk_count(h, B, timing, false);
    int rc;
    if (nch == 1) {
        if ((rc = decode_chunk(h, h->ws[0], d_obs, d_off, B, N, max_len, d_path, d_score, user, timing))) return rc;
    } else {
        // fork: chunk k runs on internal stream k % 2 so that backtrace(k) overlaps forward(k+1)
        if (max_len <= 0) return fail(CV_ERR_ARG, "max_len is required for the chunked device path");
        CUDA_TRY(cudaEventRecord(h->ev_fork, user));
        for (int k = 0; k < 2; k++) CUDA_TRY(cudaStreamWaitEvent(h->ws[k].st, h->ev_fork, 0));
        for (int k = 0; k < nch; k++) {
            const int64_t b0 = B * k / nch, b1 = B * (k + 1) / nch;
            DecodeWs &w = h->ws[k & 1];
            if ((rc = decode_chunk(h, w, d_obs, d_off + b0, b1 - b0, N, max_len, d_path,
                                   d_score ? d_score + b0 : nullptr, w.st, false)))
                return rc;
        }
        for (int k = 0; k < 2; k++) {
            CUDA_TRY(cudaEventRecord(h->ws[k].done, h->ws[k].st));
            CUDA_TRY(cudaStreamWaitEvent(user, h->ws[k].done, 0));
        }
    }
    if (sync_status || timing) {
        int *hs = (int *)h->pinned_status + 8;
        for (int k = 0; k < 2; k++)
            if (h->ws[k].misc.p) CUDA_TRY(cudaMemcpyAsync(hs + k, (int *)h->ws[k].misc.p + 16, sizeof(int), cudaMemcpyDeviceToHost, user));
            else hs[k] = 0;
        CUDA_TRY(cudaStreamSynchronize(user));
        if (timing) {
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
            h->last_ms = ms;
            CUDA_TRY(cudaEventElapsedTime(&ms, h->ev1, h->ev2));
            h->last_bt_ms = ms;
        }
        return report_status(hs, 2);
    }
    return CV_OK;
}

// cuStreamWriteValue32: a stream writes a word once everything before it in the stream is done
typedef int (*StreamWriteValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
static StreamWriteValue32Fn stream_write_value32()
{
    static StreamWriteValue32Fn fn = []() -> StreamWriteValue32Fn {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (StreamWriteValue32Fn)f;
    }();
    return fn;
}

// Streamed host path (K <= 64, large batches): ONE forward + backtrace launch over the whole batch while the
// observations still arrive chunk by chunk on a copy stream (each chunk followed by a stream write of the
// `arrived` word the forward CTAs poll) and the paths / scores of a chunk leave on another copy stream as soon as
// the backtrace has counted the chunk's sequences done (stream wait on chunk_done[c]).  The batch is ordered by
// (chunk, length), so the kernels consume the chunks in arrival order.  Against one launch per chunk this
// removes the per-chunk sort / launch / ramp costs: the copies hide behind 13 ms of kernels instead of 6 x 2.3 ms.
static int decode_streamed(cv_hmm *h, const uint32_t *obs_flat, const int64_t *seq_off, int64_t B, uint32_t *path_out,
                           double *score_out, int nch, bool *handled, uint32_t *d_path, double *d_score)
{
    *handled = false;
    const size_t ob = h->obs16 ? 2 : 4, pb = h->path8 ? 1 : 4;       // bytes per observation / per path element on the host side
    StreamWaitValue32Fn wait32 = stream_wait_value32();
    StreamWriteValue32Fn write32 = stream_write_value32();
    if (!g_tune.streamed || !g_tune.bt_concurrent || !wait32 || !write32 || nch < 2 || nch > STREAM_MAX_CHUNKS) return CV_OK;
    if (h->K > SMALL_K_MAX || (g_tune.chain_max_b < 0 ? B <= 8192 : B <= g_tune.chain_max_b) || B > 0x7fffffffLL) return CV_OK;
    const int64_t N = seq_off[B];
    DecodeWs &w = h->ws[0];
    cudaStream_t sk = w.st, s_in = h->ws[1].st, s_out = h->ws[1].st_bt;
    uint32_t *d_obs = (uint32_t *)h->obs.p;
    int64_t *d_off = (int64_t *)h->seq_off.p;
    int rc;
    if ((rc = w.order.ensure(sizeof(uint32_t) * (size_t)B)) || (rc = w.keys_in.ensure(sizeof(uint32_t) * (size_t)B)) ||
        (rc = w.keys_out.ensure(sizeof(uint32_t) * (size_t)B)) || (rc = w.vals_in.ensure(sizeof(uint32_t) * (size_t)B)) ||
        (rc = w.misc.ensure(256)))
        return rc;
    unsigned int *d_counter = (unsigned int *)w.misc.p;
    int *d_status = (int *)w.misc.p + 16;
    unsigned int *d_maxlen = (unsigned int *)w.misc.p + 17, *d_arrived = (unsigned int *)w.misc.p + 32, *d_chunk_done = d_arrived + 1;
    StreamedIO sio;
    sio.cbs.nch = nch;
    // Chunk sizes.  The forward kernel cannot start before chunk 0 is on the device and the last chunk's paths leave after
    // the kernels have ended, so both should be small -- but tiles are ordered by (chunk, length): every chunk starts with
    // its longest, most ragged tiles, and a small late chunk starts them when everything else is about to end (measured
    // with plain uneven chunks: 1/8, 3/8, 3/8, 1/8 13.45 vs 13.24 ms for four equal ones; six tapering chunks end the
    // kernels 0.5 ms later than five).  The automatic choice for large batches therefore cuts 10 / 25 / 25 / 20 / 12 / 8 %
    // and orders the LONG sequences of chunks 3.. with chunk 2 (ChunkBounds::promote_len): the late groups hold short,
    // even tiles only, a chunk's copy out fits into the work ordered behind it, and 8 % of the paths are copied after
    // the kernels.  A forced chunk count (cv_debug_set_chunks) gives equal chunks.
    sio.cbs.promote_len = 0; sio.cbs.promote_to = 0;
    if ((g_tune.chunks <= 0 && nch == 4 && g_tune.uneven_chunks) || (g_tune.uneven_chunks >= 2 && B >= 100)) {     // (2: forced, for tests)
        nch = 6;
        sio.cbs.nch = nch;
        static const int pct[7] = {0, 10, 35, 60, 80, 92, 100};
        for (int k = 0; k <= nch; k++) sio.cbs.cb[k] = B * pct[k] / 100;
        // a tile of a late group should take at most a quarter of the tile-steps a resident CTA runs for the last chunk
        const int64_t steps_per_cta = (seq_off[B] - seq_off[sio.cbs.cb[nch - 1]]) / ((int64_t)64 * 2 * h->num_sms);
        sio.cbs.promote_len = (int)std::max<int64_t>(4, std::min<int64_t>(steps_per_cta / 4, 1 << 20));
        sio.cbs.promote_to = 2;
    } else {
        for (int k = 0; k <= nch; k++) sio.cbs.cb[k] = B * k / nch;
    }
    sio.d_arrived = d_arrived; sio.d_chunk_done = d_chunk_done;

    // From here on copies from / to the caller's host buffers are in flight: every return (error or not) first drains
    // the three streams, so the caller may free or reuse its buffers as soon as this function returns.
    struct Drain { cudaStream_t s[3]; ~Drain() { for (cudaStream_t x : s) cudaStreamSynchronize(x); cudaGetLastError(); } } drain{{sk, s_in, s_out}};
    // CV_E2E_PROF=1: a timeline of this call (host clock for the enqueue / wait phases, events for the device phases)
    const bool prof = g_tune.e2e_prof != 0;
    cudaEvent_t pe[6] = {};
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    double hm_scan = 0.0, hm_enq = 0.0;
    if (prof) { for (auto &e : pe) cudaEventCreate(&e); cudaEventRecord(pe[0], sk); }
    // offsets first (the ordering needs nothing else), then the observations chunk by chunk on the copy stream
    CUDA_TRY(cudaMemsetAsync(w.misc.p, 0, 256, sk));
    CUDA_TRY(cudaMemcpyAsync(d_off, seq_off, sizeof(int64_t) * (size_t)(B + 1), cudaMemcpyHostToDevice, sk));
    CUDA_TRY(cudaEventRecord(w.ev_pre, sk));
    int *hs = (int *)h->pinned_status + 8;
    // `arrived` is zeroed first -- and the offsets have the link to themselves: the ordering (keys, sort) waits for them,
    // and two host-to-device copies in flight share the bandwidth (offsets + keys done at 0.5 instead of 0.25 ms)
    CUDA_TRY(cudaStreamWaitEvent(s_in, w.ev_pre, 0));
    for (int k = 0; k < nch; k++) {
        const int64_t e0 = seq_off[sio.cbs.cb[k]], e1 = seq_off[sio.cbs.cb[k + 1]];
        if (e1 < e0 || e1 > N) return fail(CV_ERR_ARG, "seq_off not monotone");
        if (e1 > e0) CUDA_TRY(cudaMemcpyAsync((char *)d_obs + ob * e0, (const char *)obs_flat + ob * e0, ob * (size_t)(e1 - e0), cudaMemcpyHostToDevice, s_in));
        if (write32(s_in, (unsigned long long)(uintptr_t)d_arrived, (unsigned int)(k + 1), 0x0) != 0) return fail(CV_ERR_CUDA, "cuStreamWriteValue32 failed");
    }
    CUDA_TRY(cudaEventRecord(h->ev_all_in, s_in));
    sio.all_in = h->ev_all_in;
    // the long-sequence split needs the longest length before the keys kernel: small batches only (host scan below
    // is moved up for them); large batches have nothing that long relative to their size
    LongSplit ls{0u, 0u, nullptr, nullptr, nullptr};
    int64_t max_len_host = 0;
    if (B <= (1 << 18)) {
        for (int64_t b = 0; b < B; b++) {
            const int64_t len = seq_off[b + 1] - seq_off[b];
            if (len <= 0) return len == 0 ? fail(CV_ERR_EMPTY, "empty sequence in batch (reference: usize underflow panic)")
                                          : fail(CV_ERR_ARG, "seq_off not monotone");
            max_len_host = std::max(max_len_host, len);
        }
        if (max_len_host <= 0xffffffLL && (rc = long_split_setup(h, w, B, N, max_len_host, sk, ls))) return rc;
    }
    // sort key = chunk field << shift | length: 24 bits of length when the longest one is not known yet, else just the
    // bits it needs (the sort then takes one or two 8-bit passes instead of four)
    sio.cbs.shift = (max_len_host > 0 && max_len_host <= 0xffffffLL) ? key_bits(max_len_host) : 24;
    seq_chunk_keys_kernel<<<(unsigned)((B + 255) / 256), 256, 0, sk>>>(d_off, B, sio.cbs, (uint32_t *)w.keys_in.p,
                                                                      (uint32_t *)w.vals_in.p, d_status, d_maxlen, ls);
    g_launches++;
    // Lengths are checked and the longest one is found before the launch (the history is sized from it).  Small
    // batches (a slice of a sharded batch): on the HOST while the first copies fly -- 0.1 ms per 100 k sequences and no
    // device round trip.  Large ones: by the keys kernel, read back with one ~0.2 ms synchronisation (a single host
    // thread would need ~0.7 ms for a million offsets, which delays the launch past the arrival of chunk 0).
    int64_t max_len = max_len_host;
    if (B > (1 << 18)) {
        CUDA_TRY(cudaMemcpyAsync(hs, d_status, 2 * sizeof(int), cudaMemcpyDeviceToHost, sk));  // status, longest length
        CUDA_TRY(cudaStreamSynchronize(sk));
        if (hs[0] == CV_ERR_EMPTY) return fail(CV_ERR_EMPTY, "empty sequence in batch (reference: usize underflow panic)");
        if (hs[0] == CV_ERR_ARG) return fail(CV_ERR_ARG, "seq_off not monotone");
        if (hs[0]) { *handled = false; return CV_OK; }                                         // e.g. a sequence longer than 2^24: chunked path
        max_len = (int64_t)(unsigned int)hs[1];
    }
    if (max_len > 0xffffffLL) { *handled = false; return CV_OK; }                              // does not fit the 24-bit sort key: chunked path
    if (prof) { hm_scan = host_ms(); cudaEventRecord(pe[1], sk); cudaEventRecord(pe[4], s_in); }

    const int end_bit = std::min(32, sio.cbs.shift + key_bits(nch - 1));
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, (uint32_t *)w.keys_in.p, (uint32_t *)w.keys_out.p,
                                                       (uint32_t *)w.vals_in.p, (uint32_t *)w.order.p, (int)B, 0, end_bit, sk));
    if ((rc = w.cub_tmp.ensure(tmp_bytes))) return rc;
    CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(w.cub_tmp.p, tmp_bytes, (uint32_t *)w.keys_in.p, (uint32_t *)w.keys_out.p,
                                                       (uint32_t *)w.vals_in.p, (uint32_t *)w.order.p, (int)B, 0, end_bit, sk));
    if (prof) cudaEventRecord(pe[2], sk);
    if ((rc = launch_decode_small(h, w, d_obs, d_off, B, N, d_path, d_score, d_counter, d_status, max_len, sk, false, &sio, &ls))) return rc;
    if (prof) cudaEventRecord(pe[3], sk);
    // paths / scores of chunk c leave as soon as the backtrace has counted all its sequences
    CUDA_TRY(cudaStreamWaitEvent(s_out, w.ev_pre, 0));
    for (int k = 0; k < nch; k++) {
        const int64_t b0 = sio.cbs.cb[k], b1 = sio.cbs.cb[k + 1], e0 = seq_off[b0], e1 = seq_off[b1];
        if (b1 == b0) continue;
        if (wait32(s_out, (unsigned long long)(uintptr_t)(d_chunk_done + k), (unsigned int)(b1 - b0), 0x0 /* GEQ */) != 0)
            return fail(CV_ERR_CUDA, "cuStreamWaitValue32 failed");
        CUDA_TRY(cudaMemcpyAsync((char *)path_out + pb * e0, (const char *)d_path + pb * e0, pb * (size_t)(e1 - e0), cudaMemcpyDeviceToHost, s_out));
        if (score_out) CUDA_TRY(cudaMemcpyAsync(score_out + b0, d_score + b0, sizeof(double) * (size_t)(b1 - b0), cudaMemcpyDeviceToHost, s_out));
    }
    CUDA_TRY(cudaMemcpyAsync(hs, d_status, sizeof(int), cudaMemcpyDeviceToHost, sk));
    hs[1] = 0;
    if (prof) { hm_enq = host_ms(); cudaEventRecord(pe[5], s_out); }
    CUDA_TRY(cudaStreamSynchronize(sk));
    CUDA_TRY(cudaStreamSynchronize(s_out));
    CUDA_TRY(cudaStreamSynchronize(s_in));
    if (prof) {
        const double hm_end = host_ms();
        float t[6] = {};
        for (int i = 1; i < 6; i++) cudaEventElapsedTime(&t[i], pe[0], pe[i]);
        fprintf(stderr, "[cv] streamed B=%lld nch=%d | host: lengths known %.2f, all enqueued %.2f, done %.2f ms | device (from first enqueue): "
                        "offsets+keys %.2f, sorted %.2f, forward+backtrace done %.2f, last chunk in %.2f, last copy out %.2f ms\n",
                (long long)B, nch, hm_scan, hm_enq, hm_end, t[1], t[2], t[3], t[4], t[5]);
        for (auto &e : pe) cudaEventDestroy(e);
    }
    *handled = true;
    return report_status(hs, 1);
}

static int decode_batch_host(cv_hmm *h, const uint32_t *obs_flat, const int64_t *seq_off, int64_t B, uint32_t *path_out,
                             double *score_out, uint32_t *d_path_keep, double *d_score_keep);

extern "C" int cv_decode_batch(cv_hmm *h, const uint32_t *obs_flat, const int64_t *seq_off, int64_t B,
                               uint32_t *path_out, double *score_out)
{
    return decode_batch_host(h, obs_flat, seq_off, B, path_out, score_out, nullptr, nullptr);
}

extern "C" int cv_decode_batch_keep(cv_hmm *h, const uint32_t *obs_flat, const int64_t *seq_off, int64_t B,
                                    uint32_t *path_out, double *score_out, uint32_t *d_path_keep, double *d_score_keep)
{
    if (!d_path_keep || !d_score_keep) return fail(CV_ERR_ARG, "NULL device buffer");
    return decode_batch_host(h, obs_flat, seq_off, B, path_out, score_out, d_path_keep, d_score_keep);
}

extern "C" int cv_decode_batch_f32(cv_hmm *h, const uint32_t *obs_flat, const int64_t *seq_off, int64_t B,
                                   uint32_t *path_out, double *score_out)
{
    if (!h) return fail(CV_ERR_ARG, "NULL model");
    if (!h->dA32f) return fail(CV_ERR_UNSUPPORTED, "the f32 mode is implemented for K <= %d", SMALL_K_MAX);
    h->f32mode = 1;
    const int rc = decode_batch_host(h, obs_flat, seq_off, B, path_out, score_out, nullptr, nullptr);
    h->f32mode = 0;
    return rc;
}

extern "C" int cv_decode_batch_u16u8(cv_hmm *h, const uint16_t *obs_flat, const int64_t *seq_off, int64_t B,
                                     uint8_t *path_out, double *score_out)
{
    if (!h) return fail(CV_ERR_ARG, "NULL model");
    if (h->M > 65536) return fail(CV_ERR_UNSUPPORTED, "u16 observations need M <= 65536 (M = %lld)", (long long)h->M);
    if (h->K > 256) return fail(CV_ERR_UNSUPPORTED, "u8 paths need K <= 256 (K = %d)", h->K);
    if (h->K > SMALL_K_MAX) return fail(CV_ERR_UNSUPPORTED, "the narrow host formats are implemented for K <= %d", SMALL_K_MAX);
    h->obs16 = 1; h->path8 = 1;
    const int rc = decode_batch_host(h, reinterpret_cast<const uint32_t *>(obs_flat), seq_off, B,
                                     reinterpret_cast<uint32_t *>(path_out), score_out, nullptr, nullptr);
    h->obs16 = 0; h->path8 = 0;
    return rc;
}

static int decode_batch_host(cv_hmm *h, const uint32_t *obs_flat, const int64_t *seq_off, int64_t B, uint32_t *path_out,
                             double *score_out, uint32_t *d_path_keep, double *d_score_keep)
{
    if (!h) return fail(CV_ERR_ARG, "NULL model");
    if (B < 0) return fail(CV_ERR_ARG, "negative B");
    if (B == 0) return CV_OK;
    if (!obs_flat || !seq_off || !path_out) return fail(CV_ERR_ARG, "NULL buffer");
    if (seq_off[0] != 0) return fail(CV_ERR_ARG, "seq_off[0] must be 0");
    const int64_t N = seq_off[B];
    CUDA_TRY(cudaSetDevice(h->device));
    int rc;
    if ((rc = h->obs.ensure(sizeof(uint32_t) * (size_t)N))) return rc;
    if ((rc = h->seq_off.ensure(sizeof(int64_t) * (size_t)(B + 1)))) return rc;
    if (!d_path_keep && (rc = h->path.ensure(sizeof(uint32_t) * (size_t)N))) return rc;
    if (!d_score_keep && (rc = h->score.ensure(sizeof(double) * (size_t)B))) return rc;
    // the device copies of the results: the library's own buffers, or the caller's (cv_decode_batch_keep)
    uint32_t *d_obs = (uint32_t *)h->obs.p, *d_path = d_path_keep ? d_path_keep : (uint32_t *)h->path.p;
    int64_t *d_off = (int64_t *)h->seq_off.p;
    double *d_score = d_score_keep ? d_score_keep : (double *)h->score.p;
    const bool timing = g_timing.load() != 0;
    const int nch = chunk_count(h, B, timing, true);
    const size_t ob = h->obs16 ? 2 : 4, pb = h->path8 ? 1 : 4;
    if (!timing) {
        bool handled = false;
        if ((rc = decode_streamed(h, obs_flat, seq_off, B, path_out, score_out, nch, &handled, d_path, d_score))) return rc;
        if (handled) return CV_OK;
    }
    // Chunk k: H2D of its observations -> order/forward/backtrace -> D2H of its paths, on stream k % 2, so the
    // copies of one chunk overlap the kernels of its neighbours.  Offsets are validated while the first copy flies.
    std::vector<int64_t> cb;                                            // chunk boundaries (sequence indices)
    for (int k = 0; k <= nch; k++) cb.push_back(B * k / nch);
    const int nck = (int)cb.size() - 1;
    const bool e2e_prof = g_tune.e2e_prof != 0;      // prints a per-chunk timeline (copy in / kernels / copy out)
    std::vector<cudaEvent_t> pev;
    auto mark = [&](cudaStream_t s_) { if (e2e_prof) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s_); pev.push_back(e); } };
    if (e2e_prof) { cudaDeviceSynchronize(); mark(h->ws[0].st); }
    for (int k = 0; k < nck; k++) {
        const int64_t b0 = cb[k], b1 = cb[k + 1];
        int64_t max_len = 0;
        for (int64_t b = b0; b < b1; b++) {
            const int64_t len = seq_off[b + 1] - seq_off[b];
            if (len < 0) return fail(CV_ERR_ARG, "seq_off not monotone at %lld", (long long)b);
            if (len == 0) return fail(CV_ERR_EMPTY, "sequence %lld is empty (reference: usize underflow panic)", (long long)b);
            max_len = std::max(max_len, len);
        }
        DecodeWs &w = h->ws[k & 1];
        cudaStream_t st = nch == 1 ? h->stream : w.st;
        const int64_t e0 = seq_off[b0], e1 = seq_off[b1];
        CUDA_TRY(cudaMemcpyAsync(d_off + b0, seq_off + b0, sizeof(int64_t) * (size_t)(b1 - b0 + 1), cudaMemcpyHostToDevice, st));
        mark(st);
        CUDA_TRY(cudaMemcpyAsync((char *)d_obs + ob * e0, (const char *)obs_flat + ob * e0, ob * (size_t)(e1 - e0), cudaMemcpyHostToDevice, st));
        mark(st);
        if ((rc = decode_chunk(h, w, d_obs, d_off + b0, b1 - b0, e1 - e0, max_len, d_path, d_score + b0, st, timing))) return rc;
        mark(st);
        CUDA_TRY(cudaMemcpyAsync((char *)path_out + pb * e0, (const char *)d_path + pb * e0, pb * (size_t)(e1 - e0), cudaMemcpyDeviceToHost, st));
        if (score_out)
            CUDA_TRY(cudaMemcpyAsync(score_out + b0, d_score + b0, sizeof(double) * (size_t)(b1 - b0), cudaMemcpyDeviceToHost, st));
        mark(st);
    }
    if (e2e_prof) {
        cudaDeviceSynchronize();
        for (int k = 0; k < nck; k++) {
            float t[4];
            for (int i = 0; i < 4; i++) cudaEventElapsedTime(&t[i], pev[0], pev[1 + 4 * k + i]);
            fprintf(stderr, "[cv] chunk %d (%lld seqs): in %.2f-%.2f  kernels -%.2f  out -%.2f ms\n", k, (long long)(cb[k + 1] - cb[k]), t[0], t[1], t[2], t[3]);
        }
        for (cudaEvent_t e : pev) cudaEventDestroy(e);
    }
    int *hs = (int *)h->pinned_status + 8;
    hs[0] = hs[1] = 0;
    if (nch == 1) {
        CUDA_TRY(cudaMemcpyAsync(hs, (int *)h->ws[0].misc.p + 16, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    } else {
        for (int k = 0; k < 2; k++) {
            CUDA_TRY(cudaMemcpyAsync(hs + k, (int *)h->ws[k].misc.p + 16, sizeof(int), cudaMemcpyDeviceToHost, h->ws[k].st));
            CUDA_TRY(cudaStreamSynchronize(h->ws[k].st));
        }
    }
    if (timing) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->last_ms = ms; else cudaGetLastError();
        if (cudaEventElapsedTime(&ms, h->ev1, h->ev2) == cudaSuccess) h->last_bt_ms = ms; else cudaGetLastError();
    }
    return report_status(hs, 2);
}

