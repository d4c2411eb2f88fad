// host launch logic of the large-K kernel (included by cv_api.cu)

// nact(t) = number of row blocks whose longest sequence is longer than t; arr[t] = nact(t) * NCB items.
__global__ void step_items_kernel(const uint32_t *sorted_len, int NRB, int NCB, int Tmax, long long *arr)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > Tmax) return;
    long long n = 0;
    if (t >= 1 && t < Tmax) {
        int lo = 0, hi = NRB;                   // first rb with len(rb) <= t  (lengths are descending)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (sorted_len[(size_t)mid * LG_BM] > (uint32_t)t) lo = mid + 1; else hi = mid;
        }
        n = (long long)lo * NCB;
    }
    arr[t] = n;
}

static int launch_decode_large(cv_hmm *h, DecodeWs &w, const uint32_t *d_obs, const int64_t *d_off, int64_t B, int64_t N,
                               uint32_t *d_path, double *d_score, unsigned int *d_counter, int *d_status, int64_t max_len,
                               cudaStream_t st)
{
    (void)N;
    const uint32_t *d_order = (const uint32_t *)w.order.p, *d_sorted_len = (const uint32_t *)w.keys_out.p;
    const int Kl = h->Kl, NCB = Kl / LARGE_BN;
    const int64_t NRB_all = (B + LG_BM - 1) / LG_BM;
    if (NRB_all > 0x7fffffffLL) return fail(CV_ERR_UNSUPPORTED, "batch too large");
    int rc;
    if (max_len <= 0) {   // not supplied: read the longest length back (one sync)
        uint32_t L = 0;
        CUDA_TRY(cudaMemcpyAsync(&L, d_sorted_len, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        max_len = L;
    }
    if (max_len > 0x7ffffff0LL) return fail(CV_ERR_UNSUPPORTED, "sequence too long");
    const int Tmax = (int)max_len;
    // delta history [Tmax][Kl][64 * NRB_group]: run the batch in groups of row blocks that fit in HBM
    const size_t per_rb = (size_t)Tmax * Kl * LG_BM * sizeof(double);
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    const size_t budget = std::max<size_t>(w.hist.bytes, (size_t)((free_b + w.hist.bytes) * 0.85));
    int64_t rb_per_group = (int64_t)std::min<size_t>((size_t)NRB_all, budget / std::max<size_t>(per_rb, 1));
    if (g_tune.large_group_rb > 0) rb_per_group = std::max<int64_t>(1, std::min<int64_t>(NRB_all, (int64_t)g_tune.large_group_rb));
    if (rb_per_group < 1) return fail(CV_ERR_OOM, "delta history of one row block (%zu bytes) does not fit in device memory", per_rb);
    if ((rc = w.hist.ensure(per_rb * (size_t)rb_per_group))) return rc;
    DevBuf &b_arr = w.lg_arr, &b_start = w.lg_start, &b_done = w.lg_done, &b_tmp = w.cub_tmp;
    if ((rc = b_arr.ensure(sizeof(long long) * (size_t)(Tmax + 2)))) return rc;
    if ((rc = b_start.ensure(sizeof(long long) * (size_t)(Tmax + 2)))) return rc;
    if ((rc = b_done.ensure(sizeof(unsigned int) * (size_t)rb_per_group))) return rc;
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (long long *)b_arr.p, (long long *)b_start.p, Tmax + 1, st));
    if ((rc = b_tmp.ensure(tmp_bytes))) return rc;
    CUDA_TRY(cudaFuncSetAttribute(decode_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LG_SMEM_BYTES));

    for (int64_t rb0 = 0; rb0 < NRB_all; rb0 += rb_per_group) {
        const int64_t NRB = std::min<int64_t>(rb_per_group, NRB_all - rb0), Bpad = NRB * LG_BM;
        const uint32_t *grp_len = d_sorted_len + rb0 * LG_BM;
        CUDA_TRY(cudaMemsetAsync(b_done.p, 0, sizeof(unsigned int) * (size_t)NRB, st));
        CUDA_TRY(cudaMemsetAsync(d_counter, 0, 2 * sizeof(unsigned int), st));
        CUDA_TRY(cudaMemsetAsync(w.hist.p, 0, (size_t)Kl * Bpad * sizeof(double), st));    // slab 0: delta(0) = 0.0 (viterbi.rs:6)
        // items per step for this group (its first row block holds the longest sequence of the group)
        uint32_t gmax = (uint32_t)Tmax;
        step_items_kernel<<<(Tmax + 1 + 255) / 256, 256, 0, st>>>(grp_len, (int)NRB, NCB, Tmax, (long long *)b_arr.p);
        g_launches++;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(b_tmp.p, tmp_bytes, (long long *)b_arr.p, (long long *)b_start.p, Tmax + 1, st));
        (void)gmax;
        DecodeLargeParams p;
        p.A = h->dAl; p.AT = h->dATl; p.BT = h->dBTl; p.obs = d_obs; p.seq_off = d_off; p.order = d_order;
        p.sorted_len = d_sorted_len; p.path = d_path; p.score = d_score; p.hist = (double *)w.hist.p;
        p.step_start = (const long long *)b_start.p; p.item_counter = (unsigned long long *)d_counter;
        p.done = (unsigned int *)b_done.p; p.status = d_status;
        p.M = h->M; p.B = B; p.Bpad = Bpad; p.rank0 = rb0 * LG_BM; p.K = h->K; p.Kl = Kl; p.NCB = NCB; p.NRB = (int)NRB;
        p.Tmax = Tmax;
        if (Tmax > 1) {
            decode_large_kernel<<<h->num_sms, LG_THREADS, LG_SMEM_BYTES, st>>>(p);
            g_launches++;
            CUDA_TRY(cudaGetLastError());
        }
        backtrace_large_kernel<<<(unsigned)((Bpad * 8 + 255) / 256), 256, 0, st>>>(p);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return CV_OK;
}
