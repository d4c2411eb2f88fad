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
    const uint32_t *d_order = (const uint32_t *)w.order.p, *d_sorted_len = (const uint32_t *)w.keys_out.p;
    const int Kl = h->Kl, NCB = Kl / LARGE_BN;
    const int64_t NRB = (B + LG_BM - 1) / LG_BM, Bpad = NRB * LG_BM;
    if (NRB > 0x7fffffffLL) return fail(CV_ERR_UNSUPPORTED, "batch too large");
    int rc;
    if (max_len <= 0) {   // not supplied: read the longest length back (one sync)
        uint32_t L = 0;
        CUDA_TRY(cudaMemcpyAsync(&L, d_sorted_len, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        max_len = L;
    }
    if (max_len > 0x7ffffff0LL) return fail(CV_ERR_UNSUPPORTED, "sequence too long");
    const int Tmax = (int)max_len;
    const int psi16 = Kl > 256 ? 1 : 0;
    if ((rc = w.hist.ensure((size_t)N * Kl * (psi16 ? 2 : 1)))) return rc;
    if ((rc = w.delta_g.ensure((size_t)2 * Kl * Bpad * sizeof(double)))) return rc;
    DevBuf &b_arr = w.lg_arr, &b_start = w.lg_start, &b_done = w.lg_done, &b_tmp = w.cub_tmp;
    if ((rc = b_arr.ensure(sizeof(long long) * (size_t)(Tmax + 2)))) return rc;
    if ((rc = b_start.ensure(sizeof(long long) * (size_t)(Tmax + 2)))) return rc;
    if ((rc = b_done.ensure(sizeof(unsigned int) * (size_t)NRB))) return rc;
    CUDA_TRY(cudaMemsetAsync(b_done.p, 0, sizeof(unsigned int) * (size_t)NRB, st));
    CUDA_TRY(cudaMemsetAsync(w.delta_g.p, 0, (size_t)Kl * Bpad * sizeof(double), st));   // delta(0) = 0.0 (viterbi.rs:6)

    step_items_kernel<<<(Tmax + 1 + 255) / 256, 256, 0, st>>>(d_sorted_len, (int)NRB, NCB, Tmax, (long long *)b_arr.p);
    g_launches++;
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (long long *)b_arr.p, (long long *)b_start.p, Tmax + 1, st));
    if ((rc = b_tmp.ensure(tmp_bytes))) return rc;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(b_tmp.p, tmp_bytes, (long long *)b_arr.p, (long long *)b_start.p, Tmax + 1, st));

    DecodeLargeParams p;
    p.A = h->dAl; p.BT = h->dBTl; p.obs = d_obs; p.seq_off = d_off; p.order = d_order; p.sorted_len = d_sorted_len;
    p.path = d_path; p.score = d_score; p.delta = (double *)w.delta_g.p; p.psi = w.hist.p;
    p.step_start = (const long long *)b_start.p; p.item_counter = (unsigned long long *)d_counter;
    p.done = (unsigned int *)b_done.p; p.status = d_status;
    p.M = h->M; p.B = B; p.Bpad = Bpad; p.K = h->K; p.Kl = Kl; p.NCB = NCB; p.NRB = (int)NRB; p.Tmax = Tmax;
    p.psi16 = psi16; p.zero = 0;
    if (Tmax > 1) {
        auto kern = decode_large_kernel<CVB_CELL_VARIANT>;
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LG_SMEM_BYTES));
        kern<<<h->num_sms, LG_THREADS, LG_SMEM_BYTES, st>>>(p);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    backtrace_large_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(p);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return CV_OK;
}
