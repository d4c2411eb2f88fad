// cfn_host.inl -- cv_cfn_tables: the numeric part of write_cfn (reference src/viterbi_solver/cfn.rs:82-167)
// (included by cv_api.cu).  Host: component boundaries, chain lists, lower bound; device: every sweep and the
// ordered accumulation (cfn.cuh).

extern "C" int cv_cfn_tables(cv_hmm *h, const uint32_t *obs, const uint8_t *is_seq_start, const int32_t *comp, int64_t N,
                             int32_t k, double *cost_tables, double *unary, double *lower_bound_out, int64_t *nboundaries_out,
                             double *device_ms_out)
{
    if (!h) return fail(CV_ERR_ARG, "NULL model");
    if (N <= 0) return fail(CV_ERR_EMPTY, "empty super-sequence");
    if (!obs || !is_seq_start || !comp || !cost_tables || !unary) return fail(CV_ERR_ARG, "NULL buffer");
    if (k <= 0) return fail(CV_ERR_EMPTY, "no constraint component (reference: constraint_boundaries.last().unwrap() panics, cfn.rs:108)");
    if (h->K > SMALL_K_MAX) return fail(CV_ERR_UNSUPPORTED, "cv_cfn_tables covers K <= %d", SMALL_K_MAX);
    const int K = h->K;
    for (int64_t t = 0; t < N; t++) {
        if ((int64_t)obs[t] >= h->M) return fail(CV_ERR_ARG, "observation index >= M at %lld (reference: ndarray index panic)", (long long)t);
        if (comp[t] >= k) return fail(CV_ERR_ARG, "component id %d >= k %d at %lld (reference: cost_tables index panic)", comp[t], k, (long long)t);
    }
    // cfn.rs:82-107: a boundary wherever the component of the constrained elements changes
    std::vector<int64_t> bt; std::vector<int32_t> bc;
    int32_t last_cid = -1;
    for (int64_t t = 0; t < N; t++)
        if (comp[t] >= 0) {
            if (last_cid < 0 || last_cid != comp[t]) { bt.push_back(t); bc.push_back(comp[t]); }
            last_cid = comp[t];
        }
    if (bt.empty()) return fail(CV_ERR_EMPTY, "no constrained element (reference: constraint_boundaries.last().unwrap() panics, cfn.rs:108)");
    const int64_t nb = (int64_t)bt.size(), npairs = nb - 1;

    // chains: K per boundary pair (one per n_from), 1 unary start, K unary end
    std::vector<CfnChain> chains;
    chains.reserve((size_t)npairs * K + K + 1);
    for (int64_t i = 0; i < npairs; i++)
        for (int n1 = 0; n1 < K; n1++) chains.push_back(CfnChain{bt[i], bt[i + 1], n1, 0, (i * K + n1) * (int64_t)K});
    const int64_t out_start = npairs * (int64_t)K * K, out_end = out_start + K;
    chains.push_back(CfnChain{0, bt[0], -1, 1, out_start});
    const bool has_tail = bt[nb - 1] != N - 1;
    if (has_tail) for (int n = 0; n < K; n++) chains.push_back(CfnChain{bt[nb - 1], N - 1, n, 2, out_end + n});
    const int64_t nout = out_end + K, nchains = (int64_t)chains.size();
    // boundary pairs of every unordered component pair c1 < c2, ascending; bit 0 = the pair runs c2 -> c1
    std::vector<std::vector<int64_t>> per_tab((size_t)k * k);
    for (int64_t i = 0; i < npairs; i++) {
        const int cf = bc[i], ct = bc[i + 1];
        per_tab[(size_t)std::min(cf, ct) * k + std::max(cf, ct)].push_back(i * 2 + (cf < ct ? 0 : 1));
    }
    std::vector<int64_t> plist, toff{0}; std::vector<int32_t> tcs;             // tcs = [c1 of every table | c2 of every table]
    std::vector<int32_t> t1s, t2s;
    for (int c1 = 0; c1 < k; c1++)
        for (int c2 = c1 + 1; c2 < k; c2++) {
            const auto &v = per_tab[(size_t)c1 * k + c2];
            if (v.empty()) continue;
            plist.insert(plist.end(), v.begin(), v.end());
            toff.push_back((int64_t)plist.size()); t1s.push_back(c1); t2s.push_back(c2);
        }
    const int ntab = (int)t1s.size();
    std::vector<int64_t> acc_i64(plist); acc_i64.insert(acc_i64.end(), toff.begin(), toff.end());
    tcs = t1s; tcs.insert(tcs.end(), t2s.begin(), t2s.end());

    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    DevBuf *b = h->cpb;
    int rc;
    if ((rc = b[2].ensure(sizeof(uint32_t) * (size_t)N)) || (rc = b[3].ensure((size_t)N)) || (rc = b[4].ensure(sizeof(int32_t) * (size_t)N)) ||
        (rc = b[6].ensure(sizeof(CfnChain) * (size_t)nchains)) || (rc = b[7].ensure(sizeof(int32_t) * (tcs.size() + 4))) ||
        (rc = b[8].ensure(sizeof(int64_t) * (acc_i64.size() + 4))) ||
        (rc = b[10].ensure(sizeof(double) * (size_t)nout)) || (rc = b[14].ensure(sizeof(double) * (size_t)k * k * K * K)))
        return rc;
    h->cp_N = 0;                                                                   // the CP state hooks no longer describe these buffers
    CUDA_TRY(cudaMemcpyAsync(b[2].p, obs, sizeof(uint32_t) * (size_t)N, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b[3].p, is_seq_start, (size_t)N, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b[4].p, comp, sizeof(int32_t) * (size_t)N, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b[6].p, chains.data(), sizeof(CfnChain) * (size_t)nchains, cudaMemcpyHostToDevice, st));
    if (ntab > 0) {
        CUDA_TRY(cudaMemcpyAsync(b[7].p, tcs.data(), sizeof(int32_t) * tcs.size(), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(b[8].p, acc_i64.data(), sizeof(int64_t) * acc_i64.size(), cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemsetAsync(b[10].p, 0, sizeof(double) * (size_t)nout, st));
    CUDA_TRY(cudaMemsetAsync(b[14].p, 0, sizeof(double) * (size_t)k * k * K * K, st));   // Array2::from_elem(.., 0.0) cfn.rs:116
    CpParams p{};
    p.A = h->dA; p.BT = h->dBT; p.Pi = h->dPi;
    p.obs = (const uint32_t *)b[2].p; p.start = (const uint8_t *)b[3].p; p.comp = (const int32_t *)b[4].p;
    p.N = N; p.M = h->M; p.K = K; p.Kp = h->Kp; p.G = h->G;
    const size_t smem = ((size_t)K * h->Kp + (size_t)CFN_WARPS * 2 * h->Kp) * sizeof(double);
    CUDA_TRY(cudaFuncSetAttribute(cfn_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaFuncSetAttribute(cfn_chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>((nchains + CFN_WARPS - 1) / CFN_WARPS, (int64_t)h->num_sms * 8);
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    if (K <= 32) cfn_chain_kernel<1><<<grid, 32 * CFN_WARPS, smem, st>>>(p, (const CfnChain *)b[6].p, nchains, (double *)b[10].p);
    else cfn_chain_kernel<2><<<grid, 32 * CFN_WARPS, smem, st>>>(p, (const CfnChain *)b[6].p, nchains, (double *)b[10].p);
    g_launches++;
    if (ntab > 0) {
        const int64_t *d_plist = (const int64_t *)b[8].p, *d_toff = d_plist + plist.size();
        const int32_t *d_t1 = (const int32_t *)b[7].p, *d_t2 = d_t1 + ntab;
        cfn_accumulate_kernel<<<ntab, std::min(K * K, 1024), 0, st>>>((const double *)b[10].p, d_plist, d_toff, d_t1, d_t2, K, k, (double *)b[14].p);
        g_launches++;
    }
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    CUDA_TRY(cudaGetLastError());
    std::vector<double> tail((size_t)2 * K);
    CUDA_TRY(cudaMemcpyAsync(cost_tables, b[14].p, sizeof(double) * (size_t)k * k * K * K, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(tail.data(), (double *)b[10].p + out_start, sizeof(double) * 2 * K, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (device_ms_out) { float ms = 0.f; if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) *device_ms_out = ms; else { cudaGetLastError(); *device_ms_out = 0.0; } }

    // cfn.rs:139-145: unary costs of the first / last boundary's component (+= onto 0.0)
    for (int64_t e = 0; e < (int64_t)k * K; e++) unary[e] = 0.0;
    for (int n = 0; n < K; n++) unary[(int64_t)bc[0] * K + n] += tail[n];
    for (int n = 0; n < K; n++) unary[(int64_t)bc[nb - 1] * K + n] += has_tail ? tail[K + n] : 0.0;
    // cfn.rs:149-166: lower bound = -1 + sum over k1 < k2 of the table's minimum; -inf unary costs take it
    double lb = -1.0;
    for (int k1 = 0; k1 < k; k1++)
        for (int k2 = k1 + 1; k2 < k; k2++) {
            const double *tb = cost_tables + ((size_t)k1 * k + k2) * K * K;
            double m = tb[0];
            for (int e = 1; e < K * K; e++) if (tb[e] < m) m = tb[e];
            lb += m;
        }
    for (int64_t e = 0; e < (int64_t)k * K; e++) if (unary[e] == -std::numeric_limits<double>::infinity()) unary[e] = lb;
    if (lower_bound_out) *lower_bound_out = lb;
    if (nboundaries_out) *nboundaries_out = nb;
    return CV_OK;
}
