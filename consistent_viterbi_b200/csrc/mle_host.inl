// mle_host.inl -- cv_mle: HMM::maximum_likelihood_estimation + HMM::log (reference src/hmm/hmm.rs:30-62,192-205)
// (included by cv_api.cu).  The device counts the events (mle.cuh); the host turns counts into the values the
// reference's `+= 1.0` loop would have produced, divides and takes logarithms exactly as the reference does.

namespace {

// v after `n` sequential `v += 1.0` (IEEE binary64, round to nearest even).  Inside a binade where v + 1.0 is
// exact every further +1 is exact too, so the run up to the next power of two is one exact add; only the first
// add after entering a binade (or below 1.0) can round, and that one is done as the reference does it.
double add_ones(double v, unsigned long long n)
{
    while (n > 0) {
        if (!(v >= 0.0) || !std::isfinite(v)) { v = v + 1.0; n--; continue; }       // negative / NaN / inf: literal
        const double w = v + 1.0;
        if (v >= 1.0 && w - 1.0 == v && v < 9007199254740992.0) {
            int ex; std::frexp(v, &ex);                                              // v in [2^(ex-1), 2^ex)
            const double P = std::ldexp(1.0, ex);
            const double room = P - v;                                               // exact (Sterbenz), integer-valued
            const unsigned long long m = (unsigned long long)std::min<double>((double)n, room);
            if (m == 0) { v = w; n--; continue; }
            v = v + (double)m;                                                       // exact: every partial sum is representable
            n -= m;
        } else {
            v = w; n--;
        }
    }
    return v;
}

}  // namespace

extern "C" int cv_mle(int K, int D, const uint64_t *bdims, double *a, double *b, double *pi, const uint32_t *obs_flat,
                      const int32_t *tags_flat, const int64_t *seq_off, int64_t B, int device, double *count_ms_out)
{
    if (K <= 0 || D <= 0 || !bdims || !a || !b || !pi) return fail(CV_ERR_ARG, "bad model argument");
    if (B < 0 || (B > 0 && (!obs_flat || !tags_flat || !seq_off))) return fail(CV_ERR_ARG, "bad sequence argument");
    int64_t M = 1;
    for (int d = 0; d < D; d++) M *= (int64_t)bdims[d];
    if (M <= 0) return fail(CV_ERR_ARG, "empty observation space");
    int rc = check_device(device < 0 ? 0 : device);
    if (rc) return rc;
    if (device >= 0) CUDA_TRY(cudaSetDevice(device));
    const int64_t N = B ? seq_off[B] : 0;
    for (int64_t i = 0; i < B; i++)
        if (seq_off[i + 1] <= seq_off[i]) return fail(CV_ERR_EMPTY, "sequence %lld is empty (reference: tag[0] / len()-1 panics, hmm.rs:39-40)", (long long)i);
    if (B > 0x7fffffffLL) return fail(CV_ERR_UNSUPPORTED, "more than 2^31 sequences");

    const size_t ncnt = (size_t)K * K + (size_t)K * M + 3 * (size_t)K;
    DevBuf cnt, dobs, dtags, doff, dseq, dstat;
    auto cleanup = [&]() { for (DevBuf *x : {&cnt, &dobs, &dtags, &doff, &dseq, &dstat}) x->release(); };
    std::vector<unsigned long long> hc(ncnt, 0ULL);
    int status = 0;
    float ms = 0.f;
    if (N > 0) {
        if ((rc = cnt.ensure(ncnt * 8)) || (rc = dobs.ensure((size_t)N * 4)) || (rc = dtags.ensure((size_t)N * 4)) ||
            (rc = doff.ensure((size_t)(B + 1) * 8)) || (rc = dseq.ensure((size_t)N + 16)) || (rc = dstat.ensure(16))) { cleanup(); return rc; }
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaError_t ce = cudaMemset(cnt.p, 0, ncnt * 8);
        if (ce == cudaSuccess) ce = cudaMemset(dstat.p, 0, 16);
        if (ce == cudaSuccess) ce = cudaMemcpy(dobs.p, obs_flat, (size_t)N * 4, cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMemcpy(dtags.p, tags_flat, (size_t)N * 4, cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMemcpy(doff.p, seq_off, (size_t)(B + 1) * 8, cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) {
            MleParams p;
            p.obs = (const uint32_t *)dobs.p; p.tags = (const int32_t *)dtags.p; p.seq_off = (const int64_t *)doff.p;
            p.first = (const uint8_t *)dseq.p; p.N = N; p.B = B; p.M = M; p.K = K;
            p.a = (unsigned long long *)cnt.p; p.b = p.a + (size_t)K * K; p.pi = p.b + (size_t)K * M;
            p.seen = p.pi + K; p.end = p.seen + K; p.status = (int *)dstat.p;
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device < 0 ? 0 : device);
            cudaEventRecord(e0);
            cudaMemsetAsync(dseq.p, 0, (size_t)N + 1);
            mle_mark_kernel<<<(unsigned)((B + 1 + 255) / 256), 256>>>(p.seq_off, B, (uint8_t *)dseq.p);
            const size_t smem = K <= MLE_SMEM_K ? ((size_t)K * K + 3 * K) * sizeof(unsigned int) : 0;
            const int grid = (int)std::min<int64_t>((N + 4 * MLE_THREADS - 1) / (4 * MLE_THREADS), (int64_t)sms * 8);
            mle_count_kernel<<<grid, MLE_THREADS, smem>>>(p);
            cudaEventRecord(e1);
            g_launches += 2;
            ce = cudaGetLastError();
        }
        if (ce == cudaSuccess) ce = cudaMemcpy(hc.data(), cnt.p, ncnt * 8, cudaMemcpyDeviceToHost);
        if (ce == cudaSuccess) ce = cudaMemcpy(&status, dstat.p, sizeof(int), cudaMemcpyDeviceToHost);
        if (ce == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        cleanup();
        if (ce != cudaSuccess) { cudaGetLastError(); return fail(CV_ERR_CUDA, "cv_mle: %s", cudaGetErrorString(ce)); }
    }
    if (count_ms_out) *count_ms_out = ms;
    if (status & 1) return fail(CV_ERR_ARG, "a tag is None or >= K (reference: tag[t].unwrap() / index panic, hmm.rs:39-47)");
    if (status & 2) return fail(CV_ERR_ARG, "observation index >= M (reference: ndarray index panic, hmm.rs:41)");

    const unsigned long long *ca = hc.data(), *cb = ca + (size_t)K * K, *cpi = cb + (size_t)K * M, *cseen = cpi + K, *cend = cseen + K;
    // hmm.rs:39-47 replayed per entry, then :50-59
    for (size_t e = 0; e < (size_t)K * K; e++) a[e] = add_ones(a[e], ca[e]);
    for (size_t e = 0; e < (size_t)K * M; e++) b[e] = add_ones(b[e], cb[e]);
    for (int s = 0; s < K; s++) pi[s] = add_ones(pi[s], cpi[s]);
    for (int s = 0; s < K; s++) {
        const double seen = (double)cseen[s], end = (double)cend[s];                 // sums of 1.0 from 0.0: exact below 2^53
        if (seen != end) { const double d = seen - end; for (int j = 0; j < K; j++) a[(size_t)s * K + j] /= d; }
        else for (int j = 0; j < K; j++) a[(size_t)s * K + j] = 0.0;
        pi[s] /= (double)B;
        for (int64_t m = 0; m < M; m++) b[(size_t)s * M + m] /= seen;
    }
    // HMM::log, hmm.rs:192-205: x == 0 -> -inf, else x.log(10.0) = ln(x) / ln(10)
    const double ln10 = std::log(10.0);
    auto lg = [&](double x) { return x == 0.0 ? -std::numeric_limits<double>::infinity() : std::log(x) / ln10; };
    for (size_t e = 0; e < (size_t)K * K; e++) a[e] = lg(a[e]);
    for (size_t e = 0; e < (size_t)K * M; e++) b[e] = lg(b[e]);
    for (int s = 0; s < K; s++) pi[s] = lg(pi[s]);
    return CV_OK;
}
