// host launch logic of the large-K kernel (included by cv_api.cu)
static int launch_decode_large(cv_hmm *h, const uint32_t *, const int64_t *, int64_t, int64_t, uint32_t *, double *,
                               const uint32_t *, const uint32_t *, unsigned int *, int *, cudaStream_t)
{
    return fail(CV_ERR_UNSUPPORTED, "K=%d > %d: large-K kernel not built yet", h->K, SMALL_K_MAX);
}
