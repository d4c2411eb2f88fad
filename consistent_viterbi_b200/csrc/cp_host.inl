// host control loop of the constrained solver (included by cv_api.cu)
extern "C" int cv_cp_solve(cv_hmm *h, const uint32_t *, const uint8_t *, const int32_t *, int64_t, int32_t, uint64_t,
                           uint64_t *, double *, uint64_t *, uint64_t *)
{
    (void)h;
    return fail(CV_ERR_UNSUPPORTED, "cv_cp_solve not built yet");
}
extern "C" int cv_cp_last_state(cv_hmm *, double *, uint64_t *) { return fail(CV_ERR_UNSUPPORTED, "not built yet"); }
extern "C" int cv_cp_last_ub(cv_hmm *, double *, uint64_t, uint64_t *) { return fail(CV_ERR_UNSUPPORTED, "not built yet"); }
