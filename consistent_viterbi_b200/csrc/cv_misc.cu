// cv_misc.cu -- SURVEY.md 8(f) N3: supervised MLE counts (cv_mle, reference src/hmm/hmm.rs:30-62,192-205).
#include "cv_internal.cuh"

#include <chrono>

#include "common.cuh"
#include "mle.cuh"

using namespace cvb;

#include "mle_host.inl"
