// decode_large.cuh -- batched plain Viterbi (mode R1) for K > 64 (tiled logA). Placeholder until implemented.
#pragma once
#include "common.cuh"
