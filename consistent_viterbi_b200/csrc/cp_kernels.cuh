// cp_kernels.cuh -- device kernels of the constrained (CPSolver) path. Placeholder until implemented.
#pragma once
#include "common.cuh"
