// cv_cp.cu -- constrained decode (CPSolver, reference src/viterbi_solver/cp.rs:20-152): C ABI entry points
// cv_cp_solve / cv_cp_solve_dist / cv_cp_dist_* and the parity hooks, over cp_kernels.cuh and cp_dist.cuh; and the
// CFN cost tables (cv_cfn_tables, src/viterbi_solver/cfn.rs:11-167), which reuse the same sweep kernels.
#include "cv_internal.cuh"

#include <chrono>

#include "common.cuh"
#include "cp_kernels.cuh"
#include "cp_dist.cuh"
#include "cfn.cuh"

using namespace cvb;

#include "cp_host.inl"
#include "cfn_host.inl"
