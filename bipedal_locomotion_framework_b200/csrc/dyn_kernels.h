// dyn_kernels.h -- launcher interface of dyn_kernels.cu (its own translation unit, so that the
// many unrolled instantiations of the mass-matrix solve compile beside ccm_capi.cu, not inside it).
#pragma once

#include <cuda_runtime.h>

namespace blfccm {

// Last step of System::FloatingBaseDynamicalSystem::dynamics
// (src/System/src/FloatingBaseSystemDynamics.cpp:226-248), n independent systems:
//   rhs = known;  rhs.tail(nc - 6) += joint_torques                       (:226-227)
//   acc = (mass [+ reg]).llt().solve(rhs)                                 (:235-243)
struct LltArgs {
    const double* mass;     // [n][nc][nc] row-major (iDynTree::MatrixDynSize); lower triangle is read
    const double* reg;      // [nc][nc] row-major, shared by all systems, or nullptr
    const double* known;    // [n][nc]
    const double* tau;      // [n][nc - 6] or nullptr
    double* acc;            // [n][nc]; may alias known
    long long n;
    int nc;
};

constexpr int kLltMaxFast = 64;    // warp-level kernel: up to 2 rows x 32 lanes (64 unknowns with the right-hand side as a column)
constexpr int kLltMaxCols = 128;   // block-level kernel

// path_out: lanes per system of the warp-level kernel (4, 8, 16, 32) or 0 for the block-level kernel;
// ncmax_out: the compile-time size class the call ran in.  force_general != 0: block-level kernel
// whatever the size (tests compare the two bit for bit).
cudaError_t llt_solve_launch(const LltArgs& a, cudaStream_t st, bool pdl, int force_general,
                             int* path_out, int* ncmax_out);

}  // namespace blfccm
