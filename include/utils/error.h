// Approximation-error helpers (reference include/utils/error.h, src/utils/error.cpp).  Inside cp_cals the fast error
// is evaluated on the device (model_update_kernel); these host versions exist for callers and tests that check a
// fitted model afterwards.
#ifndef CALS_B200_UTILS_ERROR_H
#define CALS_B200_UTILS_ERROR_H

#include <vector>

#include "ktensor.h"

namespace cals::error {

// sqrt(max(||X||^2 + sum_ij l_i l_j P_ij - 2 sum_ij l_j F_ij G_ij, 0))      (reference src/utils/error.cpp:64-89)
double compute_fast_error(double X_norm, const std::vector<double> &lambda, const cals::Matrix &last_factor,
                          const cals::Matrix &last_mttkrp, const cals::Matrix &gramian_hadamard);

// ||X - [[lambda; A, B, C, ..]]||_F by explicit reconstruction (reference src/utils/error.cpp:7-31; any N here).
// The workspaces are accepted for signature parity and not used.
double compute_error(const cals::Tensor &X, cals::Ktensor &ktensor, cals::Matrix &krp_workspace,
                     cals::Matrix &ten_workspace);

} // namespace cals::error
#endif
