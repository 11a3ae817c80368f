// Helpers around the hot path (reference include/utils/utils.h, src/utils/utils.cpp).
#ifndef CALS_B200_UTILS_UTILS_H
#define CALS_B200_UTILS_UTILS_H

#include <string>
#include <vector>

#include "ktensor.h"

namespace cals::utils {

std::string mode_string(std::vector<dim_t> const &modes); // "I0-I1-I2"

// Column-wise concatenation of models of equal rank (reference src/utils/utils.cpp:18-38).
cals::Ktensor concatenate_ktensors(std::vector<cals::Ktensor> const &ktensors);

// One flagged copy of `reference_ktensor` per mode-0 sample (reference src/utils/utils.cpp:40-51).
void generate_jk_ktensors(cals::Ktensor const &reference_ktensor, std::vector<cals::Ktensor> &jk_ktensor_v);

// Re-order the components of every leave-one-out model to match `ktensor`: linear sum assignment on
// B^T B_m + C^T C_m (reference src/utils/utils.cpp:53-101).
void jk_permutation_adjustment(cals::Ktensor &ktensor, std::vector<cals::Ktensor> &jk_ktensor_v);

// ||X||_jk[i] = sqrt(||X||^2 - sum_j X(i, j)^2) for every mode-0 index, computed on the device in one pass over X
// (reference src/utils/utils.cpp:103-152).
std::vector<double> calculate_jackknifing_norms(cals::Tensor const &tensor);

// Linear sum assignment (square, dense, row-major cost): col_of_row[i] = column assigned to row i.
void linear_sum_assignment(dim_t n, const double *cost_row_major, bool maximize, std::vector<int64_t> &col_of_row);

} // namespace cals::utils

namespace cals::ops {
void update_gramian(const cals::Matrix &factor, cals::Matrix &gramian);             // gramian = F^T F
void update_gramians(const cals::Ktensor &ktensor, std::vector<Matrix> &gramians);
Matrix &hadamard_but_one(std::vector<cals::Matrix> &matrices, dim_t mode);          // into matrices[mode]
void hadamard_all(std::vector<cals::Matrix> &matrices);                             // into matrices[0]
} // namespace cals::ops
#endif
