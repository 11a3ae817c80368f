// Factor-update options (reference include/utils/update.h:6-20).  Both run inside model_update_kernel
// (cp-cals_b200/csrc/update.cuh): UNCONSTRAINED = batched Cholesky solve, NNLS = row-wise active-set solve with the
// active sets kept in Ktensor::active_set between calls, as in the reference (src/utils/update.cpp:61-176).
#ifndef CALS_B200_UTILS_UPDATE_H
#define CALS_B200_UTILS_UPDATE_H

#include <string>

namespace cals::update {
enum UPDATE_METHOD { UNCONSTRAINED = 0, NNLS, LENGTH };
static const std::string update_method_names[UPDATE_METHOD::LENGTH] = {"unconstrained", "nnls"};
} // namespace cals::update
#endif
