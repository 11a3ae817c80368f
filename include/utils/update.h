// Factor-update options (reference include/utils/update.h:6-20).  Only UNCONSTRAINED exists on the B200 path
// (batched Cholesky solve inside model_update_kernel, cp-cals_b200/csrc/update.cuh); NNLS is rejected loudly.
#ifndef CALS_B200_UTILS_UPDATE_H
#define CALS_B200_UTILS_UPDATE_H

#include <string>

namespace cals::update {
enum UPDATE_METHOD { UNCONSTRAINED = 0, NNLS, LENGTH };
static const std::string update_method_names[UPDATE_METHOD::LENGTH] = {"unconstrained", "nnls"};
} // namespace cals::update
#endif
