// Line-search options and per-model state (reference include/utils/line_search.h:8-39).  Line search is off by default
// in the reference (include/cals.h:153) and is not on the B200 path: cp_cals / cp_als refuse to run with it switched
// on.  The state struct exists because MultiKtensor's registry entries carry one (reference include/multi_ktensor.h:20).
#ifndef CALS_B200_UTILS_LINE_SEARCH_H
#define CALS_B200_UTILS_LINE_SEARCH_H

#include <string>

#include "ktensor.h"

namespace cals::ls {
enum LS_METHOD { NO_ERROR_CHECKING = 0, ERROR_CHECKING_SERIAL, ERROR_CHECKING_PARALLEL, LENGTH };
static const std::string ls_method_names[LS_METHOD::LENGTH] = {"no-error-checking", "error-checking-serial",
                                                               "error-checking-parallel"};

struct LineSearchParams {
  int iter{};
  int interval{};
  bool updated_last_iter{};
  LS_METHOD method{NO_ERROR_CHECKING};
  Ktensor prev_ktensor{};
  double step{0.0};
  bool cuda{false};
  bool extrapolated{false};
  bool reversed{false};
  Ktensor backup_ktensor{};
  Tensor const *T{nullptr};
};
} // namespace cals::ls
#endif
