// Line-search options and the per-model state that goes with them (counterpart of the reference's
// include/utils/line_search.h).  The search itself runs on the device (cp-cals_b200/csrc/ls.cuh); the struct below only
// exists on the host because MultiKtensor's registry entries carry one per model (include/multi_ktensor.h).
#ifndef CALS_B200_UTILS_LINE_SEARCH_H
#define CALS_B200_UTILS_LINE_SEARCH_H

#include <string>

#include "ktensor.h"

namespace cals::ls {

// Values as in the reference (they are stored in CalsParams / AlsParams and printed by name).
enum LS_METHOD {
  NO_ERROR_CHECKING = 0,   // extrapolate blindly, return to a backup if the error went up       -> device
  ERROR_CHECKING_SERIAL,   // extrapolate a trial model, keep it only if its explicit error is lower -> device
  ERROR_CHECKING_PARALLEL, // declared upstream but never dispatched there; rejected here
  LENGTH
};

static const std::string ls_method_names[LS_METHOD::LENGTH] = {
    "no-error-checking",
    "error-checking-serial",
    "error-checking-parallel",
};

struct LineSearchParams {
  // configuration, copied from CalsParams / AlsParams when a model enters the buffer
  LS_METHOD method{NO_ERROR_CHECKING};
  int interval{};            // iterations between two extrapolations
  double step{0.0};          // extrapolation factor; 0 means cbrt(iteration)
  bool cuda{false};          // unused on this path (everything is on the device)
  Tensor const *T{nullptr};  // target tensor, for the explicit error of the error-checking methods

  // running state
  int iter{};                    // iterations since the last extrapolation
  bool updated_last_iter{};      // the previous iteration extrapolated (NO_ERROR_CHECKING)
  bool extrapolated{false};      // outcome flags of the last call
  bool reversed{false};

  // models kept on the side
  Ktensor prev_ktensor{};   // snapshot taken interval-1 iterations after the last extrapolation
  Ktensor backup_ktensor{}; // NO_ERROR_CHECKING: state right before the extrapolation
};

} // namespace cals::ls
#endif
