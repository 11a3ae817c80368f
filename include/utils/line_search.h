// Line-search options (reference include/utils/line_search.h:8-13).  Line search is off by default in the reference
// (include/cals.h:153) and is not on the B200 path: cp_cals / cp_als refuse to run with it switched on.
#ifndef CALS_B200_UTILS_LINE_SEARCH_H
#define CALS_B200_UTILS_LINE_SEARCH_H

#include <string>

namespace cals::ls {
enum LS_METHOD { NO_ERROR_CHECKING = 0, ERROR_CHECKING_SERIAL, ERROR_CHECKING_PARALLEL, LENGTH };
static const std::string ls_method_names[LS_METHOD::LENGTH] = {"no-error-checking", "error-checking-serial",
                                                               "error-checking-parallel"};
} // namespace cals::ls
#endif
