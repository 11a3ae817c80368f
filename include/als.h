// Single-model ALS entry points (reference include/als.h:21-221, src/als.cpp).  On the B200 path a single model is
// the C = R special case of the concurrent engine: cp_als enqueues one model with buffer_size = its rank, so every
// kernel is the one cp_cals uses and the two agree bit for bit on the same model.
#ifndef CALS_B200_ALS_H
#define CALS_B200_ALS_H

#include <cstdint>
#include <string>
#include <vector>

#include "ktensor.h"
#include "timer.h"
#include "utils/error.h"
#include "utils/line_search.h"
#include "utils/mttkrp.h"
#include "utils/update.h"
#include "utils/utils.h"

namespace cals {

struct JKTime {
  double pre_als_time{0.0};
  double als_time{0.0};
};

struct JKReport {
  JKTime jk_time{};
  vector<vector<Ktensor>> results; // results[b][i]: base model b with mode-0 sample i left out
};

struct AlsReport {
  int tensor_rank{0};
  dim_t n_modes{0};
  vector<dim_t> modes;
  double X_norm{0.0};

  dim_t iter{0};
  dim_t max_iter{0};
  int n_threads{0};
  int ktensor_id{0};
  dim_t ktensor_components{0};
  double tol{0.0};
  bool cuda{false};
  update::UPDATE_METHOD update_method{update::UNCONSTRAINED};

  bool line_search{false};
  int line_search_interval{0};
  double line_search_step{0.0};
  dim_t ls_performed{0};
  dim_t ls_failed{0};
  ls::LS_METHOD line_search_method{ls::NO_ERROR_CHECKING};

  uint64_t flops_per_iteration{0}; // 2 * N * nX * R

  double total_time{0.0};
  Matrix als_times{};
  Matrix mode_times{};
  Matrix mttkrp_times{};

  // CSV with the reference's column set (reference include/als.h:70-135)
  void print_header(const std::string &file_name, const std::string &sep = ";") const;
  void print_to_file(const std::string &file_name, const std::string &sep = ";") const;
};

struct AlsParams {
  update::UPDATE_METHOD update_method{update::UPDATE_METHOD::UNCONSTRAINED};
  mttkrp::MTTKRP_METHOD mttkrp_method{mttkrp::MTTKRP_METHOD::AUTO};
  cals::mttkrp::MttkrpLut mttkrp_lut{};

  dim_t max_iterations{200};
  double tol{1e-7};
  bool cuda{false}; // accepted; this library always runs on the B200

  bool line_search{false};
  int line_search_interval{5};
  double line_search_step{0};
  ls::LS_METHOD line_search_method{ls::NO_ERROR_CHECKING};

  bool cuda_no_tensor_alloc{false}; // X already resident from the previous call with the same Tensor object: skip the upload
  bool force_max_iter{false};
  bool suppress_lut_warning{false};

  int device{0}; // extension: CUDA device ordinal

  void print() const;
};

AlsReport cp_als(const Tensor &X, Ktensor &ktensor, AlsParams &params);

// Leave-one-out models fitted one by one on the row-deleted tensors (the reference's baseline for jk_cp_cals,
// reference src/als.cpp:291-388).
JKReport jk_cp_als(const Tensor &X, vector<Ktensor> &kt_vector, AlsParams &als_params);

// Loop of cp_als with the tensor uploaded once (reference src/als.cpp:390-419 runs it as an OpenMP loop).
vector<AlsReport> cp_omp_als(const Tensor &X, vector<Ktensor> &ktensor, AlsParams &params);
JKReport jk_cp_omp_als(const Tensor &X, vector<Ktensor> &kt_vector, AlsParams &params);

} // namespace cals
#endif
