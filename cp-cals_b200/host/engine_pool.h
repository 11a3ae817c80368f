// Internal to the C++ host layer: one C-ABI context (include/cals_b200.h) per CUDA device, created on first use and
// kept for the life of the process so that device allocations, TMA descriptors and pinned staging buffers are reused
// from call to call (the reference allocates its device mirrors per call, reference src/cals.cpp:142-155).
#ifndef CALS_B200_HOST_ENGINE_POOL_H
#define CALS_B200_HOST_ENGINE_POOL_H

#include <mutex>
#include <string>
#include <vector>

#include "cals.h"
#include "cals_b200.h"

namespace cals::detail {

struct EngineHandle {
  cals_b200_ctx *ctx{nullptr};
  int device{0};
  std::mutex mu; // one caller at a time per device (the reference API is not re-entrant either)
  // what is resident: identity (Tensor::get_uid), address, extents and a sampled content fingerprint of the caller's
  // tensor at the time of the last upload
  const double *resident_data{nullptr};
  std::vector<dim_t> resident_modes;
  uint64_t resident_uid{0}, resident_fp{0};
};

// Throws cals::B200Error when the device or the library is unusable (no CPU fallback).
EngineHandle &engine_for_device(int device);

// rc != 0  ->  throw B200Error("<what>: <last error of the context>")
void check(EngineHandle &e, int rc, const char *what);

// Upload X unless `may_skip` and the same buffer with the same extents is already resident.
void upload_tensor(EngineHandle &e, const Tensor &X, bool may_skip);

struct RunOptions {
  dim_t buffer_size{4200};
  dim_t max_iterations{200};
  double tol{1e-7};
  bool force_max_iter{false};
  bool always_evict_first{false};
  bool nnls{false}; // update_method == NNLS
  bool line_search{false};
  int ls_method{0};
  int ls_interval{5};
  double ls_step{0.0};
  int timing{0};
  bool pair_node{true}; // mttkrp_method != MTTKRP: modes 1 and 2 of a 3-mode tensor share one contraction
  bool skip_upload_if_resident{false};
};

struct RunResult {
  cals_b200_report rep{};
  std::vector<dim_t> cols; // active columns per global iteration
  uint64_t ls_performed{0}, ls_failed{0};
};

// Fit `models` (FIFO order) to X on one device; every Ktensor is overwritten with its result.
RunResult run_on_device(int device, const Tensor &X, const std::vector<Ktensor *> &models, const RunOptions &opt);

// Deterministic split of a FIFO model list over n_parts shards: every model goes to the shard with the smallest sum of
// ranks so far (ties: lowest shard index); order inside a shard is the queue order.
std::vector<std::vector<size_t>> shard_models(const std::vector<dim_t> &ranks, size_t n_parts);

} // namespace cals::detail
#endif
