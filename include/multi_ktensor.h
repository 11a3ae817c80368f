// cals::MultiKtensor -- the column-concatenated buffer of all concurrently fitted models (reference
// include/multi_ktensor.h:12-130, src/multi_ktensor.cpp).  It IS a Ktensor whose factor n is the I_n x (active columns)
// multi-factor; models are attached into it (their Matrix views point into the buffer) and detached with their final
// values.  Same public surface and semantics as the reference: first-fit placement, BufferFull, stable compaction.
//
// On the B200 path cp_cals does not use this host class -- the same bookkeeping runs on the device (ModelDesc /
// sched_kernel, cp-cals_b200/csrc/sched.cuh).  It is kept for callers that drive the pieces themselves, as the
// reference's MTTKRP benchmarks do: mttkrp::mttkrp(X, multi_ktensor, ws, mode, params) runs the concurrent MTTKRP of
// all attached models in one launch of the tensor-core kernel.
#ifndef CALS_B200_MULTI_KTENSOR_H
#define CALS_B200_MULTI_KTENSOR_H

#include <exception>
#include <map>

#include "ktensor.h"
#include "utils/line_search.h"

namespace cals {

struct RegistryEntry {
  Ktensor &ktensor;        // the attached model
  vector<Matrix> gramians; // one R x R Gramian per mode, of the factors as they were when the model was added
  int col;                 // first buffer column of the model
  dim_t id;                // key in the registry
  ls::LineSearchParams ls_params{};
};

typedef std::map<int, RegistryEntry> Registry;

struct BufferFull : public std::exception {
  [[nodiscard]] const char *what() const noexcept override {
    return "Buffer is full, wait until some ktensors converge.";
  }
};

class MultiKtensor : public Ktensor {
  vector<dim_t> tensor_modes;
  vector<dim_t> owner; // per buffer column: id of the model living there, 0 = free
  Registry registry;
  int start{0};
  bool cuda{false};
  bool line_search{false};
  ls::LineSearchParams ls_params{};
  bool flag_jk{false};
  dim_t next_id{1};

  int first_fit(dim_t width) const; // leftmost run of `width` free columns; throws BufferFull
  void refresh_views();             // factor n := buffer columns [0, last occupied]

public:
  MultiKtensor() = default;
  ~MultiKtensor() = default;
  MultiKtensor &operator=(MultiKtensor &&) = default;

  explicit MultiKtensor(vector<dim_t> &modes, dim_t buffer_size);

  MultiKtensor &add(Ktensor &ktensor);    // throws BufferFull when no contiguous gap is wide enough
  MultiKtensor &remove(dim_t ktensor_id); // the model gets its columns back, the buffer columns are zeroed
  MultiKtensor &compress();               // shift every model left over the holes, order preserved

  Registry &get_registry() { return registry; }
  [[nodiscard]] int get_start() const noexcept { return start; }
  [[nodiscard]] bool get_flag_jk() const noexcept { return flag_jk; }
  void set_cuda(bool value) { cuda = value; }
  void set_line_search(bool value) { line_search = value; }
  void set_line_search_params(ls::LineSearchParams &params) { ls_params = params; }
  vector<dim_t> &get_modes() { return tensor_modes; }
  int get_leftmost_id() { return owner.empty() ? -1 : static_cast<int>(owner[0]); }
};

} // namespace cals
#endif
