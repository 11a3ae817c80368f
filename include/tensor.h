// cals::Tensor -- dense FP64 tensor, column-major with mode 0 fastest (reference include/tensor.h:48-310,
// src/tensor.cpp:18-196).  Same public surface as the reference class; differences that matter on the B200 path:
//   * storage of large tensors is page-locked (cals_b200_host_alloc) so the upload to HBM runs at PCIe/C2C speed;
//   * there is no device mirror inside the object (the reference's cudata/cudata_up): device residency belongs to
//     the engine context behind the C ABI (include/cals_b200.h), which also computes norm() there when it can;
//   * copying does not print (reference src/tensor.cpp:98 does -- SURVEY appendix B).
#ifndef CALS_B200_TENSOR_H
#define CALS_B200_TENSOR_H

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <functional>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "cals_blas.h"
#include "definitions.h"

using std::function;
using std::unique_ptr;
using std::vector;

namespace cals {

namespace detail {
// Host storage: 64-byte aligned; blocks of >= 1 MiB are page-locked when a CUDA device is present.
double *host_alloc(dim_t n_doubles);
void host_free(double *p);
struct HostFree {
  void operator()(double *p) const { host_free(p); }
};
uint64_t next_tensor_uid();
} // namespace detail

// Geometry of the mode-n unfolding as a sequence of equally sized matrix blocks (reference include/tensor.h:38-44).
struct Unfolding {
  dim_t n_blocks;
  dim_t block_offset;
  dim_t rows;
  dim_t cols;
  dim_t stride;
};

class Tensor {
  int rank{0};
  dim_t n_elements{0};
  dim_t max_n_elements{0};
  vector<dim_t> modes;
  unique_ptr<double, detail::HostFree> data_up{};
  double *data{nullptr};
  // Identity of this object for the engine's "tensor already resident" test (AlsParams::cuda_no_tensor_alloc): a fresh
  // value for every constructed or copy-assigned Tensor, kept by moves, so that a new Tensor that the allocator happens
  // to place at the address of a freed one is never mistaken for it.
  uint64_t uid{detail::next_tensor_uid()};

  void allocate(dim_t n);

public:
  Tensor() = default;
  ~Tensor() = default;

  explicit Tensor(const vector<dim_t> &modes);                      // uninitialised
  explicit Tensor(const vector<dim_t> &modes, double *view_data);   // non-owning view
  explicit Tensor(const std::string &file_name);                    // text: "I0 I1 ..\n" then values, mode 0 fastest
  Tensor(dim_t mode0, dim_t mode1, double *view_data = nullptr);    // matrices
  Tensor(dim_t rank, const vector<dim_t> &modes);                   // random rank-`rank` tensor

  Tensor(Tensor &&rhs) = default;
  Tensor &operator=(Tensor &&rhs) = default;
  Tensor(const Tensor &rhs);            // deep copy unless rhs is a view
  Tensor &operator=(const Tensor &rhs);

  [[nodiscard]] dim_t get_n_elements() const noexcept { return n_elements; }
  [[nodiscard]] dim_t get_max_n_elements() const noexcept { return max_n_elements; }
  [[nodiscard]] dim_t get_n_modes() const noexcept { return modes.size(); }
  [[nodiscard]] vector<dim_t> get_modes() const noexcept { return modes; }
  [[nodiscard]] double *get_data() const noexcept { return data; }
  [[nodiscard]] int get_rank() const noexcept { return rank; }
  [[nodiscard]] bool is_view() const noexcept { return data_up == nullptr; }
  [[nodiscard]] uint64_t get_uid() const noexcept { return uid; }

  void set_data(double *new_data) noexcept {
    data = new_data;
    uid = detail::next_tensor_uid();
  }
  Tensor &reset_data() noexcept {
    data = data_up.get();
    uid = detail::next_tensor_uid();
    return *this;
  }

  double const &operator[](dim_t i) const noexcept { return data[i]; }
  double &operator[](dim_t i) noexcept { return data[i]; }

  // "Soft" resize inside the allocation (reference include/tensor.h:186-192).
  void resize(dim_t new_n_elements, vector<dim_t> &new_modes) {
    assert(new_n_elements <= max_n_elements && new_modes.size() == modes.size());
    n_elements = new_n_elements;
    modes = std::move(new_modes);
  }

  // Frobenius norm, scaled accumulation like dnrm2 (reference include/tensor.h:196).
  [[nodiscard]] double norm() const;

  Tensor &fill(const function<double()> &&f);
  Tensor &zero();
  Tensor &randomize(); // uniform(-1, 1), std::mt19937 seeded from std::random_device (reference src/tensor.cpp:121-129)

  void copy(const Tensor &ten) noexcept {
    std::copy(ten.data, ten.data + ten.n_elements, data);
    uid = detail::next_tensor_uid();
  }

  dim_t max_id(vector<bool> &mask) noexcept {
    dim_t best = 0;
    double mx = -DBL_MAX;
    for (dim_t i = 0; i < mask.size(); i++)
      if (mask[i] && data[i] > mx) {
        mx = data[i];
        best = i;
      }
    return best;
  }
  double max(vector<bool> &mask) noexcept { return data[max_id(mask)]; }
  double min() noexcept { return *std::min_element(data, data + n_elements); }

  void print(const std::string &&text = "Tensor") const;

  [[nodiscard]] Unfolding implicit_unfold(dim_t mode) const;

  void set_rank(int r) noexcept { rank = r; } // extension (the reference sets it only in the rank constructor)
};

} // namespace cals
#endif
