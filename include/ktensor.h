// cals::Ktensor -- one CP model: N factor matrices (I_n x R) and R weights (reference include/ktensor.h:24-341,
// src/ktensor.cpp).  Host-side object with the reference's interface; while cals::cp_cals runs, the model's columns
// live in the device multi-factor buffers owned by the engine (device-side replacement of MultiKtensor::add/remove,
// reference src/multi_ktensor.cpp:41-163) and the object is overwritten with the fitted values when it is evicted.
#ifndef CALS_B200_KTENSOR_H
#define CALS_B200_KTENSOR_H

#include <atomic>
#include <cmath>
#include <vector>

#include "cals_blas.h"
#include "matrix.h"

using std::vector;

namespace cals {

namespace detail {
int next_ktensor_id(); // process-wide counter (atomic, unlike reference include/ktensor.h:14)
}

struct JackKniffing {
  bool enabled{false};
  dim_t fiber{0};
  dim_t mode{0};
};

class Ktensor {
  int id{-1};
  dim_t components{0};
  dim_t iters{0};
  double fit{0.0};
  double old_fit{0.0};
  double approx_error{0.0};
  bool normalized{false};
  JackKniffing jk{};
  vector<vector<vector<bool>>> active_set; // NNLS: [mode][row][column] constrained-to-zero flags, warm start of the next call
  vector<dim_t> modes{};
  vector<double> lambda{};
  vector<Matrix> factors{};

  void init_active_sets();

public:
  Ktensor() = default;
  ~Ktensor() = default;

  Ktensor(dim_t components, const vector<dim_t> &modes);
  Ktensor(dim_t components, const vector<dim_t> &modes, dim_t jk_fiber, dim_t jk_mode = 0)
      : Ktensor(components, modes) {
    to_jk(jk_mode, jk_fiber);
  }

  Ktensor(Ktensor &&) = default;
  Ktensor &operator=(Ktensor &&) = default;
  Ktensor(const Ktensor &rhs);            // new id; factors, lambda and jk flag copied (reference :71-85)
  Ktensor &operator=(const Ktensor &rhs); // idem (reference :87-102)

  // -- getters / setters (reference include/ktensor.h:106-160) ----------------------------------------------------
  [[nodiscard]] dim_t get_components() const noexcept { return factors[0].get_cols(); }
  [[nodiscard]] dim_t get_iters() const noexcept { return iters; }
  [[nodiscard]] int get_id() const noexcept { return id; }
  [[nodiscard]] bool is_jk() const noexcept { return jk.enabled; }
  [[nodiscard]] dim_t get_jk_mode() const noexcept { return jk.mode; }
  [[nodiscard]] dim_t get_jk_fiber() const noexcept { return jk.fiber; }
  [[nodiscard]] double get_approximation_error() const noexcept { return approx_error; }
  [[nodiscard]] vector<dim_t> const &get_modes() const noexcept { return modes; }
  [[nodiscard]] dim_t get_n_modes() const noexcept { return factors.size(); }
  [[nodiscard]] vector<double> const &get_lambda() const noexcept { return lambda; }
  vector<double> &get_lambda() noexcept { return lambda; }
  [[nodiscard]] vector<Matrix> const &get_factors() const noexcept { return factors; }
  vector<Matrix> &get_factors() noexcept { return factors; }
  [[nodiscard]] Matrix const &get_factor(dim_t mode) const noexcept { return factors.at(mode); }
  Matrix &get_factor(dim_t mode) noexcept { return factors.at(mode); }
  [[nodiscard]] Matrix const &get_last_factor() const noexcept { return factors.back(); }
  vector<vector<bool>> &get_active_set(dim_t mode) noexcept { return active_set.at(mode); }

  void set_iters(dim_t v) noexcept { iters = v; }
  void set_approximation_error(double v) noexcept { approx_error = v; }
  void set_factor(int index, double *src) noexcept {
    Matrix &f = get_factor(static_cast<dim_t>(index));
    std::copy(src, src + f.get_n_elements(), f.get_data());
  }
  void set_lambda(double const *src) noexcept { std::copy(src, src + lambda.size(), lambda.begin()); }

  // fit bookkeeping (reference include/ktensor.h:178-189)
  double calculate_new_fit(double X_norm) noexcept {
    old_fit = fit;
    fit = 1 - std::fabs(approx_error) / X_norm;
    return fit;
  }
  [[nodiscard]] double get_fit_diff() const noexcept { return std::fabs(old_fit - fit); }
  [[nodiscard]] double get_fit() const noexcept { return fit; }         // extension
  [[nodiscard]] double get_old_fit() const noexcept { return old_fit; } // extension
  void set_fit(double new_fit, double prev_fit) noexcept {              // extension: results coming from the device
    fit = new_fit;
    old_fit = prev_fit;
  }

  void print(const std::string &&text = "Ktensor") const;

  // -- storage redirection (reference src/ktensor.cpp:109-135) ------------------------------------------------------
  Ktensor &attach(vector<double *> &data_ptrs, bool multi_thread = true); // copy factors there and view them
  Ktensor &detach();                                                      // copy back, zero the foreign block

  // -- normalisation (reference src/ktensor.cpp:66-107) -------------------------------------------------------------
  Ktensor &normalize();                               // every column to unit 2-norm, lambda = product of norms
  Ktensor &normalize(dim_t mode, dim_t iteration = 1); // 2-norm (iteration 1) or signed max-magnitude entry
  Ktensor &denormalize();                             // factor 0 *= lambda

  Ktensor &randomize();                   // uniform(-1,1), then normalize() (reference src/ktensor.cpp:10-17)
  Ktensor &fill(function<double()> &&f);  // factors 0..N-1 column-major from f, then normalize() (:19-28)

  Tensor to_tensor(); // dense reconstruction, mode 0 fastest (reference src/ktensor.cpp:30-64)

  Ktensor &copy(Ktensor &rhs); // values only, keeps the id (reference src/ktensor.cpp:162-179)

  // -- jackknife (reference include/ktensor.h:276-325) --------------------------------------------------------------
  Ktensor &to_jk(dim_t mode, dim_t fiber) {
    jk = JackKniffing{true, fiber, mode};
    return *this;
  }
  Ktensor to_regular(); // drop the jackknifed row: mode jk.mode shrinks by one
  void set_jk_fiber(double value) noexcept {
    if (!jk.enabled)
      return;
    Matrix &f = get_factor(jk.mode);
    for (dim_t j = 0; j < f.get_cols(); j++)
      f(jk.fiber, j) = std::isnan(value) ? NAN : f(jk.fiber, j) * value;
  }
};

} // namespace cals
#endif
