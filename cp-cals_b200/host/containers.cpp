// Host containers of the cals:: surface: Tensor, Matrix, Ktensor (reference src/tensor.cpp, src/matrix.cpp,
// src/ktensor.cpp).  Plain C++ -- no BLAS on the host side of the B200 path; the per-column loops below run once per
// call (fill / normalise / reconstruct), never inside the ALS iteration.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <new>
#include <numeric>
#include <sstream>

#include "cals_b200.h"
#include "ktensor.h"
#include "matrix.h"
#include "tensor.h"

// ---------------------------------------------------------------------------------------------------------------------
namespace {
std::atomic<int> g_threads{1};
std::atomic<int> g_ktensor_id{1};
} // namespace

void set_threads(int threads) { g_threads = threads > 0 ? threads : 1; }
int get_threads() { return g_threads; }

namespace cals {
namespace detail {

int next_ktensor_id() { return g_ktensor_id.fetch_add(1); }

// Every block carries a 64-byte header that remembers who allocated it.
namespace {
constexpr uint64_t TAG_PINNED = 0x43414c5350494e21ull, TAG_PAGED = 0x43414c5350414745ull;
constexpr dim_t HEADER_DOUBLES = 8;
constexpr dim_t PIN_THRESHOLD_BYTES = dim_t(1) << 20;
} // namespace

double *host_alloc(dim_t n_doubles) {
  const size_t bytes = (n_doubles + HEADER_DOUBLES) * sizeof(double);
  void *raw = nullptr;
  uint64_t tag = TAG_PAGED;
  if (bytes >= PIN_THRESHOLD_BYTES && (raw = cals_b200_host_alloc(bytes)) != nullptr)
    tag = TAG_PINNED;
  if (!raw)
    raw = operator new[](bytes, std::align_val_t(64));
  *static_cast<uint64_t *>(raw) = tag;
  return static_cast<double *>(raw) + HEADER_DOUBLES;
}

void host_free(double *p) {
  if (!p)
    return;
  void *raw = p - HEADER_DOUBLES;
  if (*static_cast<uint64_t *>(raw) == TAG_PINNED)
    cals_b200_host_free(raw);
  else
    operator delete[](raw, std::align_val_t(64));
}

uint64_t next_tensor_uid() {
  static std::atomic<uint64_t> counter{0};
  return ++counter;
}

} // namespace detail

// ---------------------------------------------------------------------------------------------------------------------
// Tensor
static dim_t product(const vector<dim_t> &m) { return std::accumulate(m.begin(), m.end(), dim_t(1), std::multiplies<>()); }

void Tensor::allocate(dim_t n) {
  data_up.reset(detail::host_alloc(n));
  data = data_up.get();
}

Tensor::Tensor(const vector<dim_t> &shape) : n_elements{product(shape)}, max_n_elements{n_elements}, modes{shape} {
  allocate(n_elements);
}

Tensor::Tensor(const vector<dim_t> &shape, double *view_data)
    : n_elements{product(shape)}, max_n_elements{n_elements}, modes{shape}, data{view_data} {}

Tensor::Tensor(dim_t mode0, dim_t mode1, double *view_data)
    : n_elements{mode0 * mode1}, max_n_elements{n_elements}, modes{mode0, mode1} {
  if (view_data)
    data = view_data;
  else
    allocate(n_elements);
}

// Text format of the reference's data files (reference src/tensor.cpp:35-65): first line = extents separated by
// blanks, then the values in column-major order.  The whole file is read in one go and parsed with strtod.
Tensor::Tensor(const std::string &file_name) {
  std::ifstream in(file_name, std::ios::in | std::ios::binary);
  if (!in)
    throw std::runtime_error("cals::Tensor: cannot open " + file_name);
  std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  const size_t eol = text.find('\n');
  std::istringstream head(text.substr(0, eol));
  for (dim_t v; head >> v;)
    modes.push_back(v);
  if (modes.empty())
    throw std::runtime_error("cals::Tensor: no extents on the first line of " + file_name);
  n_elements = max_n_elements = product(modes);
  allocate(n_elements);
  const char *p = text.c_str() + (eol == std::string::npos ? text.size() : eol + 1);
  dim_t got = 0;
  while (got < n_elements) {
    char *end = nullptr;
    const double v = std::strtod(p, &end);
    if (end == p)
      break;
    data[got++] = v;
    p = end;
  }
  if (got != n_elements)
    throw std::runtime_error("cals::Tensor: " + file_name + " holds " + std::to_string(got) + " values, expected " +
                             std::to_string(n_elements));
}

Tensor::Tensor(dim_t rank_, const vector<dim_t> &shape) {
  Ktensor P(rank_, shape);
  P.randomize();
  *this = P.to_tensor();
  rank = static_cast<int>(rank_);
}

Tensor::Tensor(const Tensor &rhs) : rank{rhs.rank}, n_elements{rhs.n_elements}, max_n_elements{rhs.n_elements}, modes{rhs.modes} {
  if (rhs.is_view())
    data = rhs.data; // views copy as views (reference src/tensor.cpp:87-99)
  else {
    allocate(n_elements);
    std::copy(rhs.data, rhs.data + n_elements, data);
  }
}

Tensor &Tensor::operator=(const Tensor &rhs) {
  if (this == &rhs)
    return *this;
  rank = rhs.rank;
  n_elements = max_n_elements = rhs.n_elements;
  modes = rhs.modes;
  uid = detail::next_tensor_uid();
  if (rhs.is_view()) {
    data_up.reset();
    data = rhs.data;
  } else {
    allocate(n_elements);
    std::copy(rhs.data, rhs.data + n_elements, data);
  }
  return *this;
}

double Tensor::norm() const {
  // two-pass scaled sum of squares: same result class as dnrm2 (no overflow for huge entries)
  double amax = 0.0;
  for (dim_t i = 0; i < n_elements; i++)
    amax = std::max(amax, std::fabs(data[i]));
  if (amax == 0.0 || !std::isfinite(amax))
    return amax;
  const double inv = 1.0 / amax;
  double acc[4] = {0, 0, 0, 0};
  dim_t i = 0;
  for (; i + 4 <= n_elements; i += 4)
    for (int k = 0; k < 4; k++) {
      const double v = data[i + k] * inv;
      acc[k] += v * v;
    }
  for (; i < n_elements; i++) {
    const double v = data[i] * inv;
    acc[0] += v * v;
  }
  return amax * std::sqrt((acc[0] + acc[1]) + (acc[2] + acc[3]));
}

Tensor &Tensor::fill(const function<double()> &&f) {
  for (dim_t i = 0; i < n_elements; i++)
    data[i] = f();
  uid = detail::next_tensor_uid(); // new contents: a resident device copy of the old ones is stale
  return *this;
}

Tensor &Tensor::zero() {
  std::fill(data, data + n_elements, 0.0);
  uid = detail::next_tensor_uid();
  return *this;
}

Tensor &Tensor::randomize() {
  std::random_device seed;
  std::mt19937 gen(seed());
  std::uniform_real_distribution<double> dist(-1.0, 1.0);
  for (dim_t i = 0; i < n_elements; i++)
    data[i] = dist(gen);
  uid = detail::next_tensor_uid();
  return *this;
}

Unfolding Tensor::implicit_unfold(dim_t mode) const {
  dim_t below = 1, above = 1;
  for (dim_t k = 0; k < mode; k++)
    below *= modes[k];
  for (dim_t k = mode + 1; k < modes.size(); k++)
    above *= modes[k];
  Unfolding u{};
  u.rows = modes[mode];
  if (mode == 0) { // one I0 x rest block, column-major
    u.n_blocks = 1;
    u.block_offset = 0;
    u.cols = above;
    u.stride = modes[0];
  } else if (mode + 1 == modes.size()) { // one block, stored transposed
    u.n_blocks = 1;
    u.block_offset = 0;
    u.cols = below;
    u.stride = below;
  } else { // one transposed block per index combination of the modes above
    u.n_blocks = above;
    u.block_offset = below * modes[mode];
    u.cols = below;
    u.stride = below;
  }
  return u;
}

void Tensor::print(const std::string &&text) const {
  std::cout << "----------------------------------------\n" << text << "\nModes: ";
  for (dim_t m : modes)
    std::cout << m << " ";
  std::cout << "\ndata = [ ";
  for (dim_t i = 0; i < n_elements; i++)
    std::cout << std::setw(6) << data[i] << "  ";
  std::cout << "]\n----------------------------------------" << std::endl;
}

// ---------------------------------------------------------------------------------------------------------------------
// Matrix
void Matrix::print(const std::string &&text) const {
  std::cout << "----------------------------------------\n" << text << "\n----------------------------------------\n";
  std::cout << "Rows: " << rows << ", Cols: " << cols << "\n";
  std::cout.precision(4);
  for (dim_t r = 0; r < rows; r++) {
    for (dim_t c = 0; c < cols; c++)
      std::cout << "  " << std::setw(8) << (*this)(r, c) << "  ";
    std::cout << "\n";
  }
  std::cout << "----------------------------------------" << std::endl;
}

void Matrix::info() const {
  std::cout << "nRows: " << rows << ", nCols: " << cols << ", nElements: " << get_n_elements()
            << ", maxNElements: " << get_max_n_elements() << std::endl;
}

// ---------------------------------------------------------------------------------------------------------------------
// Ktensor
static double col_norm2(const double *x, dim_t n) {
  double amax = 0.0;
  for (dim_t i = 0; i < n; i++)
    amax = std::max(amax, std::fabs(x[i]));
  if (amax == 0.0)
    return 0.0;
  double s = 0.0;
  for (dim_t i = 0; i < n; i++) {
    const double v = x[i] / amax;
    s += v * v;
  }
  return amax * std::sqrt(s);
}

void Ktensor::init_active_sets() {
  active_set.assign(modes.size(), {});
  for (dim_t n = 0; n < modes.size(); n++)
    active_set[n].assign(modes[n], vector<bool>(components, true));
}

Ktensor::Ktensor(dim_t comps, const vector<dim_t> &shape)
    : id{detail::next_ktensor_id()}, components{comps}, modes(shape), lambda(comps, 0.0), factors(shape.size()) {
  assert(comps > 0);
  for (dim_t n = 0; n < shape.size(); n++)
    factors[n] = Matrix{shape[n], comps};
  init_active_sets();
}

Ktensor::Ktensor(const Ktensor &rhs)
    : id{detail::next_ktensor_id()}, components{rhs.components}, jk{rhs.jk}, modes(rhs.modes), lambda(rhs.lambda),
      factors(rhs.factors) {
  init_active_sets();
}

Ktensor &Ktensor::operator=(const Ktensor &rhs) {
  if (this == &rhs)
    return *this;
  id = detail::next_ktensor_id();
  components = rhs.components;
  jk = rhs.jk;
  modes = rhs.modes;
  lambda = rhs.lambda;
  factors = rhs.factors;
  return *this;
}

Ktensor &Ktensor::normalize() {
  std::fill(lambda.begin(), lambda.end(), 1.0);
  for (Matrix &f : factors)
    for (dim_t c = 0; c < f.get_cols(); c++) {
      double *col = f.get_data() + c * f.get_col_stride();
      const double nrm = col_norm2(col, f.get_rows());
      const double inv = 1 / nrm;
      for (dim_t r = 0; r < f.get_rows(); r++)
        col[r] *= inv;
      lambda[c] *= nrm;
    }
  normalized = true;
  return *this;
}

Ktensor &Ktensor::normalize(dim_t mode, dim_t iteration) {
  Matrix &f = factors[mode];
  for (dim_t c = 0; c < f.get_cols(); c++) {
    double *col = f.get_data() + c * f.get_col_stride();
    if (iteration == 1)
      lambda[c] = col_norm2(col, f.get_rows());
    else {
      dim_t arg = 0; // first index of the largest magnitude, as idamax
      for (dim_t r = 1; r < f.get_rows(); r++)
        if (std::fabs(col[r]) > std::fabs(col[arg]))
          arg = r;
      lambda[c] = col[arg];
    }
    if (lambda[c] != 0) {
      const double inv = 1 / lambda[c];
      for (dim_t r = 0; r < f.get_rows(); r++)
        col[r] *= inv;
    }
  }
  return *this;
}

Ktensor &Ktensor::denormalize() {
  Matrix &f = factors[0];
  for (dim_t c = 0; c < f.get_cols(); c++)
    for (dim_t r = 0; r < f.get_rows(); r++)
      f(r, c) *= lambda[c];
  normalized = false;
  return *this;
}

Ktensor &Ktensor::randomize() {
  for (Matrix &f : factors)
    f.randomize();
  set_jk_fiber(0.0);
  return normalize();
}

Ktensor &Ktensor::fill(function<double()> &&func) {
  for (Matrix &f : factors)
    for (dim_t i = 0; i < f.get_n_elements(); i++)
      f[i] = func();
  set_jk_fiber(0.0);
  return normalize();
}

// X[i0, i1, ..] = sum_r lambda_r prod_n F_n[i_n, r]; the running products of the slow modes are kept per level so the
// inner loop over mode 0 is one multiply-add per component.
Tensor Ktensor::to_tensor() {
  const dim_t N = factors.size(), R = get_components();
  vector<dim_t> shape(N);
  for (dim_t n = 0; n < N; n++)
    shape[n] = factors[n].get_rows();
  Tensor X(shape);
  double *out = X.get_data();
  const dim_t I0 = shape[0], outer = X.get_n_elements() / I0;
  vector<double> w(R);
  vector<dim_t> idx(N, 0);
  for (dim_t o = 0; o < outer; o++) {
    for (dim_t r = 0; r < R; r++) {
      double p = lambda[r];
      for (dim_t n = 1; n < N; n++)
        p *= factors[n](idx[n], r);
      w[r] = p;
    }
    double *dst = out + o * I0;
    for (dim_t i = 0; i < I0; i++) {
      double s = 0.0;
      for (dim_t r = 0; r < R; r++)
        s += w[r] * factors[0](i, r);
      dst[i] = s;
    }
    for (dim_t n = 1; n < N; n++) {
      if (++idx[n] < shape[n])
        break;
      idx[n] = 0;
    }
  }
  return X;
}

Ktensor &Ktensor::copy(Ktensor &rhs) {
  approx_error = rhs.approx_error;
  fit = rhs.fit;
  old_fit = rhs.old_fit;
  iters = rhs.iters;
  normalized = rhs.normalized;
  lambda = rhs.lambda;
  active_set = rhs.active_set;
  for (dim_t n = 0; n < factors.size(); n++)
    factors[n].copy(rhs.get_factor(n));
  return *this;
}

Ktensor &Ktensor::attach(vector<double *> &data_ptrs, bool /*multi_thread*/) {
  assert(data_ptrs.size() == factors.size());
  for (dim_t n = 0; n < factors.size(); n++) {
    Matrix &f = factors[n];
    std::copy(f.get_data(), f.get_data() + f.get_n_elements(), data_ptrs[n]);
    f.attach(data_ptrs[n]);
  }
  return *this;
}

Ktensor &Ktensor::detach() {
  for (Matrix &f : factors) {
    double *foreign = f.get_data();
    f.detach();
    if (foreign != f.get_data()) {
      std::copy(foreign, foreign + f.get_n_elements(), f.get_data());
      std::fill(foreign, foreign + f.get_n_elements(), 0.0);
    }
  }
  return *this;
}

Ktensor Ktensor::to_regular() {
  if (!jk.enabled)
    return *this;
  vector<dim_t> shape(modes);
  shape[jk.mode] -= 1;
  Ktensor out(components, shape);
  for (dim_t n = 0; n < modes.size(); n++) {
    const Matrix &src = factors[n];
    Matrix &dst = out.get_factor(n);
    for (dim_t c = 0; c < src.get_cols(); c++)
      for (dim_t r = 0, w = 0; r < src.get_rows(); r++) {
        if (n == jk.mode && r == jk.fiber)
          continue;
        dst(w++, c) = src(r, c);
      }
  }
  out.get_lambda() = lambda;
  return out;
}

void Ktensor::print(const std::string &&text) const {
  std::cout << "----------------------------------------\n" << text << "\n----------------------------------------\n";
  std::cout << "Rank: " << get_components() << "\nNum Modes: " << get_n_modes() << "\nModes: [ ";
  for (const Matrix &f : factors)
    std::cout << f.get_rows() << " ";
  std::cout << "]\nWeights: [";
  for (double l : lambda)
    std::cout << l << " ";
  std::cout << " ] " << std::endl;
  for (const Matrix &f : factors)
    f.print("factor");
  std::cout << "----------------------------------------" << std::endl;
}

} // namespace cals
