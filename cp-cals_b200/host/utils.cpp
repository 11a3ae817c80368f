// Host helpers around the hot path: Gramian / Hadamard / error utilities for callers and tests, jackknife model
// generation and the component matching that follows jk_cp_cals (reference src/utils/utils.cpp, src/utils/error.cpp).
#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>

#include "utils/error.h"
#include "utils/utils.h"

namespace cals::utils {

std::string mode_string(std::vector<dim_t> const &modes) {
  std::string s;
  for (dim_t i = 0; i < modes.size(); i++)
    s += (i ? "-" : "") + std::to_string(modes[i]);
  return s;
}

Ktensor concatenate_ktensors(std::vector<Ktensor> const &ktensors) {
  const dim_t R = ktensors.at(0).get_components();
  Ktensor out(ktensors.size() * R, ktensors[0].get_modes());
  for (dim_t k = 0; k < ktensors.size(); k++) {
    const Ktensor &kt = ktensors[k];
    for (dim_t r = 0; r < R; r++)
      out.get_lambda()[k * R + r] = kt.get_lambda()[r];
    for (dim_t n = 0; n < kt.get_n_modes(); n++) {
      const Matrix &src = kt.get_factor(n);
      Matrix &dst = out.get_factor(n);
      for (dim_t r = 0; r < R; r++)
        for (dim_t i = 0; i < src.get_rows(); i++)
          dst(i, k * R + r) = src(i, r);
    }
  }
  return out;
}

void generate_jk_ktensors(Ktensor const &reference_ktensor, std::vector<Ktensor> &jk_ktensor_v) {
  const dim_t samples = reference_ktensor.get_modes()[0];
  if (samples <= 1)
    throw std::string("Can't do Jack-knife with just one sample."); // the reference throws a std::string here
  jk_ktensor_v.reserve(jk_ktensor_v.size() + samples);
  for (dim_t i = 0; i < samples; i++) {
    jk_ktensor_v.emplace_back(reference_ktensor);
    jk_ktensor_v.back().to_jk(0, i);
  }
}

// Shortest-augmenting-path assignment with dual potentials (Hungarian method, O(n^3)); square dense costs.
void linear_sum_assignment(dim_t n, const double *cost, bool maximize, std::vector<int64_t> &col_of_row) {
  const double INF = std::numeric_limits<double>::infinity();
  const double sign = maximize ? -1.0 : 1.0;
  // 1-based arrays; row_of_col[0] is the row being inserted
  std::vector<double> u(n + 1, 0.0), v(n + 1, 0.0), slack(n + 1);
  std::vector<dim_t> row_of_col(n + 1, 0), prev(n + 1, 0);
  std::vector<char> used(n + 1);
  for (dim_t i = 1; i <= n; i++) {
    row_of_col[0] = i;
    dim_t j0 = 0;
    std::fill(slack.begin(), slack.end(), INF);
    std::fill(used.begin(), used.end(), 0);
    do {
      used[j0] = 1;
      const dim_t i0 = row_of_col[j0];
      double delta = INF;
      dim_t j1 = 0;
      for (dim_t j = 1; j <= n; j++) {
        if (used[j])
          continue;
        const double reduced = sign * cost[(i0 - 1) * n + (j - 1)] - u[i0] - v[j];
        if (reduced < slack[j]) {
          slack[j] = reduced;
          prev[j] = j0;
        }
        if (slack[j] < delta) {
          delta = slack[j];
          j1 = j;
        }
      }
      for (dim_t j = 0; j <= n; j++) {
        if (used[j]) {
          u[row_of_col[j]] += delta;
          v[j] -= delta;
        } else
          slack[j] -= delta;
      }
      j0 = j1;
    } while (row_of_col[j0] != 0);
    do { // flip the augmenting path
      const dim_t j1 = prev[j0];
      row_of_col[j0] = row_of_col[j1];
      j0 = j1;
    } while (j0 != 0);
  }
  col_of_row.assign(n, 0);
  for (dim_t j = 1; j <= n; j++)
    col_of_row[row_of_col[j] - 1] = static_cast<int64_t>(j - 1);
}

// similarity(a, b) = <B[:, a], B_m[:, b]> + <C[:, a], C_m[:, b]> between the components of the full model and of a
// leave-one-out model.  The reference hands the column-major R x R matrix M = B^T B_m + C^T C_m to a row-major LSAP
// solver (reference src/utils/utils.cpp:74-86), i.e. it solves the assignment for M^T: row c of the cost is
// leave-one-out component c, column a is full-model component a.  With p = that assignment, every factor of the
// leave-one-out model gets new[:, c] = old[:, p[c]] (:88-99).  lambda is not permuted (as in the reference).
void jk_permutation_adjustment(Ktensor &ktensor, std::vector<Ktensor> &jk_ktensor_v) {
  const dim_t R = ktensor.get_components();
  const Matrix &B = ktensor.get_factor(1), &C = ktensor.get_factor(2);
  std::vector<double> cost(R * R);
  std::vector<int64_t> perm;
  for (Ktensor &m : jk_ktensor_v) {
    const Matrix &Bm = m.get_factor(1), &Cm = m.get_factor(2);
    for (dim_t c = 0; c < R; c++)
      for (dim_t a = 0; a < R; a++) {
        double s = 0.0;
        for (dim_t i = 0; i < B.get_rows(); i++)
          s += B(i, a) * Bm(i, c);
        double t = 0.0;
        for (dim_t i = 0; i < C.get_rows(); i++)
          t += C(i, a) * Cm(i, c);
        cost[c * R + a] = s + t;
      }
    linear_sum_assignment(R, cost.data(), true, perm);
    for (dim_t n = 0; n < m.get_n_modes(); n++) {
      Matrix &f = m.get_factor(n);
      Matrix old(f.get_rows(), f.get_cols());
      old.copy(f);
      for (dim_t c = 0; c < R; c++)
        if (static_cast<dim_t>(perm[c]) != c)
          for (dim_t i = 0; i < f.get_rows(); i++)
            f(i, c) = old(i, static_cast<dim_t>(perm[c]));
    }
  }
}

} // namespace cals::utils

namespace cals::ops {

void update_gramian(const Matrix &factor, Matrix &gramian) {
  const dim_t R = factor.get_cols(), I = factor.get_rows();
  for (dim_t j = 0; j < R; j++)
    for (dim_t i = 0; i < R; i++) {
      double s = 0.0;
      for (dim_t r = 0; r < I; r++)
        s += factor(r, i) * factor(r, j);
      gramian(i, j) = s;
    }
}

void update_gramians(const Ktensor &ktensor, std::vector<Matrix> &gramians) {
  for (dim_t n = 0; n < ktensor.get_n_modes(); n++)
    update_gramian(ktensor.get_factor(n), gramians[n]);
}

Matrix &hadamard_but_one(std::vector<Matrix> &matrices, dim_t mode) {
  Matrix &H = matrices[mode];
  for (dim_t e = 0; e < H.get_n_elements(); e++)
    H[e] = 1.0;
  for (dim_t k = 0; k < matrices.size(); k++)
    if (k != mode)
      H.hadamard(matrices[k]);
  return H;
}

void hadamard_all(std::vector<Matrix> &matrices) {
  for (dim_t k = 1; k < matrices.size(); k++)
    matrices[0].hadamard(matrices[k]);
}

} // namespace cals::ops

namespace cals::error {

double compute_fast_error(double X_norm, const std::vector<double> &lambda, const Matrix &last_factor,
                          const Matrix &last_mttkrp, const Matrix &gramian_hadamard) {
  double term2 = 0.0, term3 = 0.0;
  for (dim_t j = 0; j < gramian_hadamard.get_cols(); j++)
    for (dim_t i = 0; i < gramian_hadamard.get_rows(); i++)
      term2 += lambda[i] * lambda[j] * gramian_hadamard(i, j);
  for (dim_t j = 0; j < last_factor.get_cols(); j++)
    for (dim_t i = 0; i < last_factor.get_rows(); i++)
      term3 += lambda[j] * last_factor(i, j) * last_mttkrp(i, j);
  return std::sqrt(std::fmax(X_norm * X_norm + term2 - 2 * term3, 0));
}

double compute_error(const Tensor &X, Ktensor &ktensor, Matrix &, Matrix &) {
  Tensor approx = ktensor.to_tensor();
  if (approx.get_n_elements() != X.get_n_elements())
    throw std::invalid_argument("compute_error: model and tensor have different sizes");
  for (dim_t i = 0; i < X.get_n_elements(); i++)
    approx[i] = X[i] - approx[i];
  return approx.norm();
}

} // namespace cals::error
