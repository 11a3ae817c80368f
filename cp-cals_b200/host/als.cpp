// Single-model entry points (reference src/als.cpp:19-419) on the B200 engine.
//
// cp_als is the C = R special case of the concurrent path: one model, buffer = its rank, the same kernels in the same
// order, so cp_als and cp_cals agree on every model up to the summation order of the MTTKRP tiles (the reference pins
// the same equivalence at 1e-11, tests/cals/test_cals.cpp:13-86).
#include <algorithm>
#include <cmath>
#include <fstream>
#include <iostream>
#include <limits>

#include "als.h"
#include "cals.h"
#include "engine_pool.h"

namespace cals {

static AlsReport run_single(const Tensor &X, Ktensor &ktensor, AlsParams &p, bool may_skip_upload) {
  if (p.update_method != update::UNCONSTRAINED && p.update_method != update::NNLS)
    throw B200Error("cp_als: unknown update method");
  if (p.line_search && p.line_search_method != ls::NO_ERROR_CHECKING && p.line_search_method != ls::ERROR_CHECKING_SERIAL)
    throw B200Error("cp_als: this line search method is not on the B200 path; there is no CPU fallback");
  if (X.get_n_modes() < 3)
    throw B200Error("cp_als: tensors need at least 3 modes (the reference asserts the same, src/als.cpp:51)");

  AlsReport rep;
  rep.tensor_rank = X.get_rank();
  rep.n_modes = X.get_n_modes();
  rep.modes = X.get_modes();
  rep.max_iter = p.max_iterations;
  rep.n_threads = get_threads();
  rep.ktensor_id = ktensor.get_id();
  rep.ktensor_components = ktensor.get_components();
  rep.tol = p.tol;
  rep.cuda = true;
  rep.update_method = p.update_method;
  rep.line_search = p.line_search;
  rep.line_search_interval = p.line_search_interval;
  rep.line_search_step = p.line_search_step;
  rep.line_search_method = p.line_search_method;

  Timer total;
  total.start();
  detail::RunOptions opt;
  opt.buffer_size = ktensor.get_components();
  opt.max_iterations = p.max_iterations;
  opt.tol = p.tol;
  opt.force_max_iter = p.force_max_iter;
  opt.nnls = p.update_method == update::NNLS;
  opt.line_search = p.line_search;
  opt.ls_method = static_cast<int>(p.line_search_method);
  opt.ls_interval = p.line_search_interval;
  opt.ls_step = p.line_search_step;
  opt.pair_node = p.mttkrp_method != mttkrp::MTTKRP_METHOD::MTTKRP;
  opt.skip_upload_if_resident = may_skip_upload;
  std::vector<Ktensor *> one{&ktensor};
  const detail::RunResult r = detail::run_on_device(p.device, X, one, opt);
  total.stop();

  rep.X_norm = r.rep.x_norm;
  rep.iter = r.rep.iter;
  rep.ls_performed = r.ls_performed;
  rep.ls_failed = r.ls_failed;
  rep.total_time = total.get_time();
  rep.flops_per_iteration = 2ull * rep.n_modes * X.get_n_elements() * ktensor.get_components();
  const dim_t its = std::max<dim_t>(rep.iter, 1);
  rep.als_times = Matrix(AlsTimers::LENGTH, std::max<dim_t>(rep.max_iter, its));
  rep.mode_times = Matrix(ModeTimers::LENGTH * rep.n_modes, std::max<dim_t>(rep.max_iter, its));
  rep.mttkrp_times = Matrix(MttkrpTimers::LENGTH * rep.n_modes, std::max<dim_t>(rep.max_iter, its));
  rep.als_times.zero();
  rep.mode_times.zero();
  rep.mttkrp_times.zero();
  for (dim_t i = 0; i < rep.iter; i++)
    rep.als_times(AlsTimers::ITERATION, i) = r.rep.device_ms * 1e-3 / static_cast<double>(its);
  return rep;
}

AlsReport cp_als(const Tensor &X, Ktensor &ktensor, AlsParams &params) {
  return run_single(X, ktensor, params, params.cuda_no_tensor_alloc);
}

// The reference runs this as an OpenMP loop over models sharing one device copy of X (src/als.cpp:390-419); here the
// tensor is uploaded by the first call and the rest find it resident.
vector<AlsReport> cp_omp_als(const Tensor &X, vector<Ktensor> &ktensor_v, AlsParams &params) {
  Timer total;
  total.start();
  vector<AlsReport> reports(ktensor_v.size());
  for (size_t i = 0; i < ktensor_v.size(); i++)
    reports[i] = run_single(X, ktensor_v[i], params, i > 0 || params.cuda_no_tensor_alloc);
  total.stop();
  for (AlsReport &r : reports)
    r.total_time = total.get_time();
  return reports;
}

// ---------------------------------------------------------------------------------------------------------------------
// Leave-one-out fits on explicitly row-deleted tensors: the baseline the reference compares jk_cp_cals with
// (src/als.cpp:291-388).  Works for any number of modes (the reference views X as I0 x (I1*I2)).
static Tensor delete_mode0_row(const Tensor &X, dim_t row) {
  vector<dim_t> modes = X.get_modes();
  const dim_t I0 = modes[0], rest = X.get_n_elements() / I0;
  modes[0] -= 1;
  Tensor out(modes);
  const double *src = X.get_data();
  double *dst = out.get_data();
  for (dim_t j = 0; j < rest; j++) {
    std::copy(src + j * I0, src + j * I0 + row, dst + j * (I0 - 1));
    std::copy(src + j * I0 + row + 1, src + (j + 1) * I0, dst + j * (I0 - 1) + row);
  }
  return out;
}

static Ktensor delete_mode0_row(const Ktensor &kt, dim_t row) {
  vector<dim_t> modes = kt.get_modes();
  modes[0] -= 1;
  Ktensor out(kt.get_components(), modes);
  out.get_lambda() = kt.get_lambda();
  for (dim_t n = 0; n < kt.get_n_modes(); n++) {
    const Matrix &src = kt.get_factor(n);
    Matrix &dst = out.get_factor(n);
    if (n != 0) {
      dst.copy(src);
      continue;
    }
    for (dim_t c = 0; c < src.get_cols(); c++)
      for (dim_t r = 0, w = 0; r < src.get_rows(); r++)
        if (r != row)
          dst(w++, c) = src(r, c);
  }
  return out;
}

static JKReport jk_als_impl(const Tensor &X, vector<Ktensor> &kt_vector, AlsParams &als_params) {
  vector<Ktensor> bases(kt_vector);
  for (Ktensor &b : bases)
    b.denormalize().normalize();
  const dim_t samples = X.get_modes()[0];

  double pre_time = 0.0, als_time = 0.0;
  vector<vector<Ktensor>> jk_input(bases.size());
  for (size_t b = 0; b < bases.size(); b++) {
    jk_input[b].resize(samples);
    for (dim_t i = 0; i < samples; i++) {
      Timer pre, run;
      pre.start();
      jk_input[b][i] = delete_mode0_row(bases[b], i);
      Tensor X_jk = delete_mode0_row(X, i);
      pre.stop();
      run.start();
      cp_als(X_jk, jk_input[b][i], als_params);
      run.stop();
      pre_time += pre.get_time();
      als_time += run.get_time();
    }
  }
  for (auto &group : jk_input)
    for (Ktensor &m : group)
      m.denormalize().normalize();
  for (size_t b = 0; b < bases.size(); b++)
    utils::jk_permutation_adjustment(bases[b], jk_input[b]);
  return JKReport{JKTime{pre_time, als_time}, std::move(jk_input)};
}

JKReport jk_cp_als(const Tensor &X, vector<Ktensor> &kt_vector, AlsParams &als_params) {
  return jk_als_impl(X, kt_vector, als_params);
}

// One GPU runs one model at a time either way: the "omp" flavour only differs on the reference's CPU.
JKReport jk_cp_omp_als(const Tensor &X, vector<Ktensor> &kt_vector, AlsParams &als_params) {
  return jk_als_impl(X, kt_vector, als_params);
}

// ---------------------------------------------------------------------------------------------------------------------
void AlsParams::print() const {
  using std::cout;
  using std::endl;
  const char *rule = "---------------------------------------";
  cout << rule << "\nALS parameters\n" << rule << endl;
  cout << "Tol:             " << tol << endl;
  cout << "Max Iterations:  " << max_iterations << endl;
  cout << "Mttkrp Method:   " << mttkrp::mttkrp_method_names[mttkrp_method] << " (MTTKRP: one contraction per mode; otherwise shared where possible)" << endl;
  cout << "Update Method:   " << update::update_method_names[update_method] << endl;
  cout << "Line Search:     " << (line_search ? "true" : "false") << endl;
  if (line_search) {
    cout << "-Line Search Interval: " << line_search_interval << " iterations" << endl;
    cout << "-Line Search Method:   " << ls::ls_method_names[line_search_method] << endl;
  }
  cout << "CUDA:            true (B200, device " << device << ")" << endl << rule << endl;
}

// CSV layout of the reference (include/als.h:70-135): one line per cp_als call, timers as minima over iterations.
void AlsReport::print_header(const std::string &file_name, const std::string &sep) const {
  std::ofstream file(file_name, std::ios::out);
  for (const char *name : {"TENSOR_RANK", "TENSOR_MODES", "KTENSOR_ID", "KTENSOR_COMP", "UPDATE_METHOD", "LINE_SEARCH",
                           "MAX_ITERS", "ITER", "NUM_THREADS", "TOTAL", "FLOPS"})
    file << name << sep;
  AlsTimers at;
  ModeTimers mt;
  for (const auto &name : at.names)
    file << name << sep;
  for (dim_t n = 0; n < modes.size(); n++)
    for (const auto &name : mt.names)
      file << "MODE_" << n << "_" << name << sep;
  file << std::endl;
}

void AlsReport::print_to_file(const std::string &file_name, const std::string &sep) const {
  std::ofstream file(file_name, std::ios::app);
  file << tensor_rank << sep << utils::mode_string(modes) << sep << ktensor_id << sep << ktensor_components << sep
       << update::update_method_names[update_method] << sep << line_search << sep << max_iter << sep << iter << sep
       << n_threads << sep << total_time << sep << flops_per_iteration << sep;
  file << std::scientific;
  auto row_min = [&](const Matrix &t, dim_t row) {
    double best = std::numeric_limits<double>::max();
    for (dim_t j = 0; j + 1 < iter; j++) // the reference leaves the last iteration out (include/als.h:112-128)
      best = std::min(best, t(row, j));
    return best;
  };
  for (dim_t i = 0; i < als_times.get_rows(); i++)
    file << row_min(als_times, i) << sep;
  for (dim_t i = 0; i < mode_times.get_rows(); i++)
    file << row_min(mode_times, i) << sep;
  file << std::endl;
}

} // namespace cals
