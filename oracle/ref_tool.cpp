// TEST INFRASTRUCTURE -- not part of the product path.
//
// cals_ref: a small command-line harness (ours) over the UNMODIFIED reference library
// (compiled from /root/reference by oracle/build_ref.sh into oracle/_ref/).  It reads a
// "case file" (tensor + initial models + parameters, written by oracle/caseio.py), runs the
// reference's own entry points and writes the fitted models back:
//
//   algo 0: cals::cp_cals      (reference src/cals.cpp:19)    -- the hot path's CPU truth
//   algo 1: cals::cp_als loop  (reference src/als.cpp:19)     -- the reference's own cross-check
//   algo 2: cals::jk_cp_cals   (reference src/cals.cpp:397)
//   algo 3: cals::jk_cp_als    (reference src/als.cpp:291)
//
// Used by (a) tests/ to pin oracle/cals_oracle.c and to generate tests/golden/*.npz
// (oracle/make_golden.py), (b) bench.py's cpu_baseline / --impl reference arm.
//
// Case file (little endian):  see oracle/caseio.py for the authoritative description.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "als.h"
#include "cals.h"

namespace {

struct ModelSpec {
  int64_t rank, jk_mode, jk_fiber;
};

template <typename T> void rd(FILE *f, T *p, size_t n) {
  if (fread(p, sizeof(T), n, f) != n) {
    fprintf(stderr, "cals_ref: short read\n");
    exit(2);
  }
}
template <typename T> void wr(FILE *f, const T *p, size_t n) {
  if (fwrite(p, sizeof(T), n, f) != n) {
    fprintf(stderr, "cals_ref: short write\n");
    exit(2);
  }
}

void dump_model(FILE *f, const cals::Ktensor &kt) {
  int64_t rank = (int64_t)kt.get_components(), iters = (int64_t)kt.get_iters();
  double err = kt.get_approximation_error(), fit_diff = kt.get_fit_diff();
  wr(f, &rank, 1);
  wr(f, &iters, 1);
  wr(f, &err, 1);
  wr(f, &fit_diff, 1);
  for (dim_t n = 0; n < kt.get_n_modes(); n++) {
    int64_t rows = (int64_t)kt.get_factor(n).get_rows();
    wr(f, &rows, 1);
  }
  wr(f, kt.get_lambda().data(), (size_t)rank);
  for (dim_t n = 0; n < kt.get_n_modes(); n++) {
    const auto &F = kt.get_factor(n);
    // factors are dense column-major with col_stride == rows once detached
    for (dim_t c = 0; c < F.get_cols(); c++)
      wr(f, F.get_data() + c * F.get_col_stride(), (size_t)F.get_rows());
  }
}

} // namespace

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: cals_ref <case.in> <case.out>\n");
    return 1;
  }
  FILE *fi = fopen(argv[1], "rb");
  if (!fi) {
    perror("open input");
    return 1;
  }
  char magic[8];
  rd(fi, magic, 8);
  if (memcmp(magic, "CALSIN01", 8) != 0) {
    fprintf(stderr, "bad magic\n");
    return 1;
  }
  int64_t n_modes;
  rd(fi, &n_modes, 1);
  std::vector<int64_t> modes64(n_modes);
  rd(fi, modes64.data(), (size_t)n_modes);
  int64_t n_models, max_iter, buffer_size, flags, threads, algo, mttkrp_method;
  double tol;
  rd(fi, &n_models, 1);
  rd(fi, &max_iter, 1);
  rd(fi, &tol, 1);
  rd(fi, &buffer_size, 1);
  rd(fi, &flags, 1);
  rd(fi, &threads, 1);
  rd(fi, &algo, 1);
  rd(fi, &mttkrp_method, 1);
  int64_t ls_method, ls_interval;
  double ls_step;
  rd(fi, &ls_method, 1);
  rd(fi, &ls_interval, 1);
  rd(fi, &ls_step, 1);
  std::vector<ModelSpec> specs(n_models);
  for (auto &s : specs) {
    rd(fi, &s.rank, 1);
    rd(fi, &s.jk_mode, 1);
    rd(fi, &s.jk_fiber, 1);
  }

  std::vector<dim_t> modes(modes64.begin(), modes64.end());
  set_threads((int)threads);

  cals::Tensor X(modes);
  rd(fi, X.get_data(), X.get_n_elements());

  std::vector<cals::Ktensor> models;
  models.reserve((size_t)n_models);
  for (auto &s : specs) {
    models.emplace_back((dim_t)s.rank, modes);
    auto &kt = models.back();
    std::vector<double> buf;
    for (int64_t n = 0; n < n_modes; n++) {
      buf.resize((size_t)(modes[n] * s.rank));
      rd(fi, buf.data(), buf.size());
      kt.set_factor((int)n, buf.data());
    }
    buf.resize((size_t)s.rank);
    rd(fi, buf.data(), buf.size());
    kt.set_lambda(buf.data());
    if (s.jk_mode >= 0)
      kt.to_jk((dim_t)s.jk_mode, (dim_t)s.jk_fiber);
  }
  fclose(fi);

  const bool force_max_iter = flags & 1, always_evict_first = flags & 2, nnls = flags & 4, line_search = flags & 8;

  double seconds = 0.0;
  int64_t rep_iter = 0, rep_n_ktensors = 0, rep_comp_sum = 0;
  double x_norm = X.norm();
  std::vector<std::vector<cals::Ktensor>> jk_results;

  auto t0 = std::chrono::steady_clock::now();
  if (algo == 0) {
    cals::CalsParams p;
    p.mttkrp_method = (cals::mttkrp::MTTKRP_METHOD)mttkrp_method;
    p.max_iterations = (dim_t)max_iter;
    p.tol = tol;
    p.buffer_size = (dim_t)buffer_size;
    p.force_max_iter = force_max_iter;
    p.always_evict_first = always_evict_first;
    if (nnls)
      p.update_method = cals::update::UPDATE_METHOD::NNLS;
    if (line_search) {
      p.line_search = true;
      p.line_search_method = (cals::ls::LS_METHOD)ls_method;
      p.line_search_interval = (int)ls_interval;
      p.line_search_step = ls_step;
    }
    cals::KtensorQueue q;
    for (auto &m : models)
      q.emplace(m);
    auto rep = cals::cp_cals(X, q, p);
    seconds = rep.total_time;
    rep_iter = (int64_t)rep.iter;
    rep_n_ktensors = rep.n_ktensors;
    rep_comp_sum = rep.ktensor_comp_sum;
  } else if (algo == 1) {
    cals::AlsParams p;
    p.mttkrp_method = (cals::mttkrp::MTTKRP_METHOD)mttkrp_method;
    p.max_iterations = (dim_t)max_iter;
    p.tol = tol;
    p.force_max_iter = force_max_iter;
    p.suppress_lut_warning = true;
    if (nnls)
      p.update_method = cals::update::UPDATE_METHOD::NNLS;
    if (line_search) {
      p.line_search = true;
      p.line_search_method = (cals::ls::LS_METHOD)ls_method;
      p.line_search_interval = (int)ls_interval;
      p.line_search_step = ls_step;
    }
    for (auto &m : models) {
      auto rep = cals::cp_als(X, m, p);
      seconds += rep.total_time;
      rep_iter += (int64_t)rep.iter;
    }
    rep_n_ktensors = n_models;
  } else if (algo == 2) {
    cals::CalsParams p;
    p.mttkrp_method = (cals::mttkrp::MTTKRP_METHOD)mttkrp_method;
    p.max_iterations = (dim_t)max_iter;
    p.tol = tol;
    p.buffer_size = (dim_t)buffer_size;
    p.force_max_iter = force_max_iter;
    auto rep = cals::jk_cp_cals(X, models, p);
    seconds = rep.jk_time.als_time;
    jk_results = std::move(rep.results);
  } else if (algo == 3) {
    cals::AlsParams p;
    p.mttkrp_method = (cals::mttkrp::MTTKRP_METHOD)mttkrp_method;
    p.max_iterations = (dim_t)max_iter;
    p.tol = tol;
    p.force_max_iter = force_max_iter;
    p.suppress_lut_warning = true;
    auto rep = cals::jk_cp_als(X, models, p);
    seconds = rep.jk_time.als_time;
    jk_results = std::move(rep.results);
  } else {
    fprintf(stderr, "unknown algo %ld\n", (long)algo);
    return 1;
  }
  double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  FILE *fo = fopen(argv[2], "wb");
  if (!fo) {
    perror("open output");
    return 1;
  }
  wr(fo, "CALSOUT1", 8);
  int64_t n_out = 0;
  if (algo == 2 || algo == 3)
    for (auto &v : jk_results)
      n_out += (int64_t)v.size();
  else
    n_out = n_models;
  wr(fo, &n_out, 1);
  wr(fo, &seconds, 1);
  wr(fo, &wall, 1);
  wr(fo, &rep_iter, 1);
  wr(fo, &rep_n_ktensors, 1);
  wr(fo, &rep_comp_sum, 1);
  wr(fo, &x_norm, 1);
  if (algo == 2 || algo == 3) {
    for (auto &v : jk_results)
      for (auto &m : v)
        dump_model(fo, m);
  } else
    for (auto &m : models)
      dump_model(fo, m);
  fclose(fo);
  printf("{\"algo\": %ld, \"seconds\": %.6f, \"wall\": %.6f, \"iter\": %ld, \"threads\": %d}\n", (long)algo, seconds, wall,
         (long)rep_iter, get_threads());
  return 0;
}
