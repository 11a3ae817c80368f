// bench_e2e -- the end-to-end number of bench.py measured through the drop-in C++ API: every step is one
// cals::cp_cals(const Tensor&, KtensorQueue&, CalsParams&) call (reference include/cals.h:196) over HOST containers --
// upload of X and of the initial models, the concurrent-ALS loop on the GPU, and the fitted models copied back into
// the caller's Ktensor objects -- timed with the host clock around the call, exactly what a user of the reference who
// relinks against libcals.so sees.
//
//   bench_e2e <case.in> <steps> <warmup> [device]
//
// <case.in> is the case-file format that oracle/caseio.py writes for the reference harness (tensor, initial models,
// parameters), so bench.py hands the very same inputs to this program, to the Python wrapper and to the CPU reference.
// Prints one JSON line: seconds per step (mean, min), device-loop ms, checksum of the fits.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cals.h"

namespace {
template <typename T> void rd(FILE *f, T *p, size_t n) {
  if (fread(p, sizeof(T), n, f) != n) {
    fprintf(stderr, "bench_e2e: short read\n");
    exit(2);
  }
}
} // namespace

int main(int argc, char **argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: bench_e2e <case.in> <steps> <warmup> [device]\n");
    return 1;
  }
  const int steps = atoi(argv[2]), warmup = atoi(argv[3]), device = argc > 4 ? atoi(argv[4]) : 0;
  FILE *fi = fopen(argv[1], "rb");
  if (!fi) {
    perror("bench_e2e: open case file");
    return 1;
  }
  char magic[8];
  rd(fi, magic, 8);
  if (memcmp(magic, "CALSIN01", 8) != 0) {
    fprintf(stderr, "bench_e2e: bad magic\n");
    return 1;
  }
  int64_t n_modes, n_models, max_iter, buffer_size, flags, threads, algo, method, ls_method, ls_interval;
  double tol, ls_step;
  rd(fi, &n_modes, 1);
  std::vector<int64_t> modes64((size_t)n_modes);
  rd(fi, modes64.data(), (size_t)n_modes);
  rd(fi, &n_models, 1);
  rd(fi, &max_iter, 1);
  rd(fi, &tol, 1);
  rd(fi, &buffer_size, 1);
  rd(fi, &flags, 1);
  rd(fi, &threads, 1);
  rd(fi, &algo, 1);
  rd(fi, &method, 1);
  rd(fi, &ls_method, 1);
  rd(fi, &ls_interval, 1);
  rd(fi, &ls_step, 1);
  struct Spec {
    int64_t rank, jk_mode, jk_fiber;
  };
  std::vector<Spec> specs((size_t)n_models);
  for (auto &s : specs) {
    rd(fi, &s.rank, 1);
    rd(fi, &s.jk_mode, 1);
    rd(fi, &s.jk_fiber, 1);
  }
  std::vector<dim_t> modes(modes64.begin(), modes64.end());
  cals::Tensor X(modes); // page-locked when large (cals::detail::host_alloc)
  rd(fi, X.get_data(), (size_t)X.get_n_elements());
  std::vector<cals::Ktensor> initial;
  initial.reserve((size_t)n_models);
  size_t model_bytes = 0, lambda_bytes = 0;
  for (auto &s : specs) {
    initial.emplace_back((dim_t)s.rank, modes);
    cals::Ktensor &kt = initial.back();
    std::vector<double> buf;
    for (int64_t n = 0; n < n_modes; n++) {
      buf.resize((size_t)(modes[(size_t)n] * (dim_t)s.rank));
      rd(fi, buf.data(), buf.size());
      kt.set_factor((int)n, buf.data());
      model_bytes += buf.size() * 8;
    }
    buf.resize((size_t)s.rank);
    rd(fi, buf.data(), buf.size());
    kt.set_lambda(buf.data());
    lambda_bytes += (size_t)s.rank * 8;
    if (s.jk_mode >= 0)
      kt.to_jk((dim_t)s.jk_mode, (dim_t)s.jk_fiber);
  }
  fclose(fi);

  cals::CalsParams p;
  p.mttkrp_method = (cals::mttkrp::MTTKRP_METHOD)method;
  p.max_iterations = (dim_t)max_iter;
  p.tol = tol;
  p.buffer_size = (dim_t)buffer_size;
  p.force_max_iter = flags & 1;
  p.always_evict_first = flags & 2;
  if (flags & 4)
    p.update_method = cals::update::UPDATE_METHOD::NNLS;
  p.devices = {device};

  std::vector<cals::Ktensor> work(initial); // the caller's models: overwritten in place by every call
  double sum = 0.0, best = 1e300, device_ms = 0.0, fit_sum = 0.0;
  uint64_t iters = 0;
  try {
    for (int s = 0; s < warmup + steps; s++) {
      for (size_t m = 0; m < work.size(); m++) // fresh starting values (host-side copy, outside the timed call)
        work[m].copy(initial[m]);
      cals::KtensorQueue q;
      for (auto &kt : work)
        q.emplace(kt);
      const auto t0 = std::chrono::steady_clock::now();
      const cals::CalsReport rep = cals::cp_cals(X, q, p);
      const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (s >= warmup) {
        sum += sec;
        best = sec < best ? sec : best;
        device_ms += rep.device_ms;
      }
      iters = rep.iter;
    }
  } catch (const std::exception &e) {
    fprintf(stderr, "bench_e2e: %s\n", e.what());
    return 3;
  }
  uint64_t model_iters = 0;
  for (auto &kt : work) {
    fit_sum += kt.get_fit();
    model_iters += kt.get_iters();
  }
  printf("{\"api\": \"cals::cp_cals (libcals.so)\", \"steps\": %d, \"warmup\": %d, \"s_per_step\": %.9f, "
         "\"s_per_step_min\": %.9f, \"device_ms_per_step\": %.6f, \"global_iters\": %llu, \"model_iters_per_step\": %llu, "
         "\"mean_fit\": %.15g, \"h2d_bytes_per_step\": %zu, \"d2h_bytes_per_step\": %zu}\n",
         steps, warmup, sum / steps, best, device_ms / steps, (unsigned long long)iters,
         (unsigned long long)model_iters, fit_sum / (double)work.size(), (size_t)X.get_n_elements() * 8 + model_bytes,
         model_bytes + lambda_bytes);
  return 0;
}
