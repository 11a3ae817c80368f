// driver -- the counterpart of the reference's src/examples/driver.cpp on the B200 path: fit MIN..MAX-component
// models (COPIES random starts each) to one random dense tensor, once concurrently with cals::cp_cals and once model
// by model with cals::cp_als, and print both times.  Same options as the reference driver (-n, -c, -t) plus:
//   -g --gpus d0,d1,..      shard the model set over these CUDA devices (default 0)
//   -i --iterations K       force exactly K ALS iterations per model (default: tol 1e-5, at most 1000 as the reference)
//   --no-als                skip the cp_als loop
#include <cstdlib>
#include <iostream>
#include <numeric>
#include <sstream>
#include <string>

#include "als.h"
#include "cals.h"

using std::cerr;
using std::cout;
using std::endl;

static std::vector<dim_t> split_numbers(const std::string &s, char sep) {
  std::vector<dim_t> out;
  std::stringstream ss(s);
  for (std::string tok; std::getline(ss, tok, sep);)
    if (!tok.empty())
      out.push_back(std::strtoul(tok.c_str(), nullptr, 10));
  return out;
}

static void usage(const char *exe) {
  cout << "USAGE: " << exe << " [options]\n"
       << "  -n --nthreads THREADS       recorded in the reports only (no host BLAS on this path)\n"
       << "  -c --components MIN:MAX:COPIES   ranks of the models and random starts per rank (default 1:10:10)\n"
       << "  -t --tensor DIM0-DIM1-DIM2[-..]  extents of the random target tensor (default 210-210-210)\n"
       << "  -g --gpus d0,d1,..          CUDA devices to shard the models over (default 0)\n"
       << "  -i --iterations K           force K iterations per model\n"
       << "     --no-als                 do not run the model-by-model cp_als loop\n";
}

int main(int argc, char **argv) {
  std::vector<dim_t> modes{210, 210, 210};
  dim_t rmin = 1, rmax = 10, copies = 10, forced = 0;
  int threads = 10;
  bool run_als = true;
  std::vector<int> devices;

  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto value = [&](const char *what) -> std::string {
      if (i + 1 >= argc) {
        cerr << what << " needs an argument" << endl;
        std::exit(1);
      }
      return argv[++i];
    };
    if (a == "-h" || a == "--help") {
      usage(argv[0]);
      return 0;
    } else if (a == "-n" || a == "--nthreads")
      threads = std::atoi(value("--nthreads").c_str());
    else if (a == "-c" || a == "--components") {
      const auto v = split_numbers(value("--components"), ':');
      if (v.size() != 3 || v[0] < 1 || v[1] < v[0] || v[2] < 1) {
        cerr << "--components expects MIN:MAX:COPIES" << endl;
        return 1;
      }
      rmin = v[0], rmax = v[1], copies = v[2];
    } else if (a == "-t" || a == "--tensor") {
      modes = split_numbers(value("--tensor"), '-');
      if (modes.size() < 3) {
        cerr << "--tensor expects at least three extents DIM0-DIM1-DIM2" << endl;
        return 1;
      }
    } else if (a == "-g" || a == "--gpus") {
      for (dim_t d : split_numbers(value("--gpus"), ','))
        devices.push_back(static_cast<int>(d));
    } else if (a == "-i" || a == "--iterations")
      forced = std::strtoul(value("--iterations").c_str(), nullptr, 10);
    else if (a == "--no-als")
      run_als = false;
    else {
      cerr << "unrecognised argument " << a << endl;
      usage(argv[0]);
      return 1;
    }
  }
  set_threads(threads);

  cals::Tensor X(modes);
  X.randomize();

  std::vector<dim_t> ranks;
  for (dim_t r = rmin; r <= rmax; r++)
    ranks.insert(ranks.end(), copies, r);
  std::vector<cals::Ktensor> cals_input;
  cals_input.reserve(ranks.size());
  for (dim_t r : ranks) {
    cals_input.emplace_back(r, modes);
    cals_input.back().randomize();
  }
  std::vector<cals::Ktensor> als_input(cals_input);
  cout << "Tensor " << cals::utils::mode_string(modes) << ", " << ranks.size() << " models, ranks " << rmin << ".."
       << rmax << " x" << copies << endl;

  cals::CalsParams cp;
  cp.max_iterations = forced ? forced : 1000;
  cp.tol = 1e-5;
  cp.force_max_iter = forced != 0;
  cp.buffer_size = std::accumulate(ranks.begin(), ranks.end(), dim_t(0));
  cp.devices = devices;
  cp.print();

  // Per-process costs that have nothing to do with the algorithm -- creating the CUDA context, loading the kernels,
  // the first device and pinned allocations -- are paid by a throw-away call on a tiny problem before the clock starts.
  {
    std::vector<dim_t> tiny(modes.size(), 4);
    cals::Tensor Xw(tiny);
    Xw.randomize();
    cals::Ktensor kw(2, tiny);
    kw.randomize();
    cals::KtensorQueue qw;
    qw.emplace(kw);
    cals::CalsParams pw;
    pw.max_iterations = 2;
    pw.buffer_size = 2;
    pw.devices = devices;
    cals::cp_cals(Xw, qw, pw);
  }

  cals::KtensorQueue queue;
  for (auto &kt : cals_input)
    queue.emplace(kt);
  cals::Timer t_cals;
  t_cals.start();
  const cals::CalsReport rep = cals::cp_cals(X, queue, cp);
  t_cals.stop();
  dim_t model_iters = 0;
  double best_fit = 0.0;
  for (auto &kt : cals_input) {
    model_iters += kt.get_iters();
    best_fit = std::max(best_fit, kt.get_fit());
  }
  cout << "CALS: " << rep.iter << " iterations of the concurrent loop, " << model_iters << " model-iterations, "
       << t_cals.get_time() << " s  (" << model_iters / t_cals.get_time() << " model-iterations/s, device loop "
       << rep.device_ms << " ms, best fit " << best_fit << ")" << endl;

  double als_time = 0.0;
  if (run_als) {
    cals::AlsParams ap;
    ap.max_iterations = cp.max_iterations;
    ap.tol = cp.tol;
    ap.force_max_iter = cp.force_max_iter;
    ap.suppress_lut_warning = true;
    ap.device = devices.empty() ? 0 : devices[0];
    cals::Timer t_als;
    t_als.start();
    cals::cp_omp_als(X, als_input, ap);
    t_als.stop();
    als_time = t_als.get_time();
    double worst = 0.0;
    for (size_t i = 0; i < als_input.size(); i++)
      worst = std::max(worst, std::fabs(als_input[i].get_fit() - cals_input[i].get_fit()));
    cout << "ALS:  " << als_time << " s, largest |fit(ALS) - fit(CALS)| = " << worst << endl;
  }

  cout << "======================================================================" << endl;
  if (run_als)
    cout << "ALS time: " << als_time << endl;
  cout << "CALS time: " << t_cals.get_time() << endl;
  if (run_als)
    cout << "Speedup: " << als_time / t_cals.get_time() << endl;
  return 0;
}
