// Relational tests of the cals:: C++ surface on the B200 path, restating what the reference pins in
// tests/cals/test_cals.cpp (no golden vectors exist upstream: CALS must equal ALS model by model, and the jackknife
// flavours must equal ALS on explicitly row-deleted tensors), plus checks of the report and of the option handling.
// Needs a B200: run by pytest -m gpu (tests/test_cpp_api.py).
#include <algorithm>
#include <cmath>
#include <numeric>
#include <random>

#include "gtest/gtest.h"

#include "als.h"
#include "cals.h"
#include "multi_ktensor.h"

using cals::Ktensor;
using cals::Tensor;

namespace {

constexpr double kModelTol = 1e-11; // reference: MODEL_DIFF_ACC, tests/cals/test_cals.cpp:7

struct Uniform {
  std::mt19937 gen;
  std::uniform_real_distribution<double> dist{-1.0, 1.0};
  explicit Uniform(unsigned seed) : gen(seed) {}
  function<double()> source() {
    return [this] { return dist(gen); };
  }
};

double reconstruction_gap(Ktensor &a, Ktensor &b) {
  Tensor ta = a.to_tensor(), tb = b.to_tensor();
  for (dim_t e = 0; e < ta.get_n_elements(); e++)
    ta[e] -= tb[e];
  return ta.norm();
}

std::vector<Ktensor> random_models(const std::vector<dim_t> &ranks, const std::vector<dim_t> &modes, Uniform &u) {
  std::vector<Ktensor> out;
  out.reserve(ranks.size());
  for (dim_t r : ranks) {
    out.emplace_back(r, modes);
    out.back().fill(u.source());
  }
  return out;
}

cals::KtensorQueue queue_of(std::vector<Ktensor> &v) {
  cals::KtensorQueue q;
  for (Ktensor &k : v)
    q.emplace(k);
  return q;
}

} // namespace

// reference CalsGeneralTests.SimpleCorrectness (tests/cals/test_cals.cpp:13-86): a queue much longer than the buffer,
// tol-based stopping, models of many ranks -- every model must come out as cp_als alone produces it.
TEST(Cals, QueueOfModelsEqualsOneByOneAls) {
  std::vector<dim_t> ranks;
  for (dim_t r = 1; r <= 12; r++)
    ranks.insert(ranks.end(), 12, r);
  std::mt19937 order(7);
  std::shuffle(ranks.begin(), ranks.end(), order);

  Uniform u(0);
  Ktensor P(10, {13, 12, 11});
  P.fill(u.source());
  Tensor T = P.to_tensor();
  std::vector<Ktensor> start = random_models(ranks, T.get_modes(), u);
  std::vector<Ktensor> by_cals(start), by_als(start), by_omp(start);

  cals::CalsParams cp;
  cp.max_iterations = 1000;
  cp.tol = 1e-5;
  cp.buffer_size = 30;
  cals::AlsParams ap;
  ap.max_iterations = cp.max_iterations;
  ap.tol = cp.tol;
  ap.suppress_lut_warning = true;

  auto q = queue_of(by_cals);
  const cals::CalsReport rep = cals::cp_cals(T, q, cp);
  EXPECT_TRUE(q.empty());
  EXPECT_EQ((size_t)rep.n_ktensors, ranks.size());
  EXPECT_EQ((dim_t)rep.ktensor_comp_sum, std::accumulate(ranks.begin(), ranks.end(), dim_t(0)));
  for (Ktensor &k : by_als)
    cals::cp_als(T, k, ap);
  cals::cp_omp_als(T, by_omp, ap);

  for (size_t p = 0; p < ranks.size(); p++) {
    EXPECT_NEAR(reconstruction_gap(by_als[p], by_cals[p]), 0.0, kModelTol);
    EXPECT_NEAR(reconstruction_gap(by_cals[p], by_omp[p]), 0.0, kModelTol);
    EXPECT_EQ(by_als[p].get_iters(), by_cals[p].get_iters());
  }
}

// reference CalsLineSearchTests.LineSearchCorrectness (tests/cals/test_cals.cpp:88-179), both methods: with line search
// switched on, the queue of models through cp_cals equals cp_als model by model.
static void line_search_case(cals::ls::LS_METHOD method) {
  std::vector<dim_t> ranks;
  for (dim_t r = 1; r <= 12; r++)
    ranks.insert(ranks.end(), 6, r);
  std::mt19937 order(9);
  std::shuffle(ranks.begin(), ranks.end(), order);
  Uniform u(0);
  Ktensor P(10, {13, 12, 11});
  P.fill(u.source());
  Tensor T = P.to_tensor();
  std::vector<Ktensor> start = random_models(ranks, T.get_modes(), u);
  std::vector<Ktensor> by_cals(start), by_als(start);

  cals::CalsParams cp;
  cp.max_iterations = 1000;
  cp.tol = 1e-5;
  cp.buffer_size = 30;
  cp.line_search = true;
  cp.line_search_interval = 10;
  cp.line_search_step = 0;
  cp.line_search_method = method;
  cals::AlsParams ap;
  ap.max_iterations = cp.max_iterations;
  ap.tol = cp.tol;
  ap.line_search = true;
  ap.line_search_interval = cp.line_search_interval;
  ap.line_search_step = cp.line_search_step;
  ap.line_search_method = method;
  ap.suppress_lut_warning = true;

  auto q = queue_of(by_cals);
  const cals::CalsReport rep = cals::cp_cals(T, q, cp);
  dim_t als_performed = 0;
  for (Ktensor &k : by_als)
    als_performed += cals::cp_als(T, k, ap).ls_performed;
  EXPECT_GT(rep.ls_performed, (dim_t)0);
  EXPECT_EQ(rep.ls_performed, als_performed);
  for (size_t p = 0; p < ranks.size(); p++) {
    EXPECT_NEAR(reconstruction_gap(by_als[p], by_cals[p]), 0.0, kModelTol);
    EXPECT_EQ(by_als[p].get_iters(), by_cals[p].get_iters());
  }
}
TEST(Cals, LineSearchWithoutErrorCheckingEqualsOneByOneAls) { line_search_case(cals::ls::NO_ERROR_CHECKING); }
TEST(Cals, LineSearchWithErrorCheckingEqualsOneByOneAls) { line_search_case(cals::ls::ERROR_CHECKING_SERIAL); }

// reference CalsJackknifingTests.LogicCorrectness (tests/cals/test_cals.cpp:181-297): a model flagged "leave sample i
// out" inside cp_cals must equal cp_als on the tensor with row i of mode 0 deleted.
TEST(Cals, FlaggedModelsEqualAlsOnRowDeletedTensors) {
  Uniform u(5489u);
  const std::vector<dim_t> modes{20, 9, 12};
  const dim_t R = 5, samples = modes[0];
  Ktensor P(R, modes);
  P.fill(u.source());
  Tensor T = P.to_tensor();
  Ktensor start(R, modes);
  start.fill(u.source());

  std::vector<dim_t> small(modes);
  small[0] -= 1;
  std::vector<Tensor> T_without(samples);
  std::vector<Ktensor> als_models(samples), cals_models(samples);
  for (dim_t i = 0; i < samples; i++) {
    T_without[i] = Tensor(small);
    for (dim_t j = 0; j < modes[1] * modes[2]; j++)
      for (dim_t r = 0, w = 0; r < modes[0]; r++)
        if (r != i)
          T_without[i][j * small[0] + w++] = T[j * modes[0] + r];
    als_models[i] = Ktensor(R, small);
    als_models[i].get_lambda() = start.get_lambda();
    for (dim_t n = 0; n < 3; n++)
      for (dim_t c = 0; c < R; c++)
        for (dim_t r = 0, w = 0; r < modes[n]; r++)
          if (n != 0 || r != i)
            als_models[i].get_factor(n)(w++, c) = start.get_factor(n)(r, c);
    cals_models[i] = Ktensor(R, modes);
    cals_models[i].copy(start);
    cals_models[i].to_jk(0, i);
    cals_models[i].set_jk_fiber(0.0);
  }
  Ktensor full_als(start), full_cals(start);

  cals::CalsParams cp;
  cp.max_iterations = 1000;
  cp.tol = 1e-4;
  cp.buffer_size = 18;
  cp.force_max_iter = true;
  cals::AlsParams ap;
  ap.max_iterations = cp.max_iterations;
  ap.tol = cp.tol;
  ap.force_max_iter = true;
  ap.suppress_lut_warning = true;

  auto q = queue_of(cals_models);
  q.emplace(full_cals);
  cals::cp_cals(T, q, cp);
  for (dim_t i = 0; i < samples; i++)
    cals::cp_als(T_without[i], als_models[i], ap);
  cals::cp_als(T, full_als, ap);

  for (dim_t i = 0; i < samples; i++) {
    Ktensor reduced = cals_models[i].to_regular();
    EXPECT_NEAR(reconstruction_gap(reduced, als_models[i]), 0.0, kModelTol);
  }
  EXPECT_NEAR(reconstruction_gap(full_als, full_cals), 0.0, kModelTol);
}

// reference CalsJackknifingTests.FunctionCorrectness (tests/cals/test_cals.cpp:299-362): jk_cp_cals == jk_cp_als ==
// jk_cp_omp_als, including the post-normalisation and the component matching.
TEST(Cals, JackknifeEntryPointsAgree) {
  Uniform u(5489u);
  const std::vector<dim_t> modes{10, 21, 20};
  const dim_t R = 5;
  Ktensor P(R, modes);
  P.fill(u.source());
  Tensor T = P.to_tensor();
  std::vector<Ktensor> bases = random_models(std::vector<dim_t>(4, R), modes, u);

  cals::CalsParams cp;
  cp.max_iterations = 1000;
  cp.tol = 1e-4;
  cp.buffer_size = 18;
  cp.force_max_iter = true;
  cals::AlsParams ap;
  ap.max_iterations = cp.max_iterations;
  ap.tol = cp.tol;
  ap.force_max_iter = true;
  ap.suppress_lut_warning = true;

  for (Ktensor &b : bases)
    cals::cp_als(T, b, ap);
  cals::JKReport by_cals = cals::jk_cp_cals(T, bases, cp);
  cals::JKReport by_als = cals::jk_cp_als(T, bases, ap);
  cals::JKReport by_omp = cals::jk_cp_omp_als(T, bases, ap);
  ASSERT_EQ(by_cals.results.size(), bases.size());
  for (size_t b = 0; b < bases.size(); b++)
    for (dim_t i = 0; i < modes[0]; i++) {
      Ktensor reduced = by_cals.results[b][i].to_regular();
      EXPECT_NEAR(reconstruction_gap(reduced, by_als.results[b][i]), 0.0, kModelTol);
      EXPECT_NEAR(reconstruction_gap(by_als.results[b][i], by_omp.results[b][i]), 0.0, kModelTol);
      // the left-out row is marked NaN on the way out (reference src/cals.cpp:436)
      for (dim_t c = 0; c < R; c++)
        EXPECT_TRUE(std::isnan(by_cals.results[b][i].get_factor(0)(i, c)));
    }
}

// Sharding the model set over devices must not change any model (each model's arithmetic is independent of its
// neighbours except for the MTTKRP tile partition).  devices = {0, 0} exercises the split/merge logic on one GPU.
TEST(Cals, ShardedModelSetEqualsSingleDevice) {
  std::vector<dim_t> ranks{3, 7, 1, 4, 9, 2, 8, 5, 6, 2, 1, 10};
  Uniform u(11);
  const std::vector<dim_t> modes{24, 15, 18};
  Tensor T(modes);
  T.fill(u.source());
  std::vector<Ktensor> start = random_models(ranks, modes, u);
  std::vector<Ktensor> one(start), two(start);

  cals::CalsParams cp;
  cp.max_iterations = 12;
  cp.force_max_iter = true;
  cp.buffer_size = 20;
  auto q1 = queue_of(one);
  const cals::CalsReport r1 = cals::cp_cals(T, q1, cp);
  cp.devices = {0, 0};
  auto q2 = queue_of(two);
  const cals::CalsReport r2 = cals::cp_cals(T, q2, cp);
  EXPECT_EQ(r1.n_ktensors, r2.n_ktensors);
  EXPECT_EQ(r1.ktensor_comp_sum, r2.ktensor_comp_sum);
  EXPECT_EQ(r2.n_devices, 2);
  for (size_t p = 0; p < ranks.size(); p++) {
    EXPECT_NEAR(reconstruction_gap(one[p], two[p]), 0.0, 1e-9 * T.norm());
    EXPECT_NEAR(one[p].get_fit(), two[p].get_fit(), 1e-9);
    EXPECT_EQ(one[p].get_iters(), two[p].get_iters());
  }
}

// The concurrent MTTKRP driven by hand, as the reference's MTTKRP benchmarks do (include/experiments/
// bench_mttkrp_cals.h): models attached to a MultiKtensor, ONE mttkrp::mttkrp over the multi-factor == the MTTKRP of
// every model on its own.
TEST(Cals, MttkrpOverMultiKtensorEqualsPerModelMttkrp) {
  Uniform u(21);
  std::vector<dim_t> modes{18, 14, 11};
  Tensor T(modes);
  T.fill(u.source());
  std::vector<dim_t> ranks{3, 5, 1, 4};
  std::vector<Ktensor> models = random_models(ranks, modes, u);
  for (dim_t mode = 0; mode < 3; mode++) {
    std::vector<Ktensor> alone(models), together(models);
    std::vector<cals::Matrix> ws;
    cals::mttkrp::MttkrpParams mp;
    for (Ktensor &k : alone)
      cals::mttkrp::mttkrp(T, k, ws, mode, mp);
    cals::MultiKtensor mk(modes, 16);
    for (Ktensor &k : together)
      mk.add(k);
    cals::mttkrp::mttkrp(T, mk, ws, mode, mp);
    EXPECT_EQ(mp.flops, 2ull * T.get_n_elements() * 13);
    for (size_t p = 0; p < models.size(); p++) {
      double worst = 0.0, scale = 0.0;
      for (dim_t e = 0; e < alone[p].get_factor(mode).get_n_elements(); e++) {
        worst = std::max(worst, std::fabs(alone[p].get_factor(mode)[e] - together[p].get_factor(mode)[e]));
        scale = std::max(scale, std::fabs(alone[p].get_factor(mode)[e]));
      }
      EXPECT_LT(worst, 1e-12 * scale);
    }
    for (int id = 1; id <= 4; id++)
      mk.remove(id);
  }
}

// CalsParams::slice_mode: X sliced over two GPUs (needs two devices; skipped otherwise).
TEST(Cals, SlicedTensorOverTwoDevicesEqualsSingleDevice) {
  cals::Tensor probe(std::vector<dim_t>{4, 4, 4});
  probe.zero();
  Uniform u(31);
  const std::vector<dim_t> modes{22, 26, 31};
  Tensor T(modes);
  T.fill(u.source());
  std::vector<dim_t> ranks{2, 6, 3, 9};
  std::vector<Ktensor> start = random_models(ranks, modes, u);
  std::vector<Ktensor> one(start);
  cals::CalsParams cp;
  cp.max_iterations = 6;
  cp.force_max_iter = true;
  cp.buffer_size = 20;
  auto q1 = queue_of(one);
  cals::cp_cals(T, q1, cp);
  for (int s : {2, 1}) {
    std::vector<Ktensor> two(start);
    cp.devices = {0, 1};
    cp.slice_mode = s;
    auto q2 = queue_of(two);
    try {
      cals::cp_cals(T, q2, cp);
    } catch (const cals::B200Error &e) {
      if (std::string(e.what()).find("out of range") != std::string::npos) {
        std::cout << "[  SKIPPED ] needs two GPUs: " << e.what() << std::endl;
        return;
      }
      throw;
    }
    for (size_t p = 0; p < ranks.size(); p++) {
      EXPECT_NEAR(reconstruction_gap(one[p], two[p]), 0.0, 1e-9 * T.norm());
      EXPECT_NEAR(one[p].get_fit(), two[p].get_fit(), 1e-9);
    }
  }
}

TEST(Cals, ReportDescribesTheRun) {
  Uniform u(3);
  const std::vector<dim_t> modes{16, 10, 12};
  Tensor T(modes);
  T.fill(u.source());
  std::vector<dim_t> ranks{4, 4, 4, 4, 4};
  std::vector<Ktensor> models = random_models(ranks, modes, u);
  cals::CalsParams cp;
  cp.max_iterations = 3;
  cp.force_max_iter = true;
  cp.buffer_size = 8; // two models at a time: 3 waves of 3 iterations
  auto q = queue_of(models);
  const cals::CalsReport rep = cals::cp_cals(T, q, cp);
  EXPECT_EQ(rep.iter, (dim_t)9);
  EXPECT_EQ(rep.n_ktensors, 5);
  EXPECT_EQ(rep.ktensor_comp_sum, 20);
  EXPECT_NEAR(rep.X_norm, T.norm(), 1e-12 * T.norm());
  ASSERT_EQ(rep.cols.size(), (size_t)9);
  for (dim_t i = 0; i < 9; i++) {
    EXPECT_EQ(rep.cols[i], (dim_t)(i < 6 ? 8 : 4));
    EXPECT_EQ(rep.flops_per_iteration[i], 2ull * 3 * T.get_n_elements() * rep.cols[i]);
  }
  for (Ktensor &k : models) {
    EXPECT_EQ(k.get_iters(), (dim_t)3);
    EXPECT_NEAR(k.get_fit(), 1 - k.get_approximation_error() / rep.X_norm, 1e-14);
  }
}

TEST(Cals, OptionsOutsideThePathAreRejectedLoudly) {
  Uniform u(4);
  const std::vector<dim_t> modes{6, 5, 4};
  Tensor T(modes);
  T.fill(u.source());
  std::vector<Ktensor> models = random_models({2}, modes, u);
  cals::CalsParams cp;
  auto q = queue_of(models);
  bool thrown = false;
  cp.line_search = true;
  cp.line_search_method = cals::ls::ERROR_CHECKING_PARALLEL; // not dispatched by the reference either
  try {
    cals::cp_cals(T, q, cp);
  } catch (const cals::B200Error &) {
    thrown = true;
  }
  EXPECT_TRUE(thrown);
}

int main(int argc, char **argv) {
  set_threads(4);
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
