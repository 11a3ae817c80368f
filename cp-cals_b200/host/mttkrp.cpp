// Single-operation entry points of the cals:: surface that run on the device through the C-ABI hooks:
//   mttkrp::mttkrp                        reference src/utils/mttkrp.cpp:562-614
//   utils::calculate_jackknifing_norms    reference src/utils/utils.cpp:103-152
#include "engine_pool.h"
#include "utils/mttkrp.h"
#include "utils/utils.h"

namespace cals::mttkrp {

Matrix &mttkrp(const Tensor &X, Ktensor &u, std::vector<Matrix> & /*workspace*/, dim_t mode, MttkrpParams &params) {
  detail::EngineHandle &e = detail::engine_for_device(0);
  std::lock_guard<std::mutex> lk(e.mu);
  detail::upload_tensor(e, X, true);
  const dim_t N = u.get_n_modes(), R = u.get_components();
  std::vector<const double *> in(N);
  for (dim_t n = 0; n < N; n++)
    in[n] = u.get_factor(n).get_data();
  Matrix &G = u.get_factor(mode);
  std::vector<double> out(G.get_n_elements());
  detail::check(e, cals_b200_mttkrp(e.ctx, static_cast<int>(mode), R, in.data(), out.data(), CALS_B200_MTTKRP_DMMA, 1,
                                    nullptr),
                "cals_b200_mttkrp");
  std::copy(out.begin(), out.end(), G.get_data());
  params.flops = 2ull * X.get_n_elements() * R;
  params.memops = X.get_n_elements();
  for (dim_t n = 0; n < N; n++)
    params.memops += X.get_modes()[n] * R;
  return G;
}

MttkrpLut read_lookup_table(std::vector<dim_t> const &, int, bool, bool) { return MttkrpLut{}; }

} // namespace cals::mttkrp

namespace cals::utils {

std::vector<double> calculate_jackknifing_norms(Tensor const &tensor) {
  detail::EngineHandle &e = detail::engine_for_device(0);
  std::lock_guard<std::mutex> lk(e.mu);
  detail::upload_tensor(e, tensor, true);
  std::vector<double> out(tensor.get_modes()[0]);
  detail::check(e, cals_b200_jk_norms(e.ctx, out.data()), "cals_b200_jk_norms");
  return out;
}

} // namespace cals::utils
