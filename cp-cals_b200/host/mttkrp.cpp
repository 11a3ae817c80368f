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

Matrix &khatri_rao(const Matrix &A, const Matrix &B, Matrix &workspace, KrpParams &params) {
  if (A.get_cols() != B.get_cols())
    throw B200Error("khatri_rao: A and B must have the same number of columns");
  const dim_t IA = A.get_rows(), IB = B.get_rows(), cols = A.get_cols();
  if (workspace.get_max_n_elements() < IA * IB * cols)
    throw B200Error("khatri_rao: workspace smaller than [A.rows * B.rows x cols]");
  workspace.resize(IA * IB, cols);
  // dense copies when a matrix is a strided view (col_stride != rows)
  auto dense = [](const Matrix &M, std::vector<double> &tmp) -> const double * {
    if (M.get_col_stride() == M.get_rows())
      return M.get_data();
    tmp.resize(M.get_rows() * M.get_cols());
    for (dim_t c = 0; c < M.get_cols(); c++)
      std::copy(M.get_data() + c * M.get_col_stride(), M.get_data() + c * M.get_col_stride() + M.get_rows(),
                tmp.data() + c * M.get_rows());
    return tmp.data();
  };
  std::vector<double> ta, tb;
  detail::EngineHandle &e = detail::engine_for_device(0);
  std::lock_guard<std::mutex> lk(e.mu);
  detail::check(e, cals_b200_khatri_rao(e.ctx, dense(A, ta), IA, dense(B, tb), IB, cols, workspace.get_data()),
                "cals_b200_khatri_rao");
  params.flops += 1llu * IA * IB * cols; // reference src/utils/mttkrp.cpp:99-101
  params.memops += 1llu * IA * IB * cols + 1llu * IA * cols + 1llu * IB * cols;
  return workspace;
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
