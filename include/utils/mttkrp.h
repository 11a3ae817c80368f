// MTTKRP options and the single-operation entry point (reference include/utils/mttkrp.h:13-101).
// On the B200 path the method/LUT knobs are accepted and ignored: there is one kernel family (TMA-fed FP64
// tensor-core contraction with the Khatri-Rao rows formed on the fly, cp-cals_b200/csrc/mttkrp.cuh).
#ifndef CALS_B200_UTILS_MTTKRP_H
#define CALS_B200_UTILS_MTTKRP_H

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "ktensor.h"
#include "timer.h"

namespace cals::mttkrp {

typedef std::vector<std::map<int, int>> LUT_v;
struct MttkrpLut {
  LUT_v lut_v{};
  std::vector<int> keys_v{};
};

enum MTTKRP_METHOD { MTTKRP = 0, TWOSTEP0, TWOSTEP1, AUTO, LENGTH };
static const std::string mttkrp_method_names[MTTKRP_METHOD::LENGTH] = {"MTTKRP", "TWOSTEP0", "TWOSTEP1", "AUTO"};

struct KrpParams {
  uint64_t flops{0};
  uint64_t memops{0};
  bool cuda{false};
};

struct MttkrpParams {
  MTTKRP_METHOD method{AUTO};
  KrpParams krp_params{};
  MttkrpLut lut{};
  bool cuda{false};
  MttkrpTimers mttkrp_timers;
  uint64_t flops{0};  // out: 2 * nX * R (the KRP-GEMM count)
  uint64_t memops{0}; // out: algorithmic doubles moved: nX + sum_k I_k * R
};

// G = X_(mode) * KhatriRao(factors of u except `mode`); the result overwrites u's factor `mode` and a reference to it
// is returned (reference src/utils/mttkrp.cpp:562-614).  `workspace` is unused here (nothing is materialised).
cals::Matrix &mttkrp(const cals::Tensor &X, cals::Ktensor &u, std::vector<cals::Matrix> &workspace, dim_t mode,
                     cals::mttkrp::MttkrpParams &params);

// K = A (.) B, the explicit Khatri-Rao product of two matrices with the same number of columns:
// K(ib + B.rows * ia, c) = A(ia, c) * B(ib, c); `workspace` is resized to [A.rows * B.rows x A.cols], filled on the
// device (C ABI cals_b200_khatri_rao) and returned (reference include/utils/mttkrp.h:88-89, src/utils/mttkrp.cpp:78-103).
// cp_cals / mttkrp() never call it -- they form the Khatri-Rao rows on the fly (cp-cals_b200/csrc/mttkrp.cuh).
cals::Matrix &khatri_rao(const cals::Matrix &A, const cals::Matrix &B, cals::Matrix &workspace,
                         cals::mttkrp::KrpParams &params);

// No lookup tables on this path; returns an empty table so that callers that pre-load one keep working.
MttkrpLut read_lookup_table(std::vector<dim_t> const &modes, int threads, bool gpu = false,
                            bool suppress_warning = false);

} // namespace cals::mttkrp
#endif
