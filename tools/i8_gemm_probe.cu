// Stand-alone check of tools/pair_i8_probe.cuh (a probe, not product code) at the size of BASELINE config 2: T = X_(0)^T A_0 with R = 40000 rows, K = 200,
// C = 2100 columns, sliced on the device, contracted on the INT8 tensor cores, compared with long-double dot products
// on sampled entries, timed with CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Icp-cals_b200/csrc -o tools/i8_gemm_probe tools/i8_gemm_probe.cu -lcuda
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pair_i8_probe.cuh"

using namespace calsb200;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void make_map(PFN_encodeTiled enc, CUtensorMap *m, void *base, long long rows, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)I8_KP, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)I8_KP};
  cuuint32_t box[2] = {128, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("{\"error\": \"cuTensorMapEncodeTiled %d\"}\n", (int)r); exit(1); }
}

int main(int argc, char **argv) {
  const int R = argc > 1 ? atoi(argv[1]) : 40000, K = argc > 2 ? atoi(argv[2]) : 200, C = argc > 3 ? atoi(argv[3]) : 2100;
  const int dbg = argc > 4 ? atoi(argv[4]) : 0;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  PFN_encodeTiled enc = (PFN_encodeTiled)fn;
  const int ldX = (K + 1) / 2 * 2, ldA = ldX;
  std::vector<double> X((size_t)ldX * R), A((size_t)ldA * C);
  srand(3);
  for (auto &v : X) v = 2.0 * rand() / RAND_MAX - 1.0;
  for (auto &v : A) v = (2.0 * rand() / RAND_MAX - 1.0) * std::ldexp(1.0, rand() % 9 - 4);
  double *dX, *dA, *dT; CK(cudaMalloc(&dX, X.size() * 8)); CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dT, (size_t)R * C * 8));
  CK(cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
  const long long R_pad = ((long long)R + I8_TM - 1) / I8_TM * I8_TM, C_pad = ((long long)C + I8_TN - 1) / I8_TN * I8_TN;
  int8_t *pX, *pA; double *ea, *eb;
  CK(cudaMalloc(&pX, (size_t)I8_SL * R_pad * I8_KP)); CK(cudaMalloc(&pA, (size_t)I8_SL * C_pad * I8_KP));
  CK(cudaMemset(pX, 0, (size_t)I8_SL * R_pad * I8_KP)); CK(cudaMemset(pA, 0, (size_t)I8_SL * C_pad * I8_KP));
  CK(cudaMalloc(&ea, R * 8)); CK(cudaMalloc(&eb, C * 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms_slice_x, ms_slice_a;
  CK(cudaEventRecord(e0));
  i8_slice_kernel<<<prop.multiProcessorCount * 8, 256>>>(dX, ldX, K, R, R_pad, pX, ea);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms_slice_x, e0, e1));
  CK(cudaEventRecord(e0));
  i8_slice_kernel<<<prop.multiProcessorCount * 8, 256>>>(dA, ldA, K, C, C_pad, pA, eb);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms_slice_a, e0, e1));
  I8Maps maps; make_map(enc, &maps.A, pX, I8_SL * R_pad, I8_TM); make_map(enc, &maps.B, pA, I8_SL * C_pad, I8_TN);
  I8Geom g{}; g.R = R; g.K = K; g.R_pad = R_pad; g.C_pad = C_pad; g.ldT = R; g.sa = ea; g.sb = eb;
  CK(cudaFuncSetAttribute(pair_gemm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM_BYTES));
  pair_gemm_i8_kernel<<<prop.multiProcessorCount, I8_THREADS, I8_SMEM_BYTES>>>(maps, g, nullptr, C, dT, dbg);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    CK(cudaEventRecord(e0));
    pair_gemm_i8_kernel<<<prop.multiProcessorCount, I8_THREADS, I8_SMEM_BYTES>>>(maps, g, nullptr, C, dT, dbg);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::fmin(best, ms);
  }
  std::vector<double> T((size_t)R * C);
  CK(cudaMemcpy(T.data(), dT, T.size() * 8, cudaMemcpyDeviceToHost));
  double worst = 0.0; long long checked = 0;
  for (int smp = 0; smp < 20000; smp++) {
    const int r = smp < 2000 ? (R - 1 - smp % std::min(R, 300)) : rand() % R, c = smp < 2000 ? (C - 1 - smp % std::min(C, 70)) : rand() % C;
    long double ref = 0, mag = 0;
    for (int k = 0; k < K; k++) {
      const long double p = (long double)X[(size_t)r * ldX + k] * (long double)A[(size_t)c * ldA + k];
      ref += p; mag += fabsl(p);
    }
    worst = std::fmax(worst, (double)(fabsl((long double)T[(size_t)c * R + r] - ref) / mag));
    checked++;
  }
  const double flops = 2.0 * R * (double)K * C;
  printf("{\"check\": \"pair_gemm_i8\", \"dbg\": %d, \"R\": %d, \"K\": %d, \"C\": %d, \"slices\": %d, \"ms\": %.4f, \"fp64_equiv_tflops\": %.2f, "
         "\"slice_x_ms\": %.4f, \"slice_a_ms\": %.4f, \"worst_err_over_sum_abs\": %.3e, \"sampled\": %lld}\n",
         dbg, R, K, C, I8_SL, best, flops / (best * 1e-3) / 1e12, ms_slice_x, ms_slice_a, worst, checked);
  return 0;
}
