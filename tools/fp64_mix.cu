// Micro-benchmark: do DMMA (mma.sync m8n8k4 f64) and DFMA share one execution pipe on sm_100a?
// Every CTA runs `wd` warps of independent DMMA chains and `wf` warps of independent DFMA chains at the same time.
// If the two pipes were separate, the mixed run would take max(t_dmma, t_dfma); on a shared pipe it takes their sum.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_mix tools/fp64_mix.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int ILP = 8;

// warps [0, wd) run DMMA, warps [wd, wd + wf) run DFMA; iters_d / iters_f loop trips of ILP independent ops each
__global__ void k_mix(double *out, int wd, int iters_d, int iters_f, double seed) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp < wd) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = 0; c[i][1] = 0; }
    double a = seed + threadIdx.x * 1e-3, b = seed - threadIdx.x * 1e-3;
    for (int it = 0; it < iters_d; it++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) dmma884(c[i][0], c[i][1], a, b);
    }
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  } else {
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = i;
    double a = seed + threadIdx.x * 1e-9, b = 1e-9 * seed;
    for (int it = 0; it < iters_f; it++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static double time_ms(int grid, int threads, double *out, int wd, int itd, int itf) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_mix<<<grid, threads>>>(out, wd, itd, itf, 1.0); CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < 3; r++) {
    CK(cudaEventRecord(e0)); k_mix<<<grid, threads>>>(out, wd, itd, itf, 1.0); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double *out; CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024));
  printf("{\"gpu\": \"%s\", \"sms\": %d}\n", p.name, sms);
  // one DMMA.8x8x4 = 512 flop (16 pipe clocks per SM sub-partition at the measured peak); one warp-wide DFMA = 64 flop.
  // itf = 8 * itd makes both halves the same number of flops.
  for (int w : {4, 8}) {
    const int itd = 20000, itf = 8 * itd;
    const double fl_d = (double)sms * w * itd * ILP * 512.0, fl_f = (double)sms * w * 32.0 * itf * ILP * 2.0;
    const double td = time_ms(sms, 32 * w, out, w, itd, itf);           // DMMA warps only
    const double tf = time_ms(sms, 32 * w, out, 0, itd, itf);           // DFMA warps only
    const double tm = time_ms(sms, 64 * w, out, w, itd, itf);           // both kinds at once
    printf("{\"warps_each\": %d, \"dmma_only_ms\": %.4f, \"dmma_only_tflops\": %.2f, \"dfma_only_ms\": %.4f, "
           "\"dfma_only_tflops\": %.2f, \"mixed_ms\": %.4f, \"mixed_total_tflops\": %.2f, "
           "\"mixed_over_sum\": %.3f, \"mixed_over_max\": %.3f}\n",
           w, td, fl_d / td / 1e9, tf, fl_f / tf / 1e9, tm, (fl_d + fl_f) / tm / 1e9, tm / (td + tf),
           tm / (td > tf ? td : tf));
    // a light DFMA load (the Khatri-Rao scaling: 2 DMUL per 10 DMMA = 1/40 of the DMMA flops)
    const int itl = itd / 5;
    const double tl = time_ms(sms, 64 * w, out, w, itd, itl);
    printf("{\"warps_each\": %d, \"dfma_share_of_flops\": %.4f, \"mixed_light_ms\": %.4f, \"slowdown_vs_dmma_only\": %.4f}\n",
           w, (double)itl * 64.0 / (itd * 512.0), tl, tl / td);
  }
  return 0;
}
