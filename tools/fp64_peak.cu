// Micro-benchmark: FP64 peak of one B200 through (a) mma.sync DMMA (m8n8k4 and m16n8k16 forms) and (b) DFMA.
// Its result is the roofline denominator for the MTTKRP kernel (MEASURED_PEAKS.json has no FP64 entry).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void k_dmma884(double *out, int iters, double seed) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c[i][0] = 0; c[i][1] = 0; }
  double a = seed + threadIdx.x * 1e-3, b = seed - threadIdx.x * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_dmma16816(double *out, int iters, double seed) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
  double a[8], b[4];
  for (int j = 0; j < 8; j++) a[j] = seed + threadIdx.x * 1e-3 + j;
  for (int j = 0; j < 4; j++) b[j] = seed - threadIdx.x * 1e-3 + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_dfma(double *out, int iters, double seed) {
  double c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) c[i] = i;
  double a = seed + threadIdx.x * 1e-9, b = 1e-9 * seed;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// latency probe: one warp, dependent chain of DMMA, cycles per op
__global__ void k_dmma_lat(double *out, long long *cyc, int iters) {
  double c0 = 0, c1 = 0, a = 1.0 + threadIdx.x, b = 0.5;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) dmma884(c0, c1, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = c0 + c1;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <typename F>
double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  double *out; CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024 * 8));
  long long *cyc; CK(cudaMalloc(&cyc, 64));
  const int iters = 20000;
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    int threads = warps * 32; if (threads > 1024) continue;
    for (int ctas : {1, 2}) {
      int grid = sms * ctas;
      double ms, tf;
#define RUN884(ILP) ms = time_ms([&] { k_dmma884<ILP><<<grid, threads>>>(out, iters, 1.0); }, 3); \
      tf = (double)grid * warps * iters * ILP * 512.0 / (ms * 1e-3) / 1e12; \
      printf("{\"kernel\": \"dmma_m8n8k4\", \"ilp\": %d, \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", ILP, warps, ctas, ms, tf);
      RUN884(1) RUN884(2) RUN884(4) RUN884(8) RUN884(16)
#define RUN16816(ILP) ms = time_ms([&] { k_dmma16816<ILP><<<grid, threads>>>(out, iters / 8, 1.0); }, 3); \
      tf = (double)grid * warps * (iters / 8) * ILP * 4096.0 / (ms * 1e-3) / 1e12; \
      printf("{\"kernel\": \"dmma_m16n8k16\", \"ilp\": %d, \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", ILP, warps, ctas, ms, tf);
      RUN16816(1) RUN16816(4) RUN16816(8)
#define RUNFMA(ILP) ms = time_ms([&] { k_dfma<ILP><<<grid, threads>>>(out, iters, 1.0); }, 3); \
      tf = (double)grid * threads * (double)iters * ILP * 2.0 / (ms * 1e-3) / 1e12; \
      printf("{\"kernel\": \"dfma\", \"ilp\": %d, \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", ILP, warps, ctas, ms, tf);
      RUNFMA(1) RUNFMA(4) RUNFMA(8)
    }
  }
  // sustained: 2 s of the best config
  {
    int grid = sms * 2, threads = 512;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 0;
    for (; n < 200; n++) k_dmma884<8><<<grid, threads>>>(out, iters * 4, 1.0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double tf = (double)grid * 16 * (iters * 4.0) * 8 * 512.0 * n / (ms * 1e-3) / 1e12;
    printf("{\"kernel\": \"dmma_m8n8k4_sustained\", \"seconds\": %.3f, \"tflops\": %.3f}\n", ms * 1e-3, tf);
    CK(cudaEventRecord(e0));
    for (n = 0; n < 200; n++) k_dfma<8><<<grid, threads>>>(out, iters * 4, 1.0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    tf = (double)grid * threads * (iters * 4.0) * 8 * 2.0 * n / (ms * 1e-3) / 1e12;
    printf("{\"kernel\": \"dfma_sustained\", \"seconds\": %.3f, \"tflops\": %.3f}\n", ms * 1e-3, tf);
  }
  k_dmma_lat<<<1, 32>>>(out, cyc, 10000);
  long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("{\"kernel\": \"dmma_m8n8k4_latency\", \"cycles_per_dependent_op\": %.2f}\n", h / 10000.0);
  return 0;
}
