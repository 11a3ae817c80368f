// Probe for "what comes next" (DESIGN.md section 9): FP64-accurate contraction tiles on the tcgen05 INTEGER tensor cores.
// Not a product path.  Three questions, answered on the B200 itself:
//   1. does a hand-written tcgen05.mma.kind::i8 tile (TMA SWIZZLE_128B operands, INT32 accumulators in TMEM) give exact
//      integer results?                                                                      (check "gemm")
//   2. does the error-free slicing scheme (Ozaki): 7 signed 7-bit slices per FP64 operand, the 28 slice products with
//      i + j <= 6 accumulated per level i + j in 7 TMEM accumulators, reproduce the FP64 inner contraction
//      sum_p X[m,p] * A[c,p] of an MTTKRP stage (K = 200) to ~2^-49 of |X|max |A|max K?        (check "ozaki")
//   3. how fast is the INT8 tensor pipe of this chip?                                        (check "rate")
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/i8_probe tools/i8_probe.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int TM = 128;       // rows of the accumulator tile (TMEM lanes)
constexpr int KB = 256;       // bytes of K per operand row in shared memory: two 128-byte swizzle atoms
constexpr int SLICES = 7;     // signed 7-bit slices per FP64 value
constexpr long long SPIN = 2000000000ll; // bounded waits: a wrong descriptor must not hang the box

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// returns false when the barrier did not flip within SPIN clocks
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done && clock64() - t0 > SPIN)
      return false;
  }
  return true;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// shared-memory matrix descriptor: K-major operand, SWIZZLE_128B, rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);       // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8 rows x 128 B, bits [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B
  return d;
}
// instruction descriptor of kind::i8: S8 x S8 -> S32, both operands K-major, M x N tile
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct ProbeMaps {
  CUtensorMap A; // int8 [planes * 128 rows][KB], box [128 B x 128 rows], SWIZZLE_128B
  CUtensorMap B; // int8 [planes * N rows][KB],   box [128 B x N rows]
};

// One CTA, 128 threads.  For every (ia, ib) in `pairs` (n_pairs of them; level[q] = TMEM accumulator it adds to):
// load plane ia of A and plane ib of B, run the K loop (k_steps MMAs of 32 bytes of K), then the next pair.
// Afterwards every thread reads its row of all n_levels accumulators:
//   mode 0: raw INT32 of level 0 -> out_i32[128][N]
//   mode 1: FP64 recombination sum_s D_s * 2^(-7 (s + 2)) * 2^(ea[row] + eb[col]) -> out_f64[128][N]
template <int N>
__global__ void __launch_bounds__(128, 1)
tile_kernel(const __grid_constant__ ProbeMaps maps, const int *pairs, const int *level, int n_pairs, int n_levels,
            int k_steps, int mode, int32_t *out_i32, double *out_f64, const int *ea, const int *eb, int *status,
            int repeat) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u); // swizzle atoms want 1024-byte alignment
  unsigned char *As = smem;                    // 2 atoms x [128][128 B]
  unsigned char *Bs = smem + 2 * TM * 128;     // 2 atoms x [N][128 B]
  uint64_t *bars = (uint64_t *)(Bs + 2 * N * 128);
  uint64_t *bar_tma = bars, *bar_mma = bars + 1;
  uint32_t *tmem_slot = (uint32_t *)(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 512;

  if (threadIdx.x == 0) {
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  __shared__ int failed;
  if (threadIdx.x == 0)
    failed = 0;
  __syncthreads();

  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_i8(TM, N);
    uint32_t ph_tma = 0, ph_mma = 0;
    unsigned used = 0; // bit s: level s already holds a product
    for (int rep = 0; rep < repeat && !failed; rep++) {
      for (int q = 0; q < n_pairs; q++) {
        const int ia = pairs[2 * q], ib = pairs[2 * q + 1], lv = level[q];
        mbar_expect_tx(bar_tma, 2 * TM * 128 + 2 * N * 128);
        tma_load_2d(As, &maps.A, bar_tma, 0, ia * TM);
        tma_load_2d(As + TM * 128, &maps.A, bar_tma, 128, ia * TM);
        tma_load_2d(Bs, &maps.B, bar_tma, 0, ib * N);
        tma_load_2d(Bs + N * 128, &maps.B, bar_tma, 128, ib * N);
        if (!mbar_wait(bar_tma, ph_tma)) { failed = 1; break; }
        ph_tma ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int k = 0; k < k_steps; k++) {
          const uint32_t off = (uint32_t)(k >> 2) * (TM * 128) + (uint32_t)(k & 3) * 32;
          const uint32_t offb = (uint32_t)(k >> 2) * (N * 128) + (uint32_t)(k & 3) * 32;
          const uint64_t ad = umma_desc_sw128(smem_u32(As) + off), bd = umma_desc_sw128(smem_u32(Bs) + offb);
          umma_i8(tmem + (uint32_t)lv * N, ad, bd, idesc, (k > 0 || ((used >> lv) & 1u)) ? 1u : 0u);
        }
        used |= 1u << lv;
        umma_commit(bar_mma); // the operands may be overwritten once these MMAs have read them
        if (!mbar_wait(bar_mma, ph_mma)) { failed = 2; break; }
        ph_mma ^= 1;
      }
    }
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (failed) {
    if (threadIdx.x == 0)
      status[0] = failed;
  } else {
    const int row = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < N; c0 += 16) {
      if (mode == 0) {
        int32_t v[16];
        tmem_ld16(tmem + lane_base + c0, v);
        for (int u = 0; u < 16; u++)
          out_i32[row * N + c0 + u] = v[u];
      } else {
        double t[16];
        for (int u = 0; u < 16; u++)
          t[u] = 0.0;
        for (int s = n_levels - 1; s >= 0; s--) { // small terms first
          int32_t v[16];
          tmem_ld16(tmem + lane_base + (uint32_t)s * N + c0, v);
          const double w = exp2(-7.0 * (s + 2));
          for (int u = 0; u < 16; u++)
            t[u] += (double)v[u] * w;
        }
        for (int u = 0; u < 16; u++)
          out_f64[row * N + c0 + u] = ldexp(t[u], ea[row] + eb[c0 + u]);
      }
    }
    if (threadIdx.x == 0)
      status[0] = 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

// Raw pipe rate: every CTA (one per SM) keeps issuing MMAs on the operands already in shared memory.
template <int N>
__global__ void __launch_bounds__(128, 1)
rate_kernel(const __grid_constant__ ProbeMaps maps, int iters, int *status) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *As = smem;
  unsigned char *Bs = smem + 2 * TM * 128;
  uint64_t *bars = (uint64_t *)(Bs + 2 * N * 128);
  uint64_t *bar_tma = bars, *bar_mma = bars + 1;
  uint32_t *tmem_slot = (uint32_t *)(bars + 2);
  const int warp = threadIdx.x >> 5;
  constexpr uint32_t TMEM_COLS = 512;
  if (threadIdx.x == 0) {
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    bool ok = true;
    mbar_expect_tx(bar_tma, 2 * TM * 128 + 2 * N * 128);
    tma_load_2d(As, &maps.A, bar_tma, 0, 0);
    tma_load_2d(As + TM * 128, &maps.A, bar_tma, 128, 0);
    tma_load_2d(Bs, &maps.B, bar_tma, 0, 0);
    tma_load_2d(Bs + N * 128, &maps.B, bar_tma, 128, 0);
    ok = mbar_wait(bar_tma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = idesc_i8(TM, N);
    uint64_t ad[8], bd[8];
    for (int k = 0; k < 8; k++) {
      ad[k] = umma_desc_sw128(smem_u32(As) + (uint32_t)(k >> 2) * (TM * 128) + (uint32_t)(k & 3) * 32);
      bd[k] = umma_desc_sw128(smem_u32(Bs) + (uint32_t)(k >> 2) * (N * 128) + (uint32_t)(k & 3) * 32);
    }
    if (ok) {
      for (int it = 0; it < iters; it++) {
        const uint32_t d = tmem + (uint32_t)((it & 1) * N); // alternate two accumulators
#pragma unroll
        for (int k = 0; k < 8; k++)
          umma_i8(d, ad[k], bd[k], idesc, 1u);
      }
      umma_commit(bar_mma);
      ok = mbar_wait(bar_mma, 0);
    }
    if (!ok && blockIdx.x == 0)
      status[0] = 3;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

static PFN_encodeTiled get_encode() {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !fn) { printf("{\"error\": \"no cuTensorMapEncodeTiled\"}\n"); exit(1); }
  return (PFN_encodeTiled)fn;
}
static void make_map(PFN_encodeTiled enc, CUtensorMap *m, void *base, int rows, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)KB, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)KB};
  cuuint32_t box[2] = {128, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("{\"error\": \"cuTensorMapEncodeTiled %d\"}\n", (int)r); exit(1); }
}

// error-free slicing of one FP64 vector (a row of X or a column of A) of length K into SLICES int8 planes:
// x = 2^e * sum_i s_i * 2^(-7 (i + 1)) + (remainder below 2^(e - 49)),  |s_i| <= 127
static int slice_row(const double *x, int K, int8_t *planes, int plane_stride) {
  double mx = 0.0;
  for (int k = 0; k < K; k++) mx = std::fmax(mx, std::fabs(x[k]));
  int e = 0;
  if (mx > 0.0) { std::frexp(mx, &e); } // mx = f * 2^e, f in [0.5, 1)  ->  |x| * 2^-e < 1
  for (int k = 0; k < K; k++) {
    double r = std::ldexp(x[k], -e);
    for (int i = 0; i < SLICES; i++) {
      r *= 128.0;
      const double s = std::trunc(r);
      planes[(size_t)i * plane_stride + k] = (int8_t)s;
      r -= s;
    }
  }
  return e;
}

template <int N> static size_t tile_smem() { return 2 * TM * 128 + 2 * N * 128 + 64 + 1024; }

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  PFN_encodeTiled enc = get_encode();
  constexpr int N = 64;
  int *d_status; CK(cudaMalloc(&d_status, 16)); CK(cudaMemset(d_status, 0xff, 16));
  printf("{\"gpu\": \"%s\", \"sms\": %d}\n", prop.name, prop.multiProcessorCount);

  // ---------------------------------------------------------------- 1. plain integer tile
  {
    std::vector<int8_t> A((size_t)TM * KB), B((size_t)N * KB);
    srand(1);
    for (auto &v : A) v = (int8_t)(rand() % 255 - 127);
    for (auto &v : B) v = (int8_t)(rand() % 255 - 127);
    int8_t *dA, *dB; int32_t *dD; int *d_pairs, *d_level;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, TM * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    const int pairs[2] = {0, 0}, level[1] = {0};
    CK(cudaMalloc(&d_pairs, 8)); CK(cudaMalloc(&d_level, 4));
    CK(cudaMemcpy(d_pairs, pairs, 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_level, level, 4, cudaMemcpyHostToDevice));
    ProbeMaps maps; make_map(enc, &maps.A, dA, TM, TM); make_map(enc, &maps.B, dB, N, N);
    CK(cudaFuncSetAttribute(tile_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem<N>()));
    tile_kernel<N><<<1, 128, tile_smem<N>()>>>(maps, d_pairs, d_level, 1, 1, KB / 32, 0, dD, nullptr, nullptr, nullptr, d_status, 1);
    CK(cudaDeviceSynchronize());
    int st; CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
    std::vector<int32_t> D(TM * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int m = 0; m < TM; m++)
      for (int n = 0; n < N; n++) {
        int32_t ref = 0;
        for (int k = 0; k < KB; k++) ref += (int32_t)A[(size_t)m * KB + k] * (int32_t)B[(size_t)n * KB + k];
        bad += ref != D[m * N + n];
      }
    printf("{\"check\": \"gemm\", \"tile\": \"128x%dx%d s8 x s8 -> s32\", \"status\": %d, \"mismatches\": %lld, \"of\": %d}\n", N, KB, st, bad, TM * N);
    if (st != 0 || bad != 0) return 0; // the later checks build on this one
  }

  // ---------------------------------------------------------------- 2. FP64 inner contraction through 7 x 7-bit slices
  {
    const int K = 200; // contracted extent of BASELINE config 2
    std::vector<double> X((size_t)TM * K), Af((size_t)N * K);
    srand(2);
    for (auto &v : X) v = 2.0 * rand() / RAND_MAX - 1.0;
    for (auto &v : Af) v = (2.0 * rand() / RAND_MAX - 1.0) * std::ldexp(1.0, rand() % 7 - 3);
    std::vector<int8_t> PA((size_t)SLICES * TM * KB, 0), PB((size_t)SLICES * N * KB, 0);
    std::vector<int> ea(TM), eb(N);
    for (int m = 0; m < TM; m++) ea[m] = slice_row(&X[(size_t)m * K], K, &PA[(size_t)m * KB], TM * KB);
    for (int n = 0; n < N; n++) eb[n] = slice_row(&Af[(size_t)n * K], K, &PB[(size_t)n * KB], N * KB);
    std::vector<int> pairs, level;
    for (int i = 0; i < SLICES; i++)
      for (int j = 0; i + j < SLICES; j++) { pairs.push_back(i); pairs.push_back(j); level.push_back(i + j); }
    int8_t *dA, *dB; double *dT; int *d_pairs, *d_level, *d_ea, *d_eb;
    CK(cudaMalloc(&dA, PA.size())); CK(cudaMalloc(&dB, PB.size())); CK(cudaMalloc(&dT, TM * N * 8));
    CK(cudaMalloc(&d_pairs, pairs.size() * 4)); CK(cudaMalloc(&d_level, level.size() * 4));
    CK(cudaMalloc(&d_ea, TM * 4)); CK(cudaMalloc(&d_eb, N * 4));
    CK(cudaMemcpy(dA, PA.data(), PA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, PB.data(), PB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_pairs, pairs.data(), pairs.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_level, level.data(), level.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ea, ea.data(), TM * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_eb, eb.data(), N * 4, cudaMemcpyHostToDevice));
    ProbeMaps maps; make_map(enc, &maps.A, dA, SLICES * TM, TM); make_map(enc, &maps.B, dB, SLICES * N, N);
    CK(cudaMemset(d_status, 0xff, 16));
    tile_kernel<N><<<1, 128, tile_smem<N>()>>>(maps, d_pairs, d_level, (int)level.size(), SLICES, (K + 31) / 32, 1, nullptr, dT, d_ea, d_eb, d_status, 1);
    CK(cudaDeviceSynchronize());
    int st; CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
    std::vector<double> T(TM * N);
    CK(cudaMemcpy(T.data(), dT, T.size() * 8, cudaMemcpyDeviceToHost));
    double worst_scaled = 0.0, worst_rel = 0.0;
    for (int m = 0; m < TM; m++)
      for (int n = 0; n < N; n++) {
        long double ref = 0.0L, mag = 0.0L;
        for (int k = 0; k < K; k++) {
          ref += (long double)X[(size_t)m * K + k] * (long double)Af[(size_t)n * K + k];
          mag += std::fabs((long double)X[(size_t)m * K + k] * (long double)Af[(size_t)n * K + k]);
        }
        const double err = (double)std::fabs((long double)T[m * N + n] - ref);
        worst_scaled = std::fmax(worst_scaled, err / (double)mag);      // relative to sum |x||a| (what FP64 rounding is measured against)
        worst_rel = std::fmax(worst_rel, err / std::fabs((double)ref));  // relative to the result itself (cancellation included)
      }
    printf("{\"check\": \"ozaki\", \"K\": %d, \"slices\": %d, \"products\": %d, \"status\": %d, \"worst_err_over_sum_abs\": %.3e, "
           "\"worst_err_over_abs_result\": %.3e, \"fp64_eps\": %.3e}\n", K, SLICES, (int)level.size(), st, worst_scaled, worst_rel, 2.22e-16);
  }

  // ---------------------------------------------------------------- 3. INT8 tensor pipe rate
  {
    int8_t *dA, *dB;
    CK(cudaMalloc(&dA, (size_t)TM * KB)); CK(cudaMalloc(&dB, (size_t)256 * KB));
    CK(cudaMemset(dA, 1, (size_t)TM * KB)); CK(cudaMemset(dB, 1, (size_t)256 * KB));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20000, sms = prop.multiProcessorCount;
    auto run = [&](auto kern, int n, const char *name) {
      ProbeMaps maps; make_map(enc, &maps.A, dA, TM, TM); make_map(enc, &maps.B, dB, n, n);
      const size_t smem = 2 * TM * 128 + 2 * n * 128 + 64 + 1024;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CK(cudaMemset(d_status, 0, 16));
      kern<<<sms, 128, smem>>>(maps, 100, d_status); CK(cudaDeviceSynchronize());
      float best = 1e30f;
      for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0)); kern<<<sms, 128, smem>>>(maps, iters, d_status); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::fmin(best, ms);
      }
      int st; CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
      const double ops = 2.0 * TM * n * 32.0 * 8.0 * iters * sms;
      printf("{\"check\": \"rate\", \"mma\": \"%s\", \"status\": %d, \"ms\": %.3f, \"int8_tops\": %.1f, \"fp64_equiv_tflops_at_28_products\": %.1f}\n",
             name, st, best, ops / (best * 1e-3) / 1e12, ops / 28.0 / (best * 1e-3) / 1e12);
    };
    run(rate_kernel<64>, 64, "128x64x32");
    run(rate_kernel<128>, 128, "128x128x32");
    run(rate_kernel<256>, 256, "128x256x32");
  }
  return 0;
}
