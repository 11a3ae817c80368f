// Pair-node contraction T = X_(0)^T A_0 on the tcgen05 INTEGER tensor cores, FP64-accurate through error-free slicing
// (Ozaki scheme): every FP64 operand vector along the contracted mode (a row (i1,i2) of X_(0)^T, a column of A_0) is
// written as  2^e * sum_{i<SL} s_i 2^(-7(i+1)),  s_i signed 7-bit integers (int8); the slice products with
// i + j < SL are exact in INT32 (tcgen05.mma.kind::i8), products of equal weight i + j share one TMEM accumulator
// ("level"), and the epilogue recombines  T = 2^(ea + eb) * sum_s 2^(-7(s+2)) D_s  in FP64.
// With SL = 8 the neglected products are below 2^-56 of |x|max |a|max per term, i.e. below FP64 rounding of the sum.
//
// PROBE ONLY -- not part of the product library: nothing under cp-cals_b200/ includes this file, and libcals_b200.so
// computes every contraction on the FP64 tensor cores (DMMA).  It computes what pair_gemm_kernel (csrc/pairnode.cuh)
// computes and is driven by tools/i8_gemm_probe.cu alone.  Round-1 measurement (profiles/i8_gemm_probe_r01.json): 7 slices
// 0.86 ms against the DMMA kernel's 0.97 ms at BASELINE config 2, before the per-mode re-slicing of the factors -- below
// the 1.15x step gain that would justify a weaker accuracy contract, so the path was not wired into the engine.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "mttkrp.cuh" // smem_u32, mbar_*, tma_load_2d

namespace calsb200 {

#ifndef CALS_I8_SLICES
#define CALS_I8_SLICES 8
#endif
constexpr int I8_SL = CALS_I8_SLICES; // slices per operand == levels (8: below FP64 rounding; 7: 1.6e-14, 28 products)
constexpr int I8_TM = 128;      // rows per tile (TMEM lanes)
constexpr int I8_TN = 64;       // columns per tile: I8_SL * I8_TN = 512 TMEM columns
constexpr int I8_KP = 256;      // bytes of K per operand row: two 128-byte swizzle atoms
constexpr int I8_A_STAGES = 3;  // ring of X slice tiles
constexpr int I8_A_BYTES = I8_TM * I8_KP;
constexpr int I8_B_PLANE_BYTES = I8_TN * I8_KP;
constexpr int I8_SMEM_BYTES = I8_A_STAGES * I8_A_BYTES + I8_SL * I8_B_PLANE_BYTES + 256 + 1024;
constexpr int I8_EPI_WARPS = 8;  // two per TMEM lane quarter, each takes half of the tile's columns
constexpr int I8_THREADS = (2 + I8_EPI_WARPS) * 32; // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue

// ------------------------------------------------------------------------------------------------------------------
// Slicing: one warp per vector.  src[k + ld * v], k < K contiguous; planes[(i * n_vec_pad + v) * I8_KP + k];
// scales[v] = 2^e.
// Bytes k in [K, I8_KP) are never written (the planes are zeroed once at allocation).
__global__ void __launch_bounds__(256)
i8_slice_kernel(const double *__restrict__ src, long long ld, int K, int n_vec, long long n_vec_pad,
                int8_t *__restrict__ planes, double *__restrict__ scales) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int v = warp; v < n_vec; v += n_warps) {
    const double *x = src + (size_t)v * ld;
    double mx = 0.0;
    for (int k = lane; k < K; k += 32)
      mx = fmax(mx, fabs(x[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int e = 0;
    if (mx > 0.0)
      frexp(mx, &e); // mx = f 2^e, f in [0.5, 1): |x| 2^-e < 1
    if (lane == 0)
      scales[v] = ldexp(1.0, e);
    for (int k = lane; k < K; k += 32) {
      double r = ldexp(x[k], -e);
#pragma unroll
      for (int i = 0; i < I8_SL; i++) {
        r *= 128.0;
        const double s = trunc(r);
        planes[((size_t)i * n_vec_pad + v) * I8_KP + k] = (int8_t)(int)s;
        r -= s;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t i8_smem_desc(uint32_t smem_addr) { // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void i8_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// Waiting warps that are not on the critical path poll with a back-off, so that their try_wait traffic does not compete
// with the tensor pipe's shared-memory operand reads.
__device__ __forceinline__ void i8_wait_backoff(uint64_t *bar, uint32_t parity, unsigned ns) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done)
      break;
    __nanosleep(ns);
  }
}

// One lane of a converged warp.  The MMA / TMA issue loops are executed by the WHOLE warp (all values warp-uniform) and
// only the instruction itself is predicated on the elected lane: inside a divergent `if (lane == 0)` the compiler has to
// wrap every descriptor move to the uniform registers in a waterfall loop (~100 clocks per MMA, measured).
__device__ __forceinline__ bool i8_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void i8_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void i8_tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}

struct I8Maps {
  CUtensorMap A; // int8 [I8_SL * R_pad rows][I8_KP], box [128 B x I8_TM rows], SWIZZLE_128B
  CUtensorMap B; // int8 [I8_SL * C_pad rows][I8_KP], box [128 B x I8_TN rows]
};
struct I8Geom {
  int R, K;              // rows of T, contracted extent (<= I8_KP)
  long long R_pad, C_pad; // plane strides in rows (multiples of I8_TM / I8_TN)
  long long ldT;
  const double *sa, *sb; // scales 2^e of the X rows / factor columns
};

// Persistent CTAs, tiles (n-chunk, m-tile) with the m-tile fastest, so that a CTA keeps its factor slices (all I8_SL planes
// of 64 columns, 128 KB) in shared memory over many tiles while the X slice tiles stream through a 3-stage TMA ring.
// One elected thread issues TMA, one issues the MMAs, four warps run the epilogue.  Level s is complete once X plane s
// has been multiplied (its last product is (s, 0)), so the epilogue drains level 0 while the MMAs of planes 1.. are
// still running, and the next tile's first products may overwrite a level as soon as it has been drained: per-level
// full / empty barriers overlap the epilogue with the tensor pipe although the 8 accumulators fill all of TMEM.
// C = active columns (read from *C_ptr when given).
__device__ __forceinline__ double i8_to_double(int32_t v) { // exact, without the conversion pipe
  return __hiloint2double(0x43300000, (int)((unsigned)v ^ 0x80000000u)) - 4503601774854144.0; // 2^52 + 2^31
}

__global__ void __launch_bounds__(I8_THREADS, 1)
pair_gemm_i8_kernel(const __grid_constant__ I8Maps maps, const I8Geom g, const int *__restrict__ C_ptr, int C_fixed,
                    double *__restrict__ T, int dbg = 0) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *As = smem;                                   // [I8_A_STAGES][128 rows x 256 B as two atoms]
  unsigned char *Bs = smem + I8_A_STAGES * I8_A_BYTES;         // [I8_SL][64 rows x 256 B as two atoms]
  uint64_t *bars = (uint64_t *)(Bs + I8_SL * I8_B_PLANE_BYTES);
  uint64_t *a_full = bars, *a_empty = bars + I8_A_STAGES, *b_full = bars + 2 * I8_A_STAGES, *b_empty = b_full + 1;
  uint64_t *l_full = b_empty + 1, *l_empty = l_full + I8_SL;
  uint32_t *tmem_slot = (uint32_t *)(l_empty + I8_SL);

  const int C = C_ptr ? *C_ptr : C_fixed;
  if (C <= 0)
    return;
  const int m_tiles = (g.R + I8_TM - 1) / I8_TM, n_chunks = (C + I8_TN - 1) / I8_TN;
  const long long tiles = (long long)m_tiles * n_chunks;
  // contiguous tile range per CTA: at most two changes of n-chunk (128 KB of factor slices to reload) per CTA
  const long long t_lo = tiles * blockIdx.x / gridDim.x, t_hi = tiles * (blockIdx.x + 1) / gridDim.x;
  if (t_lo >= t_hi)
    return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_steps = (g.K + 31) >> 5;

  if (threadIdx.x == 0) {
    for (int i = 0; i < I8_A_STAGES; i++) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    mbar_init(b_full, 1);
    mbar_init(b_empty, 1);
    for (int s = 0; s < I8_SL; s++) {
      mbar_init(&l_full[s], 1);
      mbar_init(&l_empty[s], I8_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    { // ---------------------------------------------------------------- TMA producer (whole warp, one lane issues)
      int st = 0, prev_nc = -1;
      uint32_t aph = 1, bph = 1; // "empty" barriers: the first pass does not block
      for (long long t = t_lo; t < t_hi; t++) {
        const int nc = (int)(t / m_tiles), mt = (int)(t - (long long)nc * m_tiles);
        if (nc != prev_nc) {
          mbar_wait(b_empty, bph);
          bph ^= 1;
          if (i8_elect_one()) {
            mbar_expect_tx(b_full, I8_SL * I8_B_PLANE_BYTES);
            for (int j = 0; j < I8_SL; j++) {
              const int row = (int)(j * g.C_pad + (long long)nc * I8_TN);
              tma_load_2d(Bs + j * I8_B_PLANE_BYTES, &maps.B, b_full, 0, row);
              tma_load_2d(Bs + j * I8_B_PLANE_BYTES + I8_TN * 128, &maps.B, b_full, 128, row);
            }
          }
          __syncwarp();
          prev_nc = nc;
        }
        for (int i = 0; i < I8_SL; i++) {
          if ((dbg & 2) && (t != t_lo || i >= I8_A_STAGES))
            continue; // measurement aid: X slice tiles are loaded once and reused
          if (dbg & 32)
            i8_wait_backoff(&a_empty[st], aph, 100);
          else
            mbar_wait(&a_empty[st], aph);
          if (i8_elect_one()) {
            mbar_expect_tx(&a_full[st], I8_A_BYTES);
            const int row = (int)(i * g.R_pad + (long long)mt * I8_TM);
            tma_load_2d(As + st * I8_A_BYTES, &maps.A, &a_full[st], 0, row);
            tma_load_2d(As + st * I8_A_BYTES + I8_TM * 128, &maps.A, &a_full[st], 128, row);
          }
          __syncwarp();
          if (++st == I8_A_STAGES) {
            st = 0;
            aph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    { // ---------------------------------------------------------------- MMA issuer (whole warp, one lane issues)
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8_TN >> 3) << 17) | ((uint32_t)(I8_TM >> 4) << 24);
      // descriptors differ only in the start address (16-byte units): base + offset of the K step
      uint64_t a_desc[I8_A_STAGES], b_desc[I8_SL];
      uint32_t ka[8], kb[8];
#pragma unroll
      for (int i = 0; i < I8_A_STAGES; i++)
        a_desc[i] = i8_smem_desc(smem_u32(As + i * I8_A_BYTES));
#pragma unroll
      for (int j = 0; j < I8_SL; j++)
        b_desc[j] = i8_smem_desc(smem_u32(Bs + j * I8_B_PLANE_BYTES));
#pragma unroll
      for (int k = 0; k < 8; k++) {
        ka[k] = ((uint32_t)(k >> 2) * (I8_TM * 128) + (uint32_t)(k & 3) * 32) >> 4;
        kb[k] = ((uint32_t)(k >> 2) * (I8_TN * 128) + (uint32_t)(k & 3) * 32) >> 4;
      }
      int st = 0, prev_nc = -1;
      uint32_t aph = 0, bph = 0, leph = 1;
      for (long long t = t_lo; t < t_hi; t++) {
        const int nc = (int)(t / m_tiles);
        if (nc != prev_nc) {
          mbar_wait(b_full, bph);
          bph ^= 1;
          prev_nc = nc;
        }
#pragma unroll
        for (int i = 0; i < I8_SL; i++) {
          if (!((dbg & 2) && (t != t_lo || i >= I8_A_STAGES)))
            mbar_wait(&a_full[st], aph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t ad0 = a_desc[st];
#pragma unroll
          for (int j = 0; j < I8_SL - i; j++) {
            if (i == 0) { // first product of level j in this tile: the epilogue must have drained the previous tile's
              mbar_wait(&l_empty[j], leph);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t d = tmem + (uint32_t)(i + j) * I8_TN;
            if (i8_elect_one()) {
#pragma unroll
              for (int k = 0; k < 8; k++)
                if (k < k_steps)
                  i8_mma(d, ad0 + ka[k], b_desc[j] + kb[k], idesc, (i > 0 || k > 0) ? 1u : 0u);
            }
            __syncwarp();
          }
          if (i8_elect_one()) {
            i8_commit(&a_empty[st]); // this X slice tile may be overwritten once its MMAs have read it
            i8_commit(&l_full[i]);   // level i has received its last product (i, 0)
          }
          __syncwarp();
          if (++st == I8_A_STAGES) {
            st = 0;
            aph ^= 1;
          }
        }
        leph ^= 1;
        const long long tn = t + 1;
        if (tn >= t_hi || (int)(tn / m_tiles) != nc) {
          if (i8_elect_one())
            i8_commit(b_empty); // the factor slices may be replaced
          __syncwarp();
        }
      }
    }
  } else { // ---------------------------------------------------------- epilogue: TMEM -> FP64 -> T
    const int q = warp & 3;             // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;   // which 32 of the tile's 64 columns
    constexpr int HN = I8_TN / 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t lfph = 0;
    for (long long t = t_lo; t < t_hi; t++) {
      const int nc = (int)(t / m_tiles), mt = (int)(t - (long long)nc * m_tiles);
      const int row = mt * I8_TM + q * 32 + lane;
      const double sa = row < g.R ? g.sa[row] : 0.0;
      double acc[HN];
#pragma unroll
      for (int u = 0; u < HN; u++)
        acc[u] = 0.0;
#pragma unroll
      for (int s = 0; s < I8_SL; s++) {
        if (dbg & 32)
          i8_wait_backoff(&l_full[s], lfph, 100);
        else
          mbar_wait(&l_full[s], lfph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const double w = __longlong_as_double((long long)(1023 - 7 * (s + 2)) << 52); // 2^(-7(s+2))
        int32_t v0[16], v1[16];
        if (!(dbg & 4)) {
          i8_tmem_ld16(tmem + lane_base + (uint32_t)(s * I8_TN + half * HN), v0);
          i8_tmem_ld16(tmem + lane_base + (uint32_t)(s * I8_TN + half * HN + 16), v1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int u = 0; u < 16; u++)
            v0[u] = v1[u] = u;
        }
        // this warp's part of the level is in registers: hand it back to the tensor pipe
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0)
          mbar_arrive(&l_empty[s]);
        if (!(dbg & 1)) {
#pragma unroll
          for (int u = 0; u < 16; u++) {
            acc[u] = fma(i8_to_double(v0[u]), w, acc[u]);
            acc[16 + u] = fma(i8_to_double(v1[u]), w, acc[16 + u]);
          }
        }
      }
      lfph ^= 1;
      if (row < g.R && !(dbg & 1)) {
        const int col0 = nc * I8_TN + half * HN;
        if (dbg & 8) { // measurement aid: keep the arithmetic alive without the stores
          double keep = 0.0;
#pragma unroll
          for (int u = 0; u < HN; u++)
            keep += acc[u];
          if (keep == 1.2345e-300)
            T[0] = keep;
        } else {
#pragma unroll
          for (int u = 0; u < HN; u++)
            if (col0 + u < C)
              __stcs(&T[(size_t)(col0 + u) * g.ldT + row], acc[u] * ((dbg & 16) ? sa : sa * __ldg(g.sb + col0 + u)));
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

} // namespace calsb200
