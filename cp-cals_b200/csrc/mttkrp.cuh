// Concurrent MTTKRP for sm_100a.
//
// Replaces mttkrp::mttkrp / mttkrp_impl / khatri_rao* / mttkrp_twostep* (reference src/utils/mttkrp.cpp:78-614) and the
// cuBLAS + khatri_rao_cuda path (src/utils/khatri_rao.cu:10-55):
//     G[i_n, c] = sum_{i_k, k != n} X[i_0 .. i_{N-1}] * prod_{k != n} A_k[i_k, c]        for all C live columns.
//
// Formulation (fused two-step contraction, Khatri-Rao product never materialised):
//   view X as (P, M, O): P = the memory-contiguous mode p (mode 0, or mode 1 of the mode-swapped copy when n == 0),
//   M = mode n, O = all remaining "outer" modes.  Then
//     G[m, c] = sum_o  w_o[c] * ( sum_p X[p, m, o] * A_p[p, c] ),     w_o[c] = prod_{outer k} A_k[i_k(o), c]
//   The inner sum is an FP64 tensor-core contraction (mma.sync m8n8k4 -> DMMA.8x8x4): the A_p tile stays resident in
//   shared memory for a whole sweep over o, X tiles stream through a TMA + mbarrier ring, and the outer scaling w_o[c]
//   is applied to the B fragments in registers, so the KRP row (p,o) is formed on the fly and never stored.
//
// Parallelisation: stream-K.  The (m-tile, n-tile, K-chunk) space is flattened and cut into gridDim.x equal contiguous
// ranges; each CTA writes one partial tile per (m,n) pair it touches into a workspace and mttkrp_reduce_kernel sums
// the partials in a fixed order (deterministic, no atomics).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace calsb200 {

// ------------------------------------------------------------------------------------------------------------------
// Tile configuration.  K tile = 40 doubles: a dense [rows][40] shared-memory tile has a row pitch of 320 B = 8 mod 16
// doubles, which makes the paired-k LDS.128 fragment loads below bank-conflict free without padding or swizzle.
constexpr int KT = 40;           // K tile along the contiguous mode (multiple of 8, == 8 mod 16)
constexpr int OC = 8;            // outer indices per w-chunk
constexpr int NUM_MMA_WARPS = 8;
constexpr int MTTKRP_THREADS = (NUM_MMA_WARPS + 1) * 32;
constexpr int X_STAGES = 4;

template <int WM, int WN> struct TileCfg {
  static constexpr int M_TILE = 8 * WM;
  static constexpr int N_TILE = 8 * WN * NUM_MMA_WARPS;
  static constexpr int X_STAGE_BYTES = M_TILE * KT * 8;
  static constexpr int B_BYTES = N_TILE * KT * 8;
  static constexpr int W_BYTES = N_TILE * OC * 8;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + B_BYTES + X_STAGES * X_STAGE_BYTES + 2 * W_BYTES + 256;
};

struct MttkrpGeom {
  int mode;    // n
  int In;      // extent of mode n (rows of G)
  int Ip;      // extent of the contiguous mode p
  int p_mode;
  int n_outer;
  int outer_mode[CALS_MAX_OUTER]; // ascending; [0] is the fastest outer mode q
  int outer_dim[CALS_MAX_OUTER];
  int outer_lmul[CALS_MAX_OUTER]; // contribution of the outer index to TMA coordinate 1 (modes between p and n)
  int outer_umul[CALS_MAX_OUTER]; // contribution to TMA coordinate 3 (modes above n)
  int P_tiles; // ceil(Ip / KT)
  int QC;      // ceil(outer_dim[0] / OC)
  int S;       // prod(outer_dim[1..])
  int m_tiles; // ceil(In / M_TILE)
  int ldG;
  int ldF[CALS_MAX_MODES];
};

struct MttkrpMaps {
  CUtensorMap X;     // 4-D view (P, L', M, U) of the tensor copy used for this mode
  CUtensorMap B[2];  // factor p in buffer 0 / 1: 2-D (Ip, buffer_cols)
  CUtensorMap W[2];  // factor q (fastest outer mode) in buffer 0 / 1
};

// ------------------------------------------------------------------------------------------------------------------
// PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(addr), "r"(parity)
                 : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------------------------
// Stream-K bookkeeping, shared by the MTTKRP kernel and the reduce kernel (and unit-tested on the host).
struct StreamK {
  long long total; // chunks over all (m,n) pairs
  int Tp;          // chunks per pair
  int G;           // CTAs
  int kmax;        // workspace slots reserved per pair
  __host__ __device__ static StreamK make(int pairs, int Tp, int G) {
    StreamK s;
    s.Tp = Tp;
    s.total = (long long)pairs * Tp;
    // never more CTAs than chunks: every participating CTA owns a non-empty range, CTAs >= G stay idle
    s.G = (int)(s.total < (long long)G ? s.total : (long long)G);
    if (s.G < 1)
      s.G = 1;
    const long long share = s.total / s.G > 0 ? s.total / s.G : 1;
    long long k = (Tp + share - 1) / share + 1;
    s.kmax = (int)(k < Tp ? k : Tp);
    return s;
  }
  __host__ __device__ long long lo(int b) const { return s_mul(b); }
  __host__ __device__ long long hi(int b) const { return s_mul(b + 1); }
  __host__ __device__ long long s_mul(int b) const { return b >= G ? total : total * b / G; }
  // the CTA whose range contains chunk x
  __host__ __device__ int owner(long long x) const { return (int)(((x + 1) * G - 1) / total); }
  __host__ __device__ int first_cta(int pair) const { return owner((long long)pair * Tp); }
  __host__ __device__ int last_cta(int pair) const { return owner((long long)(pair + 1) * Tp - 1); }
};

// ------------------------------------------------------------------------------------------------------------------
// The DMMA kernel.  288 threads: warps 0..7 consume (LDS + DMMA), warp 8 lane 0 produces (TMA).
template <int WM, int WN>
__global__ void __launch_bounds__(MTTKRP_THREADS, 1)
mttkrp_dmma_kernel(const __grid_constant__ MttkrpMaps maps, const MttkrpGeom g, const SchedState *__restrict__ st,
                   const FactorPtrs fac, double *__restrict__ ws, int C_override) {
  using Cfg = TileCfg<WM, WN>;
  constexpr int M_TILE = Cfg::M_TILE, N_TILE = Cfg::N_TILE;

  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  double *Bs = (double *)smem;                                            // [N_TILE][KT]
  double *Xs = (double *)(smem + Cfg::B_BYTES);                           // [X_STAGES][M_TILE][KT]
  double *Ws = (double *)(smem + Cfg::B_BYTES + X_STAGES * Cfg::X_STAGE_BYTES); // [2][N_TILE][OC]
  uint64_t *bars = (uint64_t *)(smem + Cfg::B_BYTES + X_STAGES * Cfg::X_STAGE_BYTES + 2 * Cfg::W_BYTES);
  uint64_t *full_x = bars, *empty_x = bars + X_STAGES;
  uint64_t *full_w = bars + 2 * X_STAGES, *empty_w = full_w + 2;
  uint64_t *full_b = empty_w + 2, *empty_b = full_b + 1;

  const int C = C_override > 0 ? C_override : st->C;
  const int cur = C_override > 0 ? 0 : st->cur;
  if (C <= 0)
    return;
  const int n_tiles = (C + N_TILE - 1) / N_TILE;
  const int Tp = g.P_tiles * g.S * g.QC;
  const StreamK sk = StreamK::make(g.m_tiles * n_tiles, Tp, gridDim.x);
  const long long lo = sk.lo(blockIdx.x), hi = sk.hi(blockIdx.x);
  if (lo >= hi)
    return;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef CALS_B200_POISON_SMEM
  { // debug aid: any read of shared memory that no TMA has written shows up as NaN in G
    double *all = (double *)smem;
    const int nd = (Cfg::B_BYTES + X_STAGES * Cfg::X_STAGE_BYTES + 2 * Cfg::W_BYTES) / 8;
    for (int i = threadIdx.x; i < nd; i += blockDim.x)
      all[i] = __longlong_as_double(0x7ff8000000000000ll);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
#endif
  if (threadIdx.x == 0) {
    for (int i = 0; i < X_STAGES; i++) {
      mbar_init(&full_x[i], 1);
      mbar_init(&empty_x[i], NUM_MMA_WARPS);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&full_w[i], 1);
      mbar_init(&empty_w[i], NUM_MMA_WARPS);
    }
    mbar_init(full_b, 1);
    mbar_init(empty_b, NUM_MMA_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int Iq = g.outer_dim[0];

  if (warp == NUM_MMA_WARPS) {
    // ===================================== producer =====================================
    if (lane != 0)
      return;
    const CUtensorMap *tmB = &maps.B[cur], *tmW = &maps.W[cur];
    int xs = 0, wsi = 0;
    uint32_t xph = 1, wph = 1; // ring "empty" barriers: the first pass over the ring does not block
    uint32_t bph = 0;          // empty_b: k-th reload waits for the k-th release by the consumers
    int prev_key = -1;
    for (long long ch = lo; ch < hi; ch++) {
      const int pair = (int)(ch / Tp);
      int lc = (int)(ch - (long long)pair * Tp);
      const int nt = pair / g.m_tiles, mt = pair - nt * g.m_tiles;
      const int qc = lc % g.QC;
      lc /= g.QC;
      const int sl = lc % g.S, pt = lc / g.S;
      const int key = pt * n_tiles + nt;
      if (key != prev_key) {
        if (prev_key >= 0) { // wait until all consumer warps have released the previous A_p tile
          mbar_wait(empty_b, bph);
          bph ^= 1;
        }
        mbar_expect_tx(full_b, Cfg::B_BYTES);
        tma_load_2d(Bs, tmB, full_b, pt * KT, nt * N_TILE);
        prev_key = key;
      }
      // w chunk
      mbar_wait(&empty_w[wsi], wph);
      mbar_expect_tx(&full_w[wsi], Cfg::W_BYTES);
      tma_load_2d(Ws + wsi * (N_TILE * OC), tmW, &full_w[wsi], qc * OC, nt * N_TILE);
      if (++wsi == 2) {
        wsi = 0;
        wph ^= 1;
      }
      // slow outer coordinates
      int lbase = 0, ubase = 0;
      {
        int rem = sl;
        for (int k = 1; k < g.n_outer; k++) {
          const int ik = rem % g.outer_dim[k];
          rem /= g.outer_dim[k];
          lbase += ik * g.outer_lmul[k];
          ubase += ik * g.outer_umul[k];
        }
      }
      const int nvalid = min(OC, Iq - qc * OC);
      for (int o = 0; o < nvalid; o++) {
        const int iq = qc * OC + o;
        mbar_wait(&empty_x[xs], xph);
        mbar_expect_tx(&full_x[xs], Cfg::X_STAGE_BYTES);
        tma_load_4d(Xs + xs * (M_TILE * KT), &maps.X, &full_x[xs], pt * KT, lbase + iq * g.outer_lmul[0],
                    mt * M_TILE, ubase + iq * g.outer_umul[0]);
        if (++xs == X_STAGES) {
          xs = 0;
          xph ^= 1;
        }
      }
    }
    return;
  }

  // ===================================== consumers =====================================
  const int r = lane >> 2, s = lane & 3;
  const int wn0 = warp * (8 * WN);
  double acc[WM][WN][2];
#pragma unroll
  for (int i = 0; i < WM; i++)
#pragma unroll
    for (int j = 0; j < WN; j++)
      acc[i][j][0] = acc[i][j][1] = 0.0;
  double wslow[WN];
#pragma unroll
  for (int j = 0; j < WN; j++)
    wslow[j] = 1.0;

  int xs = 0, wsi = 0;
  uint32_t xph = 0, wph = 0, bph = 0;
  int prev_key = -1, prev_pair = -1, prev_sl = -1, prev_nt = -1;

  auto flush = [&](int pair) {
    const int slot = pair * sk.kmax + ((int)blockIdx.x - sk.first_cta(pair));
    double *tile = ws + (size_t)slot * (M_TILE * N_TILE);
#pragma unroll
    for (int i = 0; i < WM; i++)
#pragma unroll
      for (int j = 0; j < WN; j++) {
        double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
        *reinterpret_cast<double2 *>(tile + (i * 8 + r) * N_TILE + wn0 + j * 8 + 2 * s) = v;
        acc[i][j][0] = acc[i][j][1] = 0.0;
      }
  };

  for (long long ch = lo; ch < hi; ch++) {
    const int pair = (int)(ch / Tp);
    int lc = (int)(ch - (long long)pair * Tp);
    const int nt = pair / g.m_tiles, mt = pair - nt * g.m_tiles;
    const int qc = lc % g.QC;
    lc /= g.QC;
    const int sl = lc % g.S, pt = lc / g.S;
    // m8 / n8 sub-tiles that lie completely outside G are skipped (CTA-uniform resp. warp-uniform predicates)
    const int wm_valid = min(WM, (g.In - mt * M_TILE + 7) >> 3);
    const int wn_valid = max(0, min(WN, (C - nt * N_TILE - wn0 + 7) >> 3));
    if (pair != prev_pair) {
      if (prev_pair >= 0)
        flush(prev_pair);
      prev_pair = pair;
    }
    const int key = pt * n_tiles + nt;
    if (key != prev_key) {
      if (prev_key >= 0) { // this warp is done with the previous A_p tile
        __syncwarp();
        if (lane == 0)
          mbar_arrive(empty_b);
      }
      mbar_wait(full_b, bph);
      bph ^= 1;
      prev_key = key;
    }
    if (g.n_outer > 1 && (sl != prev_sl || nt != prev_nt)) {
      // product of the slower outer factors for this thread's columns (rare: once per sweep of the fastest outer mode)
      int idx[CALS_MAX_OUTER];
      int rem = sl;
      for (int k = 1; k < g.n_outer; k++) {
        idx[k] = rem % g.outer_dim[k];
        rem /= g.outer_dim[k];
      }
#pragma unroll
      for (int j = 0; j < WN; j++) {
        const int c = nt * N_TILE + wn0 + j * 8 + r;
        double w = 1.0;
        if (c < C)
          for (int k = 1; k < g.n_outer; k++) {
            const int md = g.outer_mode[k];
            w *= __ldg(fac.buf[cur][md] + (size_t)c * g.ldF[md] + idx[k]);
          }
        wslow[j] = w;
      }
    }
    prev_sl = sl;
    prev_nt = nt;

    mbar_wait(&full_w[wsi], wph);
    const double *Wc = Ws + wsi * (N_TILE * OC);
    const int nvalid = min(OC, Iq - qc * OC);
    const int kvalid = min(KT, g.Ip - pt * KT);
    const int ngroups = (kvalid + 7) >> 3;
    for (int o = 0; o < nvalid; o++) {
      double wv[WN];
#pragma unroll
      for (int j = 0; j < WN; j++)
        wv[j] = Wc[(wn0 + j * 8 + r) * OC + o] * wslow[j];
      mbar_wait(&full_x[xs], xph);
      const double *Xt = Xs + xs * (M_TILE * KT);
#pragma unroll
      for (int gk = 0; gk < KT / 8; gk++) {
        if (gk < ngroups) { // warp-uniform: rows of the K tile beyond Ip are zero-filled by TMA, skip them
          double2 a[WM], b[WN];
#pragma unroll
          for (int i = 0; i < WM; i++)
            a[i] = *reinterpret_cast<const double2 *>(Xt + (i * 8 + r) * KT + gk * 8 + 2 * s);
#pragma unroll
          for (int j = 0; j < WN; j++) {
            b[j] = *reinterpret_cast<const double2 *>(Bs + (wn0 + j * 8 + r) * KT + gk * 8 + 2 * s);
            b[j].x *= wv[j];
            b[j].y *= wv[j];
          }
          // k = 8*gk + 2s in the first pass, 8*gk + 2s + 1 in the second: the two updates of one accumulator are
          // WM*WN instructions apart (DMMA latency is ~26 cycles)
#pragma unroll
          for (int i = 0; i < WM; i++)
            if (i < wm_valid)
#pragma unroll
              for (int j = 0; j < WN; j++)
                if (j < wn_valid)
                  dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
          for (int i = 0; i < WM; i++)
            if (i < wm_valid)
#pragma unroll
              for (int j = 0; j < WN; j++)
                if (j < wn_valid)
                  dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
        }
      }
      __syncwarp();
      if (lane == 0)
        mbar_arrive(&empty_x[xs]);
      if (++xs == X_STAGES) {
        xs = 0;
        xph ^= 1;
      }
    }
    __syncwarp();
    if (lane == 0)
      mbar_arrive(&empty_w[wsi]);
    if (++wsi == 2) {
      wsi = 0;
      wph ^= 1;
    }
  }
  flush(prev_pair);
}

// ------------------------------------------------------------------------------------------------------------------
// Sum the stream-K partial tiles of every (m,n) pair in CTA order and write G (In x C, column-major, ld = ldG).
// One CTA per 32x32 patch of G; reads are coalesced along columns of the row-major tiles, writes along rows of G.
template <int M_TILE, int N_TILE>
__global__ void mttkrp_reduce_kernel(const MttkrpGeom g, const SchedState *__restrict__ st,
                                     const double *__restrict__ ws, double *__restrict__ G, int grid_ctas,
                                     int C_override) {
  const int C = C_override > 0 ? C_override : st->C;
  const int c0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  if (c0 >= C)
    return;
  __shared__ double t[32][33];
  const int n_tiles = (C + N_TILE - 1) / N_TILE;
  const int Tp = g.P_tiles * g.S * g.QC;
  const StreamK sk = StreamK::make(g.m_tiles * n_tiles, Tp, grid_ctas);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 256 threads: ty in 0..7
  for (int rr = ty; rr < 32; rr += 8) {
    const int m = m0 + rr, c = c0 + tx;
    double sum = 0.0;
    if (m < g.In && c < C) {
      const int mt = m / M_TILE, nt = c / N_TILE;
      const int pair = nt * g.m_tiles + mt;
      const int nslots = sk.last_cta(pair) - sk.first_cta(pair) + 1;
      const double *p = ws + (size_t)pair * sk.kmax * (M_TILE * N_TILE) + (m - mt * M_TILE) * N_TILE + (c - nt * N_TILE);
      for (int k = 0; k < nslots; k++)
        sum += p[(size_t)k * (M_TILE * N_TILE)];
    }
    t[rr][tx] = sum;
  }
  __syncthreads();
  for (int cc = ty; cc < 32; cc += 8) {
    const int m = m0 + tx, c = c0 + cc;
    if (m < g.In && c < C)
      G[(size_t)c * g.ldG + m] = t[tx][cc];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Cross-check kernel: one thread per output element, plain loops over the ORIGINAL tensor layout.  Not a product
// path (selected only through the CALS_B200_MTTKRP_NAIVE test hook).
struct NaiveGeom {
  int n_modes, mode;
  int dims[CALS_MAX_MODES];
  int ldF[CALS_MAX_MODES];
  long long xstride[CALS_MAX_MODES]; // element strides of the padded device copy of X
  int ldG;
};

__global__ void mttkrp_naive_kernel(const NaiveGeom g, const SchedState *__restrict__ st, const double *__restrict__ X,
                                    const FactorPtrs fac, double *__restrict__ G, int C_override) {
  const int C = C_override > 0 ? C_override : st->C;
  const int cur = C_override > 0 ? 0 : st->cur;
  const int In = g.dims[g.mode];
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (long long)In * C)
    return;
  const int i = (int)(tid % In), c = (int)(tid / In);
  int idx[CALS_MAX_MODES];
  for (int k = 0; k < g.n_modes; k++)
    idx[k] = 0;
  idx[g.mode] = i;
  double sum = 0.0;
  while (true) {
    long long off = 0;
    double w = 1.0;
    for (int k = 0; k < g.n_modes; k++) {
      off += idx[k] * g.xstride[k];
      if (k != g.mode)
        w *= fac.buf[cur][k][(size_t)c * g.ldF[k] + idx[k]];
    }
    sum += X[off] * w;
    int k = 0;
    for (; k < g.n_modes; k++) {
      if (k == g.mode)
        continue;
      if (++idx[k] < g.dims[k])
        break;
      idx[k] = 0;
    }
    if (k == g.n_modes)
      break;
  }
  G[(size_t)c * g.ldG + i] = sum;
}

} // namespace calsb200
