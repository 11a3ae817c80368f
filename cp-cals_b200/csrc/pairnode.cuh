// Pair nodes: a one-level dimension tree over the modes.  Two modes take their MTTKRPs from ONE tensor-sized
// contraction T that does not involve either of their factors:
//
//   3 modes (pair = modes 1, 2):   T[(i1,i2), c] = sum_{i0} X[i0,i1,i2] * A_0[i0,c]           pair_gemm_kernel (FP64 tensor cores)
//   4 modes (pairs (0,1), (2,3)):  T0[(i0,i1), c] = sum_{i2,i3} X * A_2[i2,c] * A_3[i3,c]      mttkrp_dmma_kernel on a 3-way view
//                                  T1[(i2,i3), c] = sum_{i0,i1} X * A_0[i0,c] * A_1[i1,c]      (set up in engine.cu)
//   leaves, for a pair (f, s) with T rows (i_f, i_s), i_f fastest:
//                                  G_f[i_f, c] = sum_{i_s} T[(i_f,i_s), c] * A_s[i_s, c]        pair_leaf_slow_kernel (HBM-bound)
//                                  G_s[i_s, c] = sum_{i_f} T[(i_f,i_s), c] * A_f[i_f, c]        pair_leaf_fast_kernel (HBM-bound)
//
// ALS updates mode f, then mode s, and T holds neither factor, so both leaves see exactly the factors the reference's
// per-mode MTTKRPs see (src/cals.cpp:214-222 calls mttkrp::mttkrp once per mode): the sums are the same, only their
// association differs.  An ALS iteration then costs two tensor-sized contractions instead of three (four).  This is the
// reference's own two-step idea (mttkrp_twostep0/1, src/utils/mttkrp.cpp:460-553: "mode1/TS0 uses the same T as mode0/TS0")
// carried across modes; T lives in HBM.  CalsParams::mttkrp_method == MTTKRP selects one full MTTKRP per mode instead.
#pragma once
#include "mttkrp.cuh"

namespace calsb200 {

constexpr int PAIR_STAGES = 2;

template <int WM, int WN> struct PairCfg {
  static constexpr int M_TILE = 8 * WM;
  static constexpr int N_TILE = 8 * WN * NUM_MMA_WARPS;
  static constexpr int X_BYTES = M_TILE * KT * 8;
  static constexpr int B_BYTES = N_TILE * KT * 8;
  static constexpr int STAGE_BYTES = X_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = PAIR_STAGES * STAGE_BYTES + 256;
};

struct PairMaps {
  CUtensorMap X;    // 2-D view (I0, I1*I2) of the tensor in the caller's order: box [KT x M_TILE]
  CUtensorMap B[2]; // factor of mode 0 in buffer 0 / 1: box [KT x N_TILE]
};

struct PairGeom {
  int R;   // rows of T = I1 * I2
  int Ip;  // contracted extent (I0)
  int E1;  // extent of the fast index of a T row (I1)
  int E2;  // extent of the slow index (I2)
  int mode_fast, mode_slow; // the two modes of the pair (1 and 2)
  // Slab mode (3-mode tensor sliced along mode 1 or 2 over several GPUs): this device's T covers rows
  // [off, off + E) of the sliced mode's factor and of its G; 0 otherwise.
  int off_fast, off_slow;
  long long ldT;
  int ldF[CALS_MAX_MODES]; // leading dimensions of the factor buffers == those of G for the same mode
};

// T = X_(0)^T A_0.  Persistent CTAs over the (m-tile, n-tile) grid, m-tile fastest, so that the CTAs running at the
// same time share the factor tiles in L2 and the cheap tiles of the ragged last n-tile come last.  Both operands of a
// K tile travel together through a two-stage TMA ring; the warp layout and the inner loop are those of
// mttkrp_dmma_kernel (mma_stage) without the outer weight.
template <int WM, int WN>
__global__ void __launch_bounds__(MTTKRP_THREADS, 1)
pair_gemm_kernel(const __grid_constant__ PairMaps maps, const PairGeom g, const SchedState *__restrict__ st,
                 double *__restrict__ T) {
  using Cfg = PairCfg<WM, WN>;
  constexpr int M_TILE = Cfg::M_TILE, N_TILE = Cfg::N_TILE;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t *bars = (uint64_t *)(smem + PAIR_STAGES * Cfg::STAGE_BYTES);
  uint64_t *full = bars, *empty = bars + PAIR_STAGES;

  const int C = st->C, cur = st->cur;
  if (C <= 0)
    return;
  const int m_tiles = (g.R + M_TILE - 1) / M_TILE;
  const int n_tiles = (C + N_TILE - 1) / N_TILE;
  const long long tiles = (long long)m_tiles * n_tiles;
  const int P_tiles = (g.Ip + KT - 1) / KT;
  if ((long long)blockIdx.x >= tiles)
    return;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < PAIR_STAGES; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NUM_MMA_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= NUM_MMA_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp != NUM_MMA_WARPS || lane != 0)
      return;
    const CUtensorMap *tmB = &maps.B[cur];
    int sidx = 0;
    uint32_t ph = 1;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int nt = (int)(t / m_tiles), mt = (int)(t - (long long)nt * m_tiles);
      for (int pt = 0; pt < P_tiles; pt++) {
        unsigned char *stage = smem + sidx * Cfg::STAGE_BYTES;
        mbar_wait(&empty[sidx], ph);
        mbar_expect_tx(&full[sidx], Cfg::STAGE_BYTES);
        tma_load_2d(stage, &maps.X, &full[sidx], pt * KT, mt * M_TILE);
        tma_load_2d(stage + Cfg::X_BYTES, tmB, &full[sidx], pt * KT, nt * N_TILE);
        if (++sidx == PAIR_STAGES) {
          sidx = 0;
          ph ^= 1;
        }
      }
    }
    return;
  }

  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int r = lane >> 2, s = lane & 3;
  double acc[WM][WN][2];
  const double ones[WN] = {1.0, 1.0, 1.0, 1.0};
  int sidx = 0;
  uint32_t ph = 0;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int nt = (int)(t / m_tiles), mt = (int)(t - (long long)nt * m_tiles);
    const int m0 = mt * M_TILE, c0 = nt * N_TILE;
    const int nm = min(WM, (g.R - m0 + 7) >> 3);
    int nn = min(WN, (C - c0 + 63) >> 6);
    if (c0 + 64 * (nn - 1) + 8 * warp >= C) // this warp's n8 group of the last octet starts beyond C (see mttkrp.cuh)
      nn--;
#pragma unroll
    for (int i = 0; i < WM; i++)
#pragma unroll
      for (int j = 0; j < WN; j++)
        acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int pt = 0; pt < P_tiles; pt++) {
      const double *Xs = (const double *)(smem + sidx * Cfg::STAGE_BYTES);
      const double *Bw = (const double *)(smem + sidx * Cfg::STAGE_BYTES + Cfg::X_BYTES) + warp * 8 * KT;
      const int kvalid = min(KT, g.Ip - pt * KT);
      mbar_wait(&full[sidx], ph);
      mma_stage<WM, WN, false>(acc, Xs, Bw, ones, (kvalid + 7) >> 3, r, s, nm, nn);
      __syncwarp();
      if (lane == 0)
        mbar_arrive(&empty[sidx]);
      if (++sidx == PAIR_STAGES) {
        sidx = 0;
        ph ^= 1;
      }
    }
    // rows m0 + 8i + r, columns c0 + 64j + 8*warp + 2s (+1): for one register the 8 lanes of equal s write 8 consecutive
    // rows of one column (64 B runs)
#pragma unroll
    for (int i = 0; i < WM; i++) {
      const int row = m0 + i * 8 + r;
      if (row < g.R) {
#pragma unroll
        for (int j = 0; j < WN; j++) {
          const int col = c0 + j * 64 + warp * 8 + 2 * s;
          if (col < C)
            T[(size_t)col * g.ldT + row] = acc[i][j][0];
          if (col + 1 < C)
            T[(size_t)(col + 1) * g.ldT + row] = acc[i][j][1];
        }
      }
    }
  }
}

// Slab mode over several GPUs: a leaf's result is this device's PARTIAL sum (or its own row block) and goes to the exchange
// buffer that comm.cuh's exchange kernel collects, exactly as mttkrp_reduce_kernel does for a full MTTKRP.
struct LeafExchange {
  double *xbuf; // nullptr: write G
  unsigned long long xcap, seq_base;
  int n_modes;
};
__device__ __forceinline__ double *leaf_output(double *G, const LeafExchange &x, const SchedState *st, int mode) {
  if (x.xbuf == nullptr)
    return G;
  const unsigned long long seq = x.seq_base + st->global_iter * (unsigned long long)x.n_modes + mode + 1;
  return x.xbuf + (seq & 1ull) * x.xcap;
}

constexpr int LEAF_DEPTH = 16; // independent loads of T in flight per thread
// Scheduling fence: the values must all be in registers here, so every load above has been issued before anything
// below runs (the hardware issues in order; an FMA waiting for its operand would hold up the loads behind it).
__device__ __forceinline__ void all_loaded(double (&v)[LEAF_DEPTH]) {
  asm volatile("" ::"d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "d"(v[4]), "d"(v[5]), "d"(v[6]), "d"(v[7]), "d"(v[8]),
               "d"(v[9]), "d"(v[10]), "d"(v[11]), "d"(v[12]), "d"(v[13]), "d"(v[14]), "d"(v[15]));
}

// G_fast[i1, c] = sum_{i2} T[i1 + E1*i2, c] * A_slow[i2, c].  One CTA per column and 256-row chunk of i1; thread = row.
__global__ void __launch_bounds__(256)
pair_leaf_slow_kernel(const PairGeom g, const SchedState *__restrict__ st, const FactorPtrs fac,
                      const double *__restrict__ T, double *__restrict__ G, const LeafExchange x) {
  const int C = st->C, cur = st->cur;
  const int c = blockIdx.x;
  if (c >= C)
    return;
  G = leaf_output(G, x, st, g.mode_fast);
  extern __shared__ double wcol[]; // E2
  const double *W = fac.buf[cur][g.mode_slow] + (size_t)c * g.ldF[g.mode_slow] + g.off_slow;
  for (int k = threadIdx.x; k < g.E2; k += blockDim.x)
    wcol[k] = W[k];
  __syncthreads();
  const int i1 = blockIdx.y * blockDim.x + threadIdx.x;
  if (i1 >= g.E1)
    return;
  const double *t = T + (size_t)c * g.ldT + i1;
  double sum = 0.0;
  int k = 0;
  // LEAF_DEPTH loads are issued before the first one is consumed (all_loaded keeps the compiler from interleaving the
  // FMAs with the loads, which it otherwise does at 4 loads in flight): with one CTA per column, a grid of a few hundred
  // columns -- the per-GPU shard of a strong-scaled run -- streams T at (bytes in flight per CTA) / latency.
  for (; k + LEAF_DEPTH <= g.E2; k += LEAF_DEPTH) {
    double v[LEAF_DEPTH];
#pragma unroll
    for (int u = 0; u < LEAF_DEPTH; u++)
      v[u] = __ldcs(t + (size_t)(k + u) * g.E1);
    all_loaded(v);
#pragma unroll
    for (int u = 0; u < LEAF_DEPTH; u++)
      sum += v[u] * wcol[k + u];
  }
  for (; k < g.E2; k++)
    sum += __ldcs(t + (size_t)k * g.E1) * wcol[k];
  G[(size_t)c * g.ldF[g.mode_fast] + i1 + g.off_fast] = sum;
}

// G_slow[i2, c] = sum_{i1} T[i1 + E1*i2, c] * A_fast[i1, c].  One CTA per column; a warp takes LEAF_ROWS values of i2 at a
// time (independent loads in flight), lanes run over i1.
constexpr int LEAF_ROWS = 4;
static_assert(4 * LEAF_ROWS == LEAF_DEPTH, "the fast leaf batches four lane-strides of LEAF_ROWS rows");
__global__ void __launch_bounds__(256)
pair_leaf_fast_kernel(const PairGeom g, const SchedState *__restrict__ st, const FactorPtrs fac,
                      const double *__restrict__ T, double *__restrict__ G, const LeafExchange x) {
  const int C = st->C, cur = st->cur;
  const int c = blockIdx.x;
  if (c >= C)
    return;
  G = leaf_output(G, x, st, g.mode_slow);
  extern __shared__ double wcol[]; // E1
  const double *W = fac.buf[cur][g.mode_fast] + (size_t)c * g.ldF[g.mode_fast] + g.off_fast;
  for (int k = threadIdx.x; k < g.E1; k += blockDim.x)
    wcol[k] = W[k];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const double *tc = T + (size_t)c * g.ldT;
  double *out = G + (size_t)c * g.ldF[g.mode_slow] + g.off_slow;
  for (int b2 = (blockIdx.y * nw + warp) * LEAF_ROWS; b2 < g.E2; b2 += gridDim.y * nw * LEAF_ROWS) {
    const double *t[LEAF_ROWS];
    double sum[LEAF_ROWS];
#pragma unroll
    for (int u = 0; u < LEAF_ROWS; u++) {
      t[u] = tc + (size_t)min(b2 + u, g.E2 - 1) * g.E1; // rows beyond E2 re-read the last row; their sums are dropped
      sum[u] = 0.0;
    }
    int i1 = lane;
    for (; i1 + 96 < g.E1; i1 += 128) { // four lane-strides of i1 x LEAF_ROWS rows = 16 loads in flight per thread
      double v[LEAF_DEPTH];
#pragma unroll
      for (int q = 0; q < 4; q++)
#pragma unroll
        for (int u = 0; u < LEAF_ROWS; u++)
          v[q * LEAF_ROWS + u] = __ldcs(t[u] + i1 + 32 * q);
      all_loaded(v);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const double w = wcol[i1 + 32 * q];
#pragma unroll
        for (int u = 0; u < LEAF_ROWS; u++)
          sum[u] += v[q * LEAF_ROWS + u] * w;
      }
    }
    for (; i1 < g.E1; i1 += 32) {
      const double w = wcol[i1];
      double v[LEAF_ROWS];
#pragma unroll
      for (int u = 0; u < LEAF_ROWS; u++)
        v[u] = __ldcs(t[u] + i1);
#pragma unroll
      for (int u = 0; u < LEAF_ROWS; u++)
        sum[u] += v[u] * w;
    }
#pragma unroll
    for (int u = 0; u < LEAF_ROWS; u++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        sum[u] += __shfl_xor_sync(0xffffffffu, sum[u], o);
      if (lane == 0 && b2 + u < g.E2)
        out[b2 + u] = sum[u];
    }
  }
}

} // namespace calsb200
