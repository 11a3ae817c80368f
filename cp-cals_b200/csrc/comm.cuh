// Cross-GPU combination of partial MTTKRPs for a tensor sliced along one mode (BASELINE config 5):
// every GPU holds a slab of X and computes, for every mode n,
//     n != sliced mode : a full-size PARTIAL sum  G_r (I_n x C)           ->  G = sum_r G_r      (all-reduce)
//     n == sliced mode : its own row block of G                           ->  G = rows of G_r    (all-gather)
// The reference has no counterpart (single device); north_star asks for NCCL or an own kernel over NVLink here.
//
// This is an own kernel over NVLink peer memory, fused with the barrier it needs: one launch per mode and GPU
//   1. publishes "my partial for exchange #seq is in my exchange buffer" by storing seq into a flag word in every
//      peer's memory (st.release.sys over NVLink),
//   2. waits until the flags of all peers in its own memory have reached seq (ld.acquire.sys spin, bounded),
//   3. reads the partials of ALL ranks straight from their HBM (coalesced 16-byte ld.global on peer-mapped pointers)
//      and adds them in rank order -- the same order on every GPU, so the replicated factor state stays bit-identical
//      across GPUs without any further synchronisation -- writing G locally.
// Exchange buffers are double-buffered by the parity of seq; together with the per-exchange barrier this makes the
// buffer a rank overwrites in exchange k+1 one that no peer can still be reading (they read it in exchange k-1 and
// have signalled exchange k since).
//
// Volume per GPU and mode: (W-1) * 8 * I_n * C bytes pulled over NVLink (config 5, W = 8: 71 MB ~ 0.1 ms at the measured
// 770 GB/s) against ~10 ms of DMMA work for the slab -- the exchange is not worth overlapping tile by tile.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace calsb200 {

constexpr int CALS_MAX_PEERS = 16;
constexpr int COMM_FLAG_BYTES = 256; // flag area at the start of every exchange block (CALS_MAX_PEERS x u64, padded)

struct CommParams {
  int rank, world;
  int slice_mode;
  int cuts[CALS_MAX_PEERS + 1];                     // slab boundaries along the sliced mode
  const double *peer_x[CALS_MAX_PEERS];             // every rank's exchange buffers: [2][cap] doubles
  unsigned long long *peer_flags[CALS_MAX_PEERS];   // every rank's flag words: [world]
  unsigned long long cap;                           // doubles per exchange buffer
  unsigned long long seq_base;                      // see mttkrp_reduce_kernel
  long long spin_limit;                             // clock64 ticks before a missing peer is declared lost
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// grid: any (grid-stride); block: 256.  rows/ld: extent and pitch of mode `mode`'s factor (full extent).
__global__ void __launch_bounds__(256)
exchange_sum_kernel(const CommParams cp, SchedState *st, double *__restrict__ G, int mode, int n_modes, int rows,
                    int ld) {
  const int C = st->C;
  if (C <= 0)
    return;
  const unsigned long long seq = cp.seq_base + st->global_iter * (unsigned long long)n_modes + mode + 1;

  if (blockIdx.x == 0 && threadIdx.x < cp.world && (int)threadIdx.x != cp.rank) {
    __threadfence_system();
    st_release_sys(cp.peer_flags[threadIdx.x] + cp.rank, seq);
  }
  __shared__ int lost;
  if (threadIdx.x == 0)
    lost = 0;
  __syncthreads();
  if ((int)threadIdx.x < cp.world && (int)threadIdx.x != cp.rank && !st->comm_error) {
    const unsigned long long *f = cp.peer_flags[cp.rank] + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < seq) {
      if (clock64() - t0 > cp.spin_limit) {
        lost = 1;
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (lost) { // a peer never arrived: flag the run as failed (the host reports it) and stop waiting from now on
    if (threadIdx.x == 0)
      st->comm_error = 1;
    return;
  }

  const unsigned long long boff = (seq & 1ull) * cp.cap;
  const int ld2 = ld >> 1; // ld is even; rows are handled in pairs
  const long long total = (long long)ld2 * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int m2 = (int)(e % ld2);
    const long long c = e / ld2;
    const int m = 2 * m2;
    if (m >= rows)
      continue;
    const size_t off = (size_t)c * ld + m;
    double2 acc;
    if (mode == cp.slice_mode) {
      int owner = 0;
      while (owner + 1 < cp.world && m >= cp.cuts[owner + 1]) // cuts are even, so a row pair never straddles two slabs
        owner++;
      acc = *reinterpret_cast<const double2 *>(cp.peer_x[owner] + boff + off);
    } else {
      acc = *reinterpret_cast<const double2 *>(cp.peer_x[0] + boff + off);
      for (int r = 1; r < cp.world; r++) {
        const double2 v = *reinterpret_cast<const double2 *>(cp.peer_x[r] + boff + off);
        acc.x += v.x;
        acc.y += v.y;
      }
    }
    *reinterpret_cast<double2 *>(G + off) = acc;
  }
}

} // namespace calsb200
