// Device-side model queue: the MultiKtensor bookkeeping of the reference without host round trips.
//
//   admission  : FIFO, stop at the first model that does not fit           reference src/cals.cpp:182-192,
//                (first-fit == append, because the buffer is compacted)    src/multi_ktensor.cpp:14-39,41-130
//   eviction   : copy the model's columns back to its own ("home") storage reference src/multi_ktensor.cpp:132-163,
//                                                                           src/ktensor.cpp:127-135
//   compaction : stable shift of the live models to the left               reference src/multi_ktensor.cpp:188-264
//
// Compaction is a gather from the old buffer into the other one of a ping-pong pair (in-place overlapping shifts would
// need a grid-wide ordering); newly admitted models are gathered from their home columns in the same pass.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"
#include "mttkrp.cuh"

namespace calsb200 {

struct SchedParams {
  SchedState *st;
  ModelDesc *models;
  int *live;       // [max_live] model ids in column order (current)
  int *live_tmp;   // scratch
  int *gather_src; // [buffer_cols] new column j <- (>=0: old buffer column, <0: home column -1-src)
  int *evict_dst;  // [buffer_cols] old column j -> home column, or -1
  int *host_flags; // mapped pinned host memory: [0] done, [1] global_iter (low 31 bits), [2] n_live
  PlanArgs plans;  // MTTKRP work partition per mode, rebuilt whenever the live column count changes
  unsigned *iter_cols; // [ITER_LOG_CAP] active column count of every executed iteration (CalsReport::cols)
};
constexpr int ITER_LOG_CAP = 65536;

// Thread 0 does the queue bookkeeping (the live list is short, <= buffer_cols entries, and this runs once per CALS
// iteration); then one thread per mode rebuilds that mode's MTTKRP plan if the column count changed.
__device__ void sched_serial(const SchedParams &p);

__global__ void sched_kernel(const SchedParams p) {
  __shared__ int replan;
  if (threadIdx.x == 0) {
    const int C_before = p.st->C;
    sched_serial(p);
    replan = (p.st->C != C_before && p.st->C > 0) ? 1 : 0;
  }
  __syncthreads();
  if (replan && (int)threadIdx.x < p.plans.n_modes)
    mttkrp_make_plan(p.plans.plan[threadIdx.x], p.plans.shape[threadIdx.x], p.st->C, p.plans.G);
}

__device__ void sched_serial(const SchedParams &p) {
  SchedState *st = p.st;
  if (st->done) {
    st->changed = 0;
    return;
  }
  const int B = st->buffer_cols;
  bool changed = false;

  // always_evict_first: evict the leftmost live model, whatever its state (reference src/cals.cpp:348-354)
  if ((st->flags & 2u) && st->global_iter > 0 && st->n_live > 0) {
    p.models[p.live[0]].state = MODEL_EVICT;
    st->n_evict = 1;
  }

  int n_live = st->n_live, col = 0, kept = 0;
  if (st->n_evict > 0) {
    for (int j = 0; j < B; j++)
      p.evict_dst[j] = -1;
    for (int li = 0; li < n_live; li++) {
      const int m = p.live[li];
      ModelDesc &md = p.models[m];
      if (md.state == MODEL_EVICT) {
        for (int j = 0; j < md.rank; j++)
          p.evict_dst[md.col + j] = md.home_col + j;
        md.state = MODEL_DONE;
      } else {
        for (int j = 0; j < md.rank; j++)
          p.gather_src[col + j] = md.col + j;
        md.col = col;
        col += md.rank;
        p.live_tmp[kept++] = m;
      }
    }
    for (int li = 0; li < kept; li++)
      p.live[li] = p.live_tmp[li];
    n_live = kept;
    changed = true;
    st->n_evict = 0;
  } else {
    col = st->C;
  }

  // admission
  bool admitted = false;
  while (st->next < st->n_models) {
    ModelDesc &md = p.models[st->next];
    if (col + md.rank > B)
      break;
    if (!changed && !admitted) {
      // first structural change of this iteration without evictions: identity map for the resident columns
      for (int j = 0; j < B; j++)
        p.evict_dst[j] = -1;
      for (int j = 0; j < col; j++)
        p.gather_src[j] = j;
    }
    admitted = true;
    for (int j = 0; j < md.rank; j++)
      p.gather_src[col + j] = -1 - (md.home_col + j);
    md.col = col;
    md.state = MODEL_LIVE;
    md.iters = 1; // reference src/multi_ktensor.cpp:96
    col += md.rank;
    p.live[n_live++] = st->next;
    st->n_admitted += 1;
    st->comp_sum += md.rank;
    st->next += 1;
  }
  changed = changed || admitted;

  st->n_live = n_live;
  st->C = col;
  st->changed = changed ? 1 : 0;
  if (changed)
    st->cur ^= 1;
  if (n_live == 0 && st->next >= st->n_models) {
    st->done = 1;
  } else {
    if (st->global_iter < (unsigned long long)ITER_LOG_CAP)
      p.iter_cols[st->global_iter] = (unsigned)col;
    st->global_iter += 1; // reference: rep.iter counts executed loop bodies (src/cals.cpp:175-176)
    st->col_iter_sum += (unsigned long long)col;
  }
  p.host_flags[1] = (int)(st->global_iter & 0x7fffffff);
  p.host_flags[2] = n_live;
  __threadfence_system();
  p.host_flags[0] = st->done;
}

// Column mover.  grid = (ceil(buffer_cols / COLS_PER_CTA), n_modes, 2): z == 0 copies evicted columns old buffer ->
// home, z == 1 gathers the new buffer from (old buffer | home).
constexpr int MOVE_COLS = 4;
__global__ void __launch_bounds__(256)
move_kernel(const Geom geo, const FactorPtrs fac, const SchedState *__restrict__ st,
            const int *__restrict__ gather_src, const int *__restrict__ evict_dst) {
  if (!st->changed)
    return;
  const int n = blockIdx.y;
  const int rows = geo.dims[n], ld = geo.ldF[n];
  const int cur = st->cur; // already flipped by the scheduler: cur is the NEW buffer
  const double *oldb = fac.buf[cur ^ 1][n];
  double *newb = fac.buf[cur][n];
  double *home = fac.home[n];
  const int j0 = blockIdx.x * MOVE_COLS;
  if (blockIdx.z == 0) {
    for (int jj = 0; jj < MOVE_COLS; jj++) {
      const int j = j0 + jj;
      if (j >= st->buffer_cols)
        break;
      const int dst = evict_dst[j];
      if (dst < 0)
        continue;
      for (int r = threadIdx.x; r < rows; r += blockDim.x)
        home[(size_t)dst * ld + r] = oldb[(size_t)j * ld + r];
    }
  } else {
    for (int jj = 0; jj < MOVE_COLS; jj++) {
      const int j = j0 + jj;
      if (j >= st->C)
        break;
      const int src = gather_src[j];
      const double *s = src >= 0 ? oldb + (size_t)src * ld : home + (size_t)(-1 - src) * ld;
      for (int r = threadIdx.x; r < rows; r += blockDim.x)
        newb[(size_t)j * ld + r] = s[r];
    }
  }
}

} // namespace calsb200
