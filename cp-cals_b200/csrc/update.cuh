// Per-model small operations, one CTA per live model, fused into a single kernel per mode:
//
//   H = hadamard of the other modes' Gramians     ops::hadamard_but_one      reference src/utils/utils.cpp:161-172
//   L = chol_lower(H)                               dpotrf('L')                reference src/utils/update.cpp:183-185
//   F = G * L^-T * L^-1 (row-wise two solves)       2 x dtrsm                  reference src/utils/update.cpp:187-190
//   jackknife row := 0                              Ktensor::set_jk_fiber(0)   reference include/ktensor.h:316-325
//   lambda, column scaling                          Ktensor::normalize(n, it)  reference src/ktensor.cpp:66-83
//   Gram_n = F^T F                                  ops::update_gramian        reference src/utils/utils.cpp:174-178
// and, after the last mode,
//   P = hadamard of all Gramians, fast error, fit   ops::hadamard_all, error::compute_fast_error,
//                                                   Ktensor::calculate_new_fit reference src/utils/utils.cpp:156,
//                                                   src/utils/error.cpp:64-89, include/ktensor.h:178-183
//   eviction predicate / iters++                    reference src/cals.cpp:336-347
//
// The MTTKRP result G is read from its own buffer and the updated factor is written to the multi-factor buffer, so the
// last mode's G is still intact for the error term (the reference keeps a copy: G_last, src/cals.cpp:75,230-234).
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"
#include "mttkrp.cuh"

namespace calsb200 {

constexpr int UPDATE_THREADS = 256;

struct UpdateParams {
  int mode;
  int n_modes;
  int rows;      // I_n
  int ld;        // ldF[n] == ldG
  int chunk_rows; // rows staged in shared memory at a time (multiple of 32)
  int chunk_pitch; // odd pitch (in doubles) of the staged chunk
  int max_rank;
  const double *G;         // MTTKRP result, ld x C
  double *F[2];            // multi-factor buffers for this mode
  double *gram_pool;
  double *lambda_home;     // per home column
  const double *x_norms_jk; // modes[0] entries or nullptr
  ModelDesc *models;
  const int *live;
  SchedState *st;
  // fused reduction of the stream-K partial tiles (replaces mttkrp_reduce_kernel on the iteration path): when
  // plan != nullptr the MTTKRP result is not read from G but summed here, per (m,n) pair in segment order, from the
  // partial-tile workspace the DMMA kernel has just written; for the last mode the sum is also written to G (the
  // error term needs it after the solve).
  const int *plan;
  const double *ws;
  int plan_ctas;  // gridDim.x of the DMMA kernel (layout of the plan tables)
  int tile_elems; // M_TILE * N_TILE of this mode's DMMA kernel
  int n_tile;     // N_TILE
  double *G_out;  // == G (mutable alias)
  // non-negative update (update_method == NNLS)
  int nnls;                 // 0: Cholesky solve, 1: row-wise active-set NNLS
  int nnls_warps;           // warps of the CTA that work on rows (each owns a scratch block in shared memory)
  int rows_before;          // sum of the extents of the modes below this one (offset into a model's active sets)
  unsigned char *act_pool;  // per model: for every mode an I_n x R block, row-major, 1 = constrained to zero
};

__device__ __forceinline__ double block_sum(double v, double *scratch) {
  // scratch: >= 32 doubles of shared memory
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0)
    scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < (blockDim.x >> 5) ? scratch[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0)
      scratch[0] = t;
  }
  __syncthreads();
  t = scratch[0];
  __syncthreads();
  return t;
}


// ------------------------------------------------------------------------------------------------------------------
// Row-wise active-set NNLS: update::update_factor_non_negative_constrained (reference src/utils/update.cpp:61-176),
// one warp per factor row, everything in shared memory.  Scratch per warp (doubles): y d w s sp [5R], Gp [R*R];
// ints: act [R], map [R+1].
struct NnlsScratch {
  double *y, *d, *w, *s, *sp, *Gp;
  int *act, *map;
};

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Solve H_pp x = y_p for the passive set (dposv('L'), reference src/utils/update.cpp:18-46).  Returns the size of the
// passive set, or -1 when the Cholesky meets a non-positive pivot (CholFail).
__device__ int nnls_solve_passive(const double *H, int R, const NnlsScratch &q, int lane) {
  __syncwarp();
  if (lane == 0) {
    int p = 0;
    for (int i = 0; i < R; i++)
      if (!q.act[i])
        q.map[p++] = i;
    q.map[R] = p;
  }
  __syncwarp();
  const int p = q.map[R];
  double *Gp = q.Gp, *sp = q.sp;
  for (int e = lane; e < p * p; e += 32) {
    const int a = e % p, b = e / p;
    Gp[a + b * p] = H[q.map[a] + q.map[b] * R];
  }
  for (int a = lane; a < p; a += 32)
    sp[a] = q.y[q.map[a]];
  __syncwarp();
  for (int j = 0; j < p; j++) {
    const double dj = Gp[j + j * p];
    if (!(dj > 0.0))
      return -1; // warp-uniform: every lane read the same value
    const double sd = sqrt(dj);
    __syncwarp();
    if (lane == 0)
      Gp[j + j * p] = sd;
    for (int i = j + 1 + lane; i < p; i += 32)
      Gp[i + j * p] /= sd;
    __syncwarp();
    const int t = p - j - 1;
    for (int e = lane; e < t * t; e += 32) {
      const int i = j + 1 + e % t, k = j + 1 + e / t;
      if (k <= i)
        Gp[i + k * p] -= Gp[i + j * p] * Gp[k + j * p];
    }
    __syncwarp();
  }
  for (int j = 0; j < p; j++) { // L z = y_p
    const double z = sp[j] / Gp[j + j * p];
    __syncwarp();
    if (lane == 0)
      sp[j] = z;
    for (int i = j + 1 + lane; i < p; i += 32)
      sp[i] -= Gp[i + j * p] * z;
    __syncwarp();
  }
  for (int j = p - 1; j >= 0; j--) { // L^T x = z
    const double x = sp[j] / Gp[j + j * p];
    __syncwarp();
    if (lane == 0)
      sp[j] = x;
    for (int i = lane; i < j; i += 32)
      sp[i] -= Gp[j + i * p] * x;
    __syncwarp();
  }
  return p;
}

__device__ __forceinline__ void nnls_scatter(double *dst, int R, int p, const NnlsScratch &q, int lane) {
  for (int i = lane; i < R; i += 32)
    dst[i] = 0.0;
  __syncwarp();
  for (int a = lane; a < p; a += 32)
    dst[q.map[a]] = q.sp[a];
  __syncwarp();
}
__device__ __forceinline__ double nnls_min_sp(int p, const NnlsScratch &q, int lane) {
  double m = 1.7976931348623157e308;
  for (int a = lane; a < p; a += 32)
    m = fmin(m, q.sp[a]);
  return warp_min(m);
}
__device__ __forceinline__ int nnls_count_passive(int R, const NnlsScratch &q, int lane) {
  int n = 0;
  for (int i = lane; i < R; i += 32)
    n += !q.act[i];
  return warp_sum_int(n);
}
__device__ __forceinline__ void nnls_multipliers(const double *H, int R, const NnlsScratch &q, int lane) {
  for (int i = lane; i < R; i += 32) { // w = y - H d   (reference src/utils/update.cpp:49-57)
    double t = 0.0;
    for (int j = 0; j < R; j++)
      t += H[i + j * R] * q.d[j];
    q.w[i] = q.y[i] - t;
  }
  __syncwarp();
}

// One row: y = the row of the MTTKRP result in S (pitch `pitch`), result written back to S, active set updated in
// `arow` (global memory).  Returns 1 when a passive-block Cholesky failed inside the main loop or an iteration cap was
// hit (the reference has no caps; they only guard the GPU against a non-terminating corner case).
__device__ int nnls_row(const double *H, int R, double tol, double *Srow, int pitch, unsigned char *arow,
                        const NnlsScratch &q, int lane) {
  int any = 0;
  for (int i = lane; i < R; i += 32) {
    const double yi = Srow[i * pitch];
    int a = arow[i];
    if (yi > 0.0)
      a = 0;
    q.y[i] = yi;
    q.d[i] = 0.0;
    q.act[i] = a;
    any |= !a;
  }
  any = __any_sync(0xffffffffu, any);
  __syncwarp();
  int p = 0;
  if (any) { // warm start from the previous active set (reference :88-117)
    p = nnls_solve_passive(H, R, q, lane);
    bool failed = p < 0;
    for (int guard = 0; !failed; guard++) {
      nnls_scatter(q.d, R, p, q, lane);
      if (!(nnls_min_sp(p, q, lane) <= tol))
        break;
      for (int i = lane; i < R; i += 32)
        if (q.d[i] <= tol) {
          q.d[i] = 0.0;
          q.act[i] = 1;
        }
      __syncwarp();
      if (nnls_count_passive(R, q, lane) == 0 || guard > 2 * R) {
        failed = true;
        break;
      }
      p = nnls_solve_passive(H, R, q, lane);
      failed = p < 0;
    }
    if (failed) {
      for (int i = lane; i < R; i += 32) {
        q.act[i] = 1;
        q.d[i] = 0.0;
      }
      __syncwarp();
    }
  }
  nnls_multipliers(H, R, q, lane);
  int trouble = 0;
  for (int outer = 0;; outer++) { // main loop (reference :122-170)
    double best = -1.7976931348623157e308;
    int bi = 0x7fffffff;
    for (int i = lane; i < R; i += 32)
      if (q.act[i] && q.w[i] > best) { // strictly greater: the first index of the maximum within this lane's sequence
        best = q.w[i];
        bi = i;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    if (bi == 0x7fffffff || !(best > tol))
      break;
    if (outer > 3 * R) {
      trouble = 1;
      break;
    }
    if (lane == 0)
      q.act[bi] = 0;
    p = nnls_solve_passive(H, R, q, lane);
    if (p < 0) {
      trouble = 1;
      break;
    }
    bool bad = false;
    for (int inner = 0; nnls_min_sp(p, q, lane) <= tol; inner++) { // inner loop (reference :131-155)
      nnls_scatter(q.s, R, p, q, lane);
      double alpha = 1.7976931348623157e308;
      for (int i = lane; i < R; i += 32)
        if (!q.act[i] && q.s[i] <= tol)
          alpha = fmin(alpha, q.d[i] / (q.d[i] - q.s[i]));
      alpha = warp_min(alpha);
      for (int i = lane; i < R; i += 32) {
        const double di = q.d[i] + alpha * (q.s[i] - q.d[i]);
        q.d[i] = di;
        if (fabs(di) < tol && !q.act[i]) {
          q.act[i] = 1;
          q.d[i] = 0.0;
        }
      }
      __syncwarp();
      if (nnls_count_passive(R, q, lane) == 0 || inner > 3 * R) {
        bad = true;
        break;
      }
      p = nnls_solve_passive(H, R, q, lane);
      if (p < 0) {
        bad = true;
        break;
      }
    }
    if (bad) {
      trouble = 1;
      break;
    }
    nnls_scatter(q.d, R, p, q, lane);
    nnls_multipliers(H, R, q, lane);
  }
  for (int i = lane; i < R; i += 32) {
    Srow[i * pitch] = q.d[i];
    arow[i] = (unsigned char)q.act[i];
  }
  __syncwarp();
  return trouble;
}

template <bool NNLS>
__global__ void __launch_bounds__(UPDATE_THREADS, 2)
model_update_kernel(const UpdateParams p) {
  SchedState *st = p.st;
  if ((int)blockIdx.x >= st->n_live)
    return;
  const int m = p.live[blockIdx.x];
  ModelDesc &md = p.models[m];
  const int R = md.rank, col = md.col, cur = st->cur;
  const int rows = p.rows, ld = p.ld, n = p.mode, N = p.n_modes;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int iters = md.iters;

  extern __shared__ double sm[];
  double *Hs = sm;                 // R x R : H, then its Cholesky factor (lower)
  double *Gs = Hs + R * R;         // R x R : new Gramian of mode n
  double *cstat = Gs + R * R;      // 2R : per column reduction value / index ; then lambda, 1/lambda
  double *red = cstat + 2 * R;     // 32
  double *S = red + 32;            // chunk_rows x R, pitch chunk_pitch
  const int pitch = p.chunk_pitch, CR = p.chunk_rows;
  // NNLS scratch (only allocated when p.nnls): per working warp 5R + R*R doubles and 2R + 1 ints, behind S
  double *nn_base = S + (size_t)pitch * p.max_rank;

  double *grams = p.gram_pool + md.gram_off;
  const double *Gm = p.G + (size_t)col * ld;
  double *Fm = p.F[cur] + (size_t)col * ld;

  // ---- H = hadamard of the other Gramians ----
  for (int e = tid; e < R * R; e += nthr) {
    double h = 1.0;
    for (int k = 0; k < N; k++)
      if (k != n)
        h *= grams[(size_t)k * R * R + e];
    Hs[e] = h;
    Gs[e] = 0.0;
  }
  for (int e = tid; e < 2 * R; e += nthr)
    cstat[e] = (e < R) ? -1.0 : 0.0; // value (sum of squares starts at 0 below; max-abs starts at -1), index
  __syncthreads();
  if (iters == 1)
    for (int e = tid; e < R; e += nthr)
      cstat[e] = 0.0;

  // ---- Cholesky, right-looking, in shared memory (unconstrained update only; NNLS needs H itself) ----
  int chol_fail = 0;
  double nnls_tol = 0.0;
  if (NNLS) { // tol = 10 * eps * ||H||_1 * R   (reference src/utils/update.cpp:65-66; one_norm = max column sum)
    double mx = -1.0;
    for (int j = tid; j < R; j += nthr) {
      double cs = 0.0;
      for (int i = 0; i < R; i++)
        cs += fabs(Hs[i + j * R]);
      mx = fmax(mx, cs);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0)
      red[tid >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int wv = 1; wv < (nthr >> 5); wv++)
      mx = fmax(mx, red[wv]);
    nnls_tol = 10 * 2.2204e-16 * mx * (double)R;
    __syncthreads();
  }
  for (int j = 0; j < (NNLS ? 0 : R); j++) {
    __syncthreads();
    const double d = Hs[j + j * R];
    if (!(d > 0.0))
      chol_fail = 1; // LAPACK would stop here (info = j+1); the reference only logs it and carries on
    const double sd = sqrt(d);
    __syncthreads();
    if (tid == 0)
      Hs[j + j * R] = sd;
    for (int i = j + 1 + tid; i < R; i += nthr)
      Hs[i + j * R] /= sd;
    __syncthreads();
    // trailing update of the lower triangle: H[i][k] -= L[i][j] * L[k][j],  j < k <= i < R
    const int t = R - j - 1;
    for (int e = tid; e < t * t; e += nthr) {
      const int i = j + 1 + e % t, k = j + 1 + e / t;
      if (k <= i)
        Hs[i + k * R] -= Hs[i + j * R] * Hs[k + j * R];
    }
  }
  __syncthreads();

  const bool jk_here = (md.jk_mode == n);
  const int jk_row = md.jk_fiber;
  const bool single = rows <= CR;

  // ---- pass 1: solve rows, jackknife row, column statistics ----
  for (int r0 = 0; r0 < rows; r0 += CR) {
    const int nr = min(CR, rows - r0);
    __syncthreads();
    if (p.plan != nullptr) {
      // columns fastest: a warp reads runs of R consecutive doubles of a row-major partial tile
      const PlanView pv = plan_view(p.plan, p.plan_ctas);
      for (int e = tid; e < nr * R; e += nthr) {
        const int j = e % R, rr = e / R;
        const int m = r0 + rr, cg = col + j;
        const int mt = plan_tile_of(m >> 3, pv.In8, pv.m_tiles);
        const int nt = plan_tile_of(cg >> 6, pv.NO, pv.n_tiles);
        const int pair = nt * pv.m_tiles + mt;
        const int s0 = pv.pair_seg0[pair], s1 = pv.pair_seg0[pair + 1];
        const double *q = p.ws + (size_t)s0 * p.tile_elems +
                          (size_t)(m - 8 * plan_m8_start(mt, pv.In8, pv.m_tiles)) * p.n_tile +
                          (cg - 64 * plan_oct_start(nt, pv.NO, pv.n_tiles));
        double sum = 0.0;
        for (int k = s0; k < s1; k++, q += p.tile_elems)
          sum += *q;
        S[rr + j * pitch] = sum;
      }
      if (n == N - 1) { // keep G for the error term
        __syncthreads();
        double *Gw = p.G_out + (size_t)col * ld;
        for (int e = tid; e < nr * R; e += nthr) {
          const int rr = e % nr, j = e / nr;
          Gw[(size_t)j * ld + r0 + rr] = S[rr + j * pitch];
        }
      }
    } else {
      for (int e = tid; e < nr * R; e += nthr) {
        const int rr = e % nr, j = e / nr;
        S[rr + j * pitch] = Gm[(size_t)j * ld + r0 + rr];
      }
    }
    __syncthreads();
    if (NNLS) { // warp per row, active sets warm-started from the previous iteration
      const int warp = tid >> 5, lane = tid & 31;
      if (warp < p.nnls_warps) {
        const size_t per_warp = (size_t)5 * R + (size_t)R * R + (size_t)(2 * R + 2) / 2 + 1; // doubles
        double *wb = nn_base + per_warp * warp;
        NnlsScratch q;
        q.y = wb;
        q.d = wb + R;
        q.w = wb + 2 * R;
        q.s = wb + 3 * R;
        q.sp = wb + 4 * R;
        q.Gp = wb + 5 * R;
        q.act = reinterpret_cast<int *>(wb + 5 * R + (size_t)R * R);
        q.map = q.act + R;
        unsigned char *act = p.act_pool + md.act_off + (size_t)p.rows_before * R;
        int trouble = 0;
        for (int rr = warp; rr < nr; rr += p.nnls_warps)
          trouble += nnls_row(Hs, R, nnls_tol, S + rr, pitch, act + (size_t)(r0 + rr) * R, q, lane);
        if (trouble && lane == 0)
          chol_fail = 1;
      }
    } else {
      for (int rr = tid; rr < nr; rr += nthr) {
        double *row = S + rr;
        // y L^T = g   (forward, right-looking)
        for (int j = 0; j < R; j++) {
          const double y = row[j * pitch] / Hs[j + j * R];
          row[j * pitch] = y;
          for (int k = j + 1; k < R; k++)
            row[k * pitch] -= y * Hs[k + j * R];
        }
        // x L = y     (backward, right-looking)
        for (int j = R - 1; j >= 0; j--) {
          const double x = row[j * pitch] / Hs[j + j * R];
          row[j * pitch] = x;
          for (int k = 0; k < j; k++)
            row[k * pitch] -= x * Hs[j + k * R];
        }
      }
    }
    __syncthreads();
    if (jk_here && jk_row >= r0 && jk_row < r0 + nr) { // Ktensor::set_jk_fiber(0.0)
      for (int j = tid; j < R; j += nthr)
        S[(jk_row - r0) + j * pitch] *= 0.0;
      __syncthreads();
    }
    // column statistics: warp per column
    {
      const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
      for (int j = warp; j < R; j += nw) {
        const double *cj = S + j * pitch;
        if (iters == 1) {
          double ss = 0.0;
          for (int rr = lane; rr < nr; rr += 32)
            ss += cj[rr] * cj[rr];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
          if (lane == 0)
            cstat[j] += ss;
        } else {
          double best = -1.0;
          int bi = 0x7fffffff;
          for (int rr = lane; rr < nr; rr += 32) {
            const double a = fabs(cj[rr]);
            if (a > best) { // strictly greater keeps the first index within this lane's (increasing) sequence
              best = a;
              bi = r0 + rr;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) {
              best = ob;
              bi = oi;
            }
          }
          if (lane == 0 && best > cstat[j]) { // earlier chunks win ties (first index of the maximum, as idamax)
            cstat[j] = best;
            cstat[R + j] = (double)bi;
          }
        }
      }
    }
    if (!single) {
      __syncthreads();
      for (int e = tid; e < nr * R; e += nthr) {
        const int rr = e % nr, j = e / nr;
        Fm[(size_t)j * ld + r0 + rr] = S[rr + j * pitch];
      }
    }
  }
  __syncthreads();

  // ---- lambda ----
  for (int j = tid; j < R; j += nthr) {
    double lam;
    if (iters == 1)
      lam = sqrt(cstat[j]);
    else {
      const int bi = (int)cstat[R + j];
      lam = single ? S[bi + j * pitch] : Fm[(size_t)j * ld + bi];
    }
    p.lambda_home[md.home_col + j] = lam;
    cstat[j] = lam;
    cstat[R + j] = (lam != 0.0) ? 1.0 / lam : 1.0;
  }
  __syncthreads();

  // ---- pass 2: scale, write factor, accumulate Gramian (and <F, G> per column for the error) ----
  const bool last_mode = (n == N - 1);
  double term3 = 0.0;
  for (int r0 = 0; r0 < rows; r0 += CR) {
    const int nr = min(CR, rows - r0);
    __syncthreads();
    for (int e = tid; e < nr * R; e += nthr) {
      const int rr = e % nr, j = e / nr;
      const double raw = single ? S[rr + j * pitch] : Fm[(size_t)j * ld + r0 + rr];
      const double v = raw * cstat[R + j];
      S[rr + j * pitch] = v;
      Fm[(size_t)j * ld + r0 + rr] = v;
      if (last_mode)
        term3 += cstat[j] * v * Gm[(size_t)j * ld + r0 + rr];
    }
    __syncthreads();
    for (int e = tid; e < R * R; e += nthr) {
      const int i = e % R, j = e / R;
      const double *ci = S + i * pitch, *cj = S + j * pitch;
      double g = 0.0;
      for (int rr = 0; rr < nr; rr++)
        g += ci[rr] * cj[rr];
      Gs[e] += g;
    }
  }
  __syncthreads();
  for (int e = tid; e < R * R; e += nthr)
    grams[(size_t)n * R * R + e] = Gs[e];
  if (__syncthreads_or(chol_fail) && tid == 0)
    md.chol_info += 1;

  if (!last_mode)
    return;

  // ---- fast error, fit, eviction decision ----
  double term2 = 0.0;
  for (int e = tid; e < R * R; e += nthr) {
    const int i = e % R, j = e / R;
    double pe = Gs[e];
    for (int k = 0; k < N - 1; k++)
      pe *= grams[(size_t)k * R * R + e];
    term2 += cstat[i] * cstat[j] * pe;
  }
  term2 = block_sum(term2, red);
  term3 = block_sum(term3, red);
  if (tid == 0) {
    double xn = st->x_norm;
    if (md.jk_mode >= 0 && p.x_norms_jk)
      xn = p.x_norms_jk[md.jk_fiber];
    const double err = sqrt(fmax(xn * xn + term2 - 2.0 * term3, 0.0));
    const double old_fit = md.fit;
    const double fit = 1.0 - fabs(err) / st->x_norm;
    md.error = err;
    md.old_fit = old_fit;
    md.fit = fit;
    if (!st->ls_enabled) // with line search the decision is taken after the extrapolation step (ls.cuh)
      decide_eviction(md, st);
  }
}

// Gramians of the incoming factors of every queued model (MultiKtensor::add computes them on admission,
// reference src/multi_ktensor.cpp:89-95; here all at once before the loop starts).  One CTA per (model, mode).
__global__ void __launch_bounds__(256)
init_grams_kernel(const Geom geo, const FactorPtrs fac, const ModelDesc *__restrict__ models, double *gram_pool) {
  const int m = blockIdx.x, n = blockIdx.y;
  const ModelDesc &md = models[m];
  const int R = md.rank, rows = geo.dims[n], ld = geo.ldF[n];
  const double *F = fac.home[n] + (size_t)md.home_col * ld;
  double *gr = gram_pool + md.gram_off + (size_t)n * R * R;
  for (int e = threadIdx.x; e < R * R; e += blockDim.x) {
    const int i = e % R, j = e / R;
    const double *ci = F + (size_t)i * ld, *cj = F + (size_t)j * ld;
    double g = 0.0;
    for (int r = 0; r < rows; r++)
      g += ci[r] * cj[r];
    gr[e] = g;
  }
}

} // namespace calsb200
