/* TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the CP-CALS hot path.  NOT the product.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library;
 * the product path (cp-cals_b200/csrc) never calls into it and has no CPU fallback.
 *
 * Plain C (no BLAS/LAPACK) restatement of the per-iteration body of cals::cp_cals, written from the
 * algorithm description, each function citing the reference file:line it follows (paths relative to
 * /root/reference).  The reference's BLAS/LAPACK calls (dgemm, dtrsm, dpotrf, dnrm2, idamax, dscal; external
 * OpenBLAS, un-vendored) are restated by their textbook definitions.
 *
 * Parity pin: tests/test_oracle.py checks this file against the UNMODIFIED reference library (oracle/_ref, built
 * from /root/reference by oracle/build_ref.sh; live comparison where that build exists) and against
 * tests/golden/ *.npz, which were produced by that same reference build (oracle/make_golden.py).  The reference itself ships no golden vectors
 * (its tests are relational: CALS == ALS, fast error == explicit error).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int64_t rank;
  int64_t jk_mode;  /* -1: regular model (reference include/ktensor.h:18-22) */
  int64_t jk_fiber;
  double *factors;  /* in/out: mode-0 block (I_0 x R col-major), then mode 1, ... */
  double *lambda;   /* out: R */
  int64_t iters;    /* out */
  double error;     /* out: approx_error */
  double fit;       /* out */
  double old_fit;   /* out */
  int64_t chol_fail; /* out: number of non-positive Cholesky pivots met (reference only logs dpotrf info) */
  unsigned char *active; /* NNLS only, in/out, may be NULL (= all constraints active, a fresh Ktensor): per mode an
                          * I_n x R block, row-major [row * R + col], 1 = constrained to zero (reference
                          * include/ktensor.h:36 active_set) */
} cals_oracle_model;

typedef struct {
  int64_t iter;          /* CalsReport::iter  (global iterations) */
  int64_t n_ktensors;    /* CalsReport::n_ktensors */
  int64_t ktensor_comp_sum;
  double x_norm;
  int64_t ls_performed; /* CalsReport::ls_performed / ls_failed (include/cals.h:52-53) */
  int64_t ls_failed;
} cals_oracle_report;

/* Line-search options (reference include/utils/line_search.h:8-39, CalsParams include/cals.h:153-156).
 * method: 0 NO_ERROR_CHECKING, 1 ERROR_CHECKING_SERIAL.  step == 0 -> cbrt(model iteration) (src/cals.cpp:317-318). */
typedef struct {
  int enabled;
  int method;
  int interval;
  double step;
} cals_oracle_ls;

/* ---------------------------------------------------------------------------------------------------------------
 * Tensor::norm  (include/tensor.h:196, cblas_dnrm2 over all elements), called at src/cals.cpp:36. */
double cals_oracle_norm(int64_t n, const double *x) {
  double s = 0.0;
  for (int64_t i = 0; i < n; i++)
    s += x[i] * x[i];
  return sqrt(s);
}

/* utils::calculate_jackknifing_norms  (src/utils/utils.cpp:103-152):
 * out[i] = sqrt( sum_all X^2 - sum_j X(i,j)^2 ), X viewed as I_0 x (nX/I_0) column-major. */
void cals_oracle_jk_norms(int n_modes, const int64_t *modes, const double *X, double *out) {
  int64_t I0 = modes[0], cols = 1;
  for (int n = 1; n < n_modes; n++)
    cols *= modes[n];
  for (int64_t i = 0; i < I0; i++)
    out[i] = 0.0;
  for (int64_t j = 0; j < cols; j++)
    for (int64_t i = 0; i < I0; i++)
      out[i] += X[i + j * I0] * X[i + j * I0];
  double total = 0.0;
  for (int64_t i = 0; i < I0; i++)
    total += out[i];
  for (int64_t i = 0; i < I0; i++)
    out[i] = sqrt(total - out[i]);
}

/* ---------------------------------------------------------------------------------------------------------------
 * MTTKRP for one mode over C concatenated columns (src/utils/mttkrp.cpp:562-614; semantics of mttkrp_impl
 * :218-328 = X_(n) * KRP with the KRP rows ordered lowest remaining mode fastest, :179-216).
 *   G[i,c] = sum_{i_k, k != n} X[i_0..i_{N-1}] * prod_{k != n} A_k[i_k, c]
 * X column-major (mode 0 fastest).  factors[k] is I_k x C with leading dimension ld[k]; G is I_n x C, ldG.
 * Evaluation order: three-way view (L, I_n, U): inner sum over the lower modes, then over the upper modes
 * (one of the orders the reference itself uses, mttkrp_twostep :450-560; its own tests accept all orders). */
static void build_krp_column(int n_lo, int n_hi, const int64_t *modes, const double *const *factors,
                             const int64_t *ld, int64_t c, double *out) {
  /* out[r] for r enumerating modes n_lo..n_hi-1, lowest fastest */
  int64_t len = 1;
  out[0] = 1.0;
  for (int k = n_lo; k < n_hi; k++) {
    const double *a = factors[k] + c * ld[k];
    /* expand in place from the back so that mode k becomes the slowest index so far */
    for (int64_t ik = modes[k] - 1; ik >= 0; ik--)
      for (int64_t r = len - 1; r >= 0; r--)
        out[ik * len + r] = out[r] * a[ik];
    len *= modes[k];
  }
}

void cals_oracle_mttkrp(int n_modes, const int64_t *modes, const double *X, int mode, int64_t C,
                        const double *const *factors, const int64_t *ld, double *G, int64_t ldG) {
  int64_t L = 1, U = 1, In = modes[mode];
  for (int k = 0; k < mode; k++)
    L *= modes[k];
  for (int k = mode + 1; k < n_modes; k++)
    U *= modes[k];
#pragma omp parallel
  {
    double *kl = (double *)malloc(sizeof(double) * (size_t)L);
    double *ku = (double *)malloc(sizeof(double) * (size_t)U);
#pragma omp for schedule(dynamic, 1)
    for (int64_t c = 0; c < C; c++) {
      build_krp_column(0, mode, modes, factors, ld, c, kl);
      build_krp_column(mode + 1, n_modes, modes, factors, ld, c, ku);
      double *g = G + c * ldG;
      for (int64_t i = 0; i < In; i++)
        g[i] = 0.0;
      for (int64_t u = 0; u < U; u++) {
        const double *Xu = X + u * L * In;
        double w = ku[u];
        if (L == 1) {
          for (int64_t i = 0; i < In; i++)
            g[i] += w * Xu[i];
        } else {
          for (int64_t i = 0; i < In; i++) {
            const double *x = Xu + i * L;
            double s = 0.0;
            for (int64_t l = 0; l < L; l++)
              s += x[l] * kl[l];
            g[i] += w * s;
          }
        }
      }
    }
    free(kl);
    free(ku);
  }
}

/* ---------------------------------------------------------------------------------------------------------------
 * Per-model small operations.  All matrices column-major. */

/* ops::update_gramian (src/utils/utils.cpp:174-178): gram = F^T F, full R x R. */
static void gramian(const double *F, int64_t rows, int64_t R, double *gram) {
  for (int64_t j = 0; j < R; j++)
    for (int64_t i = 0; i < R; i++) {
      double s = 0.0;
      for (int64_t r = 0; r < rows; r++)
        s += F[r + i * rows] * F[r + j * rows];
      gram[i + j * R] = s;
    }
}

/* ops::hadamard_but_one (src/utils/utils.cpp:161-172): H = prod_{k != mode} gram_k, written into gram_mode. */
static void hadamard_but_one(double *const *grams, int n_modes, int mode, int64_t R) {
  double *H = grams[mode];
  for (int64_t e = 0; e < R * R; e++)
    H[e] = 1.0;
  for (int k = 0; k < n_modes; k++)
    if (k != mode)
      for (int64_t e = 0; e < R * R; e++)
        H[e] *= grams[k][e];
}

/* dpotrf('L') as called at src/utils/update.cpp:183: unblocked lower Cholesky in place.  Returns the LAPACK-style
 * info (0 ok, j+1 = first non-positive pivot; the factorisation stops there, as LAPACK's does). */
static int64_t cholesky_lower(double *A, int64_t R) {
  for (int64_t j = 0; j < R; j++) {
    double d = A[j + j * R];
    for (int64_t k = 0; k < j; k++)
      d -= A[j + k * R] * A[j + k * R];
    if (!(d > 0.0))
      return j + 1;
    d = sqrt(d);
    A[j + j * R] = d;
    for (int64_t i = j + 1; i < R; i++) {
      double s = A[i + j * R];
      for (int64_t k = 0; k < j; k++)
        s -= A[i + k * R] * A[j + k * R];
      A[i + j * R] = s / d;
    }
  }
  return 0;
}

/* update::update_factor_unconstrained (src/utils/update.cpp:178-192):
 * F <- G * L^-T (dtrsm Right,Lower,Trans) then F <- F * L^-1 (dtrsm Right,Lower,NoTrans), in place. */
static void solve_in_place(double *F, int64_t rows, int64_t R, const double *Lm) {
  for (int64_t r = 0; r < rows; r++) {
    /* y L^T = g  : forward */
    for (int64_t j = 0; j < R; j++) {
      double s = F[r + j * rows];
      for (int64_t k = 0; k < j; k++)
        s -= F[r + k * rows] * Lm[j + k * R];
      F[r + j * rows] = s / Lm[j + j * R];
    }
    /* x L = y : backward */
    for (int64_t j = R - 1; j >= 0; j--) {
      double s = F[r + j * rows];
      for (int64_t k = j + 1; k < R; k++)
        s -= F[r + k * rows] * Lm[k + j * R];
      F[r + j * rows] = s / Lm[j + j * R];
    }
  }
}


/* ---------------------------------------------------------------------------------------------------------------
 * update::update_factor_non_negative_constrained (src/utils/update.cpp:61-176): row-wise active-set NNLS
 * (Bro & de Jong's fast NNLS on the normal equations), warm-started from the active set of the previous iteration.
 * F holds the MTTKRP result G on entry (rows x R, ld = rows), H is the Hadamard product of the other Gramians
 * (R x R, NOT factored), active is rows x R row-major.  dposv('L') (src/utils/update.cpp:40) is restated as an
 * unblocked Cholesky + two triangular solves.  Returns the number of rows on which a Cholesky of a passive block
 * failed inside the main loop (the reference would terminate on the uncaught CholFail there). */
static int nnls_solve_passive(const double *H, int64_t R, const double *y, const unsigned char *act, double *Gp,
                              double *sp, int64_t *map, int64_t *p_out) {
  int64_t p = 0;
  for (int64_t i = 0; i < R; i++)
    if (!act[i])
      map[p++] = i;
  *p_out = p;
  for (int64_t a = 0; a < p; a++) {
    sp[a] = y[map[a]];
    for (int64_t b = 0; b < p; b++)
      Gp[a + b * p] = H[map[a] + map[b] * R];
  }
  if (cholesky_lower(Gp, p) != 0)
    return 1; /* CholFail (src/utils/update.cpp:43-44) */
  for (int64_t j = 0; j < p; j++) { /* L z = y_p */
    double t = sp[j];
    for (int64_t k = 0; k < j; k++)
      t -= Gp[j + k * p] * sp[k];
    sp[j] = t / Gp[j + j * p];
  }
  for (int64_t j = p - 1; j >= 0; j--) { /* L^T x = z */
    double t = sp[j];
    for (int64_t k = j + 1; k < p; k++)
      t -= Gp[k + j * p] * sp[k];
    sp[j] = t / Gp[j + j * p];
  }
  return 0;
}

static double vec_min(const double *v, int64_t n) {
  double m = v[0];
  for (int64_t i = 1; i < n; i++)
    if (v[i] < m)
      m = v[i];
  return m;
}

static int64_t nnls_update(double *F, int64_t rows, int64_t R, const double *H, unsigned char *active) {
  /* tol = 10 * eps * ||H||_1 * n  (src/utils/update.cpp:65-66; one_norm = max column sum, include/matrix.h:113-121) */
  double one_norm = -1.0e308;
  for (int64_t c = 0; c < R; c++) {
    double t = 0.0;
    for (int64_t r = 0; r < R; r++)
      t += fabs(H[r + c * R]);
    if (t > one_norm)
      one_norm = t;
  }
  const double tol = 10 * 2.2204e-16 * one_norm * (double)R;
  double *y = (double *)malloc(sizeof(double) * (size_t)(5 * R + R * R));
  double *d = y + R, *w = d + R, *s = w + R, *sp = s + R, *Gp = sp + R;
  int64_t *map = (int64_t *)malloc(sizeof(int64_t) * (size_t)R);
  int64_t failures = 0;
  for (int64_t row = 0; row < rows; row++) {
    unsigned char *act = active + row * R;
    int64_t p = 0;
    int any_passive = 0;
    for (int64_t i = 0; i < R; i++) {
      d[i] = 0.0;
      y[i] = F[row + rows * i];
      if (y[i] > 0)
        act[i] = 0;
      any_passive |= !act[i];
    }
    if (any_passive) { /* warm start from the previous active set (:88-117) */
      int failed = nnls_solve_passive(H, R, y, act, Gp, sp, map, &p);
      while (!failed) {
        for (int64_t a = 0, i = 0; i < R; i++)
          d[i] = act[i] ? 0.0 : sp[a++];
        if (!(vec_min(sp, p) <= tol))
          break;
        int64_t left = 0;
        for (int64_t i = 0; i < R; i++) {
          if (d[i] <= tol) {
            d[i] = 0.0;
            act[i] = 1;
          }
          left += !act[i];
        }
        if (left == 0) { /* ZeroPassiveSet */
          failed = 1;
          break;
        }
        failed = nnls_solve_passive(H, R, y, act, Gp, sp, map, &p);
      }
      if (failed) { /* catch block (:112-115) */
        for (int64_t i = 0; i < R; i++) {
          act[i] = 1;
          d[i] = 0.0;
        }
      }
    }
    /* w = y - H d  (:119, :50-57) */
    for (int64_t i = 0; i < R; i++) {
      double t = 0.0;
      for (int64_t j = 0; j < R; j++)
        t += H[i + j * R] * d[j];
      w[i] = y[i] - t;
    }
    for (;;) { /* main loop (:122-170) */
      int64_t m = -1;
      double best = -1.0e308;
      for (int64_t i = 0; i < R; i++)
        if (act[i] && w[i] > best) { /* Tensor::max_id: first index of the maximum among the active ones */
          best = w[i];
          m = i;
        }
      if (m < 0 || !(best > tol))
        break;
      act[m] = 0;
      if (nnls_solve_passive(H, R, y, act, Gp, sp, map, &p)) {
        failures++;
        break;
      }
      int bad = 0;
      while (vec_min(sp, p) <= tol) { /* inner loop (:131-155) */
        for (int64_t a = 0, i = 0; i < R; i++)
          s[i] = act[i] ? 0.0 : sp[a++];
        double alpha = 1.7976931348623157e308;
        for (int64_t i = 0; i < R; i++)
          if (!act[i] && s[i] <= tol) {
            double t = d[i] / (d[i] - s[i]);
            if (t < alpha)
              alpha = t;
          }
        int64_t left = 0;
        for (int64_t i = 0; i < R; i++) {
          d[i] = d[i] + alpha * (s[i] - d[i]);
          if (fabs(d[i]) < tol && !act[i]) {
            act[i] = 1;
            d[i] = 0.0;
          }
          left += !act[i];
        }
        if (left == 0 || nnls_solve_passive(H, R, y, act, Gp, sp, map, &p)) {
          bad = 1;
          break;
        }
      }
      if (bad) {
        failures++;
        break;
      }
      for (int64_t a = 0, i = 0; i < R; i++)
        d[i] = act[i] ? 0.0 : sp[a++];
      for (int64_t i = 0; i < R; i++) {
        double t = 0.0;
        for (int64_t j = 0; j < R; j++)
          t += H[i + j * R] * d[j];
        w[i] = y[i] - t;
      }
    }
    for (int64_t i = 0; i < R; i++)
      F[row + rows * i] = d[i];
  }
  free(y);
  free(map);
  return failures;
}

/* Ktensor::normalize(mode, iteration) (src/ktensor.cpp:66-83). */
static void normalize_mode(double *F, int64_t rows, int64_t R, int64_t iteration, double *lambda) {
  for (int64_t c = 0; c < R; c++) {
    double *col = F + c * rows;
    double l;
    if (iteration == 1) {
      double s = 0.0;
      for (int64_t r = 0; r < rows; r++)
        s += col[r] * col[r];
      l = sqrt(s);
    } else {
      int64_t idx = 0;
      double best = -1.0;
      for (int64_t r = 0; r < rows; r++)
        if (fabs(col[r]) > best) { /* idamax: first index of the largest magnitude */
          best = fabs(col[r]);
          idx = r;
        }
      l = col[idx];
    }
    lambda[c] = l;
    if (l != 0.0) {
      double inv = 1.0 / l;
      for (int64_t r = 0; r < rows; r++)
        col[r] *= inv;
    }
  }
}

/* error::compute_fast_error (src/utils/error.cpp:64-89) on P = hadamard of all gramians (utils.cpp:156-159). */
static double fast_error(double x_norm, const double *lambda, const double *F_last, const double *G_last,
                         int64_t rows, int64_t R, double *const *grams, int n_modes) {
  double term2 = 0.0, term3 = 0.0;
  for (int64_t j = 0; j < R; j++)
    for (int64_t i = 0; i < R; i++) {
      double p = grams[0][i + j * R];
      for (int k = 1; k < n_modes; k++)
        p *= grams[k][i + j * R];
      term2 += lambda[i] * lambda[j] * p;
    }
  for (int64_t j = 0; j < R; j++)
    for (int64_t i = 0; i < rows; i++)
      term3 += lambda[j] * F_last[i + j * rows] * G_last[i + j * rows];
  return sqrt(fmax(x_norm * x_norm + term2 - 2.0 * term3, 0.0));
}


/* ---------------------------------------------------------------------------------------------------------------
 * Line search (src/utils/line_search.cpp).  A model's state is (factors, lambda, error, fit, old_fit, iters); the
 * reference keeps two extra Ktensors per live model: prev_ktensor (snapshot taken `interval-1` iterations after the
 * last extrapolation, src/cals.cpp:203-212) and backup_ktensor (NO_ERROR_CHECKING only: state right before an
 * extrapolation, to return to if the error went up). */
typedef struct {
  double *factors, *lambda; /* same layout as cals_oracle_model */
  double error, fit, old_fit;
  int64_t iters;
} ls_snapshot;

static void snapshot_take(ls_snapshot *s, const cals_oracle_model *m, int64_t n_fac) { /* Ktensor::copy, ktensor.cpp:162-179 */
  memcpy(s->factors, m->factors, sizeof(double) * (size_t)n_fac);
  memcpy(s->lambda, m->lambda, sizeof(double) * (size_t)m->rank);
  s->error = m->error;
  s->fit = m->fit;
  s->old_fit = m->old_fit;
  s->iters = m->iters;
}
static void snapshot_restore(const ls_snapshot *s, cals_oracle_model *m, int64_t n_fac) {
  memcpy(m->factors, s->factors, sizeof(double) * (size_t)n_fac);
  memcpy(m->lambda, s->lambda, sizeof(double) * (size_t)m->rank);
  m->error = s->error;
  m->fit = s->fit;
  m->old_fit = s->old_fit;
  m->iters = s->iters;
}
/* Ktensor::denormalize (src/ktensor.cpp:101-107) */
static void denormalize(double *factors, const double *lambda, int64_t rows0, int64_t R) {
  for (int64_t c = 0; c < R; c++)
    for (int64_t r = 0; r < rows0; r++)
      factors[r + c * rows0] *= lambda[c];
}
/* Ktensor::normalize() (src/ktensor.cpp:85-99): every column to unit 2-norm, lambda = product of the norms */
static void normalize_all(int n_modes, const int64_t *modes, int64_t R, double *factors, double *lambda) {
  for (int64_t c = 0; c < R; c++)
    lambda[c] = 1.0;
  int64_t off = 0;
  for (int n = 0; n < n_modes; n++) {
    double *F = factors + off;
    for (int64_t c = 0; c < R; c++) {
      double s = 0.0;
      for (int64_t r = 0; r < modes[n]; r++)
        s += F[r + c * modes[n]] * F[r + c * modes[n]];
      double nrm = sqrt(s), inv = 1.0 / nrm;
      for (int64_t r = 0; r < modes[n]; r++)
        F[r + c * modes[n]] *= inv;
      lambda[c] *= nrm;
    }
    off += modes[n] * R;
  }
}
static void all_gramians(int n_modes, const int64_t *modes, int64_t R, const double *factors, double *const *grams) {
  int64_t off = 0;
  for (int n = 0; n < n_modes; n++) {
    gramian(factors + off, modes[n], R, grams[n]);
    off += modes[n] * R;
  }
}
/* error::compute_error (src/utils/error.cpp:7-31; 3 modes there, any N here): ||X - [[lambda; factors]]||_F by explicit
 * reconstruction. */
static double explicit_error(int n_modes, const int64_t *modes, const double *X, int64_t R, const double *factors,
                             const double *lambda) {
  int64_t nX = 1, foff[17];
  foff[0] = 0;
  for (int n = 0; n < n_modes; n++) {
    nX *= modes[n];
    foff[n + 1] = foff[n] + modes[n] * R;
  }
  int64_t idx[16] = {0};
  double *w = (double *)malloc(sizeof(double) * (size_t)R);
  double acc = 0.0;
  const int64_t I0 = modes[0];
  for (int64_t o = 0; o < nX / I0; o++) {
    for (int64_t r = 0; r < R; r++) {
      double p = lambda[r];
      for (int n = 1; n < n_modes; n++)
        p *= factors[foff[n] + idx[n] + r * modes[n]];
      w[r] = p;
    }
    for (int64_t i = 0; i < I0; i++) {
      double m = 0.0;
      for (int64_t r = 0; r < R; r++)
        m += w[r] * factors[i + r * I0];
      double dlt = X[o * I0 + i] - m;
      acc += dlt * dlt;
    }
    for (int n = 1; n < n_modes; n++) {
      if (++idx[n] < modes[n])
        break;
      idx[n] = 0;
    }
  }
  free(w);
  return sqrt(acc);
}

/* ---------------------------------------------------------------------------------------------------------------
 * cals::cp_cals (src/cals.cpp:19-395) with the MultiKtensor buffer bookkeeping (src/multi_ktensor.cpp).
 *
 * flags: bit0 force_max_iter, bit1 always_evict_first, bit2 update_method == NNLS.  Models are consumed FIFO from
 * models[0..n_models).
 * Because models never interact numerically, the buffer only decides WHEN a model runs (global iteration count,
 * admission order); it is restated through the occupancy vector exactly as the reference keeps it. */
typedef struct {
  int live;
  int64_t col;
  double **grams; /* n_modes matrices R x R */
  double *G;      /* scratch: max_rows x R  (MTTKRP result for the current mode) */
  double *G_last;
  unsigned char *active; /* NNLS: the model's active sets (caller's buffer, or owned when the caller passed none) */
  int own_active;
  /* line search (RegistryEntry::ls_params, include/multi_ktensor.h:20) */
  int ls_iter, ls_updated_last;
  ls_snapshot prev, backup;
} live_entry;

int cals_oracle_cp_cals_ls(int n_modes, const int64_t *modes, const double *X, int64_t n_models,
                           cals_oracle_model *models, int64_t max_iter, double tol, int64_t buffer_size, int flags,
                           const cals_oracle_ls *lsp, cals_oracle_report *rep);

int cals_oracle_cp_cals(int n_modes, const int64_t *modes, const double *X, int64_t n_models,
                        cals_oracle_model *models, int64_t max_iter, double tol, int64_t buffer_size, int flags,
                        cals_oracle_report *rep) {
  cals_oracle_ls off = {0, 0, 0, 0.0};
  return cals_oracle_cp_cals_ls(n_modes, modes, X, n_models, models, max_iter, tol, buffer_size, flags, &off, rep);
}

int cals_oracle_cp_cals_ls(int n_modes, const int64_t *modes, const double *X, int64_t n_models,
                           cals_oracle_model *models, int64_t max_iter, double tol, int64_t buffer_size, int flags,
                           const cals_oracle_ls *lsp, cals_oracle_report *rep) {
  if (n_modes > 16)
    return -1;
  const cals_oracle_ls ls = *lsp;
  const int force_max_iter = flags & 1, always_evict_first = flags & 2, nnls = flags & 4;
  int64_t nX = 1, max_rows = 0, sum_rows = 0;
  for (int n = 0; n < n_modes; n++)
    sum_rows += modes[n];
  for (int n = 0; n < n_modes; n++) {
    nX *= modes[n];
    if (modes[n] > max_rows)
      max_rows = modes[n];
  }
  const double x_norm = cals_oracle_norm(nX, X);
  double *x_norms_jk = NULL;

  int64_t *occupancy = (int64_t *)calloc((size_t)buffer_size, sizeof(int64_t)); /* 0 free, else model index+1 */
  live_entry *live = (live_entry *)calloc((size_t)n_models, sizeof(live_entry));
  int64_t next = 0, n_live = 0;
  rep->iter = 0;
  rep->n_ktensors = 0;
  rep->ktensor_comp_sum = 0;
  rep->x_norm = x_norm;
  rep->ls_performed = 0;
  rep->ls_failed = 0;
  for (int64_t m = 0; m < n_models; m++) {
    models[m].iters = 0;
    models[m].error = 0.0;
    models[m].fit = 0.0;
    models[m].old_fit = 0.0;
    models[m].chol_fail = 0;
  }

  do {
    rep->iter++;
    /* admission: FIFO, first-fit, stop at the first model that does not fit (cals.cpp:182-192,
     * multi_ktensor.cpp:14-39) */
    while (next < n_models) {
      int64_t R = models[next].rank, run = 0, pos = -1;
      for (int64_t i = 0; i < buffer_size && run < R; i++) {
        if (occupancy[i] == 0) {
          if (run == 0)
            pos = i;
          run++;
        } else
          run = 0;
      }
      if (pos < 0 || run != R)
        break;
      cals_oracle_model *mm = &models[next];
      live_entry *le = &live[next];
      for (int64_t i = 0; i < R; i++)
        occupancy[pos + i] = next + 1;
      le->live = 1;
      le->col = pos;
      le->grams = (double **)malloc(sizeof(double *) * (size_t)n_modes);
      le->G = (double *)malloc(sizeof(double) * (size_t)(max_rows * R));
      le->G_last = (double *)malloc(sizeof(double) * (size_t)(modes[n_modes - 1] * R));
      int64_t off = 0;
      for (int n = 0; n < n_modes; n++) { /* gramians of the incoming factors (multi_ktensor.cpp:89-95) */
        le->grams[n] = (double *)malloc(sizeof(double) * (size_t)(R * R));
        gramian(mm->factors + off, modes[n], R, le->grams[n]);
        off += modes[n] * R;
      }
      le->active = mm->active;
      le->own_active = 0;
      if (nnls && !le->active) { /* fresh Ktensor: every constraint active (include/ktensor.h:66) */
        le->active = (unsigned char *)malloc((size_t)(sum_rows * R));
        memset(le->active, 1, (size_t)(sum_rows * R));
        le->own_active = 1;
      }
      le->ls_iter = 0;
      le->ls_updated_last = 0;
      if (ls.enabled) { /* multi_ktensor.cpp:103-111 */
        ls_snapshot *snaps[2] = {&le->prev, &le->backup};
        for (int k = 0; k < 2; k++) {
          snaps[k]->factors = (double *)calloc((size_t)(sum_rows * R), sizeof(double));
          snaps[k]->lambda = (double *)calloc((size_t)R, sizeof(double));
          snaps[k]->error = snaps[k]->fit = snaps[k]->old_fit = 0.0;
          snaps[k]->iters = 0;
        }
      }
      mm->iters = 1; /* multi_ktensor.cpp:96 */
      if (mm->jk_mode >= 0 && !x_norms_jk) { /* cals.cpp:198-200 */
        x_norms_jk = (double *)malloc(sizeof(double) * (size_t)modes[0]);
        cals_oracle_jk_norms(n_modes, modes, X, x_norms_jk);
      }
      rep->n_ktensors++;
      rep->ktensor_comp_sum += R;
      n_live++;
      next++;
    }

    /* line search: time to remember the model as it is now? (cals.cpp:203-212) */
    if (ls.enabled)
      for (int64_t m = 0; m < n_models; m++)
        if (live[m].live && live[m].ls_iter == ls.interval - 1)
          snapshot_take(&live[m].prev, &models[m], sum_rows * models[m].rank);

    /* modes loop (cals.cpp:220-276); per model the concatenated MTTKRP is its own columns' MTTKRP.
     * Models are independent (the reference runs them under `omp parallel for`, cals.cpp:239,281). */
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t m = 0; m < n_models; m++) {
      if (!live[m].live)
        continue;
      cals_oracle_model *mm = &models[m];
      live_entry *le = &live[m];
      int64_t R = mm->rank;
      int64_t foff[17];
      foff[0] = 0;
      for (int n = 0; n < n_modes; n++)
        foff[n + 1] = foff[n] + modes[n] * R;
      const double *fptr[16];
      int64_t ld[16];
      for (int n = 0; n < n_modes; n++) {
        int64_t rows = modes[n];
        for (int k = 0; k < n_modes; k++) {
          fptr[k] = mm->factors + foff[k];
          ld[k] = modes[k];
        }
        cals_oracle_mttkrp(n_modes, modes, X, n, R, fptr, ld, le->G, rows);
        if (n == n_modes - 1) /* cals.cpp:230-234 */
          memcpy(le->G_last, le->G, sizeof(double) * (size_t)(rows * R));
        double *F = mm->factors + foff[n];
        memcpy(F, le->G, sizeof(double) * (size_t)(rows * R)); /* G overwrites factor n (mttkrp.cpp:327,613) */
        hadamard_but_one(le->grams, n_modes, n, R);                 /* cals.cpp:242 */
        if (nnls) {                                                 /* cals.cpp:247-248 */
          int64_t aoff = 0;
          for (int k = 0; k < n; k++)
            aoff += modes[k] * R;
          mm->chol_fail += nnls_update(F, rows, R, le->grams[n], le->active + aoff);
        } else {
          if (cholesky_lower(le->grams[n], R) != 0)                 /* update.cpp:183-185 */
            mm->chol_fail++;
          solve_in_place(F, rows, R, le->grams[n]);                 /* update.cpp:187-190 */
        }
        if (mm->jk_mode == n)                                       /* cals.cpp:250-251, ktensor.h:316-325 */
          for (int64_t c = 0; c < R; c++)
            F[mm->jk_fiber + c * rows] *= 0.0;
        normalize_mode(F, rows, R, mm->iters, mm->lambda);          /* cals.cpp:253 */
        gramian(F, rows, R, le->grams[n]);                          /* cals.cpp:255 */
      }
      /* error + fit (cals.cpp:281-303) */
      double xn = x_norm;
      if (mm->jk_mode >= 0)
        xn = x_norms_jk[mm->jk_fiber];
      int64_t rows_last = modes[n_modes - 1];
      mm->error = fast_error(xn, mm->lambda, mm->factors + foff[n_modes - 1], le->G_last, rows_last, R, le->grams,
                             n_modes);
      mm->old_fit = mm->fit; /* ktensor.h:178-183: fit uses the GLOBAL norm, also for jk models (cals.cpp:302) */
      mm->fit = 1.0 - fabs(mm->error) / x_norm;

      /* line search (cals.cpp:309-333, line_search.cpp:219-268) */
      if (ls.enabled && !(ls.method == 0 && mm->iters >= max_iter)) {
        const int64_t n_fac = sum_rows * R;
        const double step = ls.step == 0.0 ? cbrt((double)mm->iters) : ls.step;
        int extrapolated = 0, reversed = 0;
        le->ls_iter++;
        if (ls.method == 0) {
          if (le->ls_updated_last) {
            le->ls_updated_last = 0;
            if (le->backup.error < mm->error) { /* the extrapolation made it worse: go back */
              reversed = 1;
              le->ls_iter = 0;
              snapshot_restore(&le->backup, mm, n_fac);
              all_gramians(n_modes, modes, R, mm->factors, le->grams);
            }
          }
          if (le->ls_iter == ls.interval) {
            extrapolated = 1;
            le->ls_iter = 0;
            le->ls_updated_last = 1;
            snapshot_take(&le->backup, mm, n_fac);
            /* line_search_no_error_checking (line_search.cpp:23-68) */
            denormalize(mm->factors, mm->lambda, modes[0], R);
            denormalize(le->prev.factors, le->prev.lambda, modes[0], R);
            for (int64_t e = 0; e < n_fac; e++)
              mm->factors[e] += step * (mm->factors[e] - le->prev.factors[e]);
            normalize_all(n_modes, modes, R, mm->factors, mm->lambda);
            mm->error = 1.7976931348623157e308;
            mm->old_fit = mm->fit;
            mm->fit = 1.0 - fabs(mm->error) / 1.0;
            all_gramians(n_modes, modes, R, mm->factors, le->grams);
          }
        } else if (le->ls_iter == ls.interval) { /* ERROR_CHECKING_SERIAL: line_search_error_checking (:84-150) */
          extrapolated = 1;
          le->ls_iter = 0;
          double *lsf = le->prev.factors; /* the snapshot doubles as the trial model (ls_ktensor) */
          for (int64_t e = 0; e < n_fac; e++)
            lsf[e] = mm->factors[e] + step * (mm->factors[e] - lsf[e]);
          memcpy(le->prev.lambda, mm->lambda, sizeof(double) * (size_t)R);
          /* compute_error: denormalize, ||X - M||, normalize() (error.cpp:13,28) */
          denormalize(lsf, le->prev.lambda, modes[0], R);
          double ones[64];
          double *one_l = R <= 64 ? ones : (double *)malloc(sizeof(double) * (size_t)R);
          for (int64_t c = 0; c < R; c++)
            one_l[c] = 1.0;
          const double e_new = explicit_error(n_modes, modes, X, R, lsf, one_l);
          if (one_l != ones)
            free(one_l);
          normalize_all(n_modes, modes, R, lsf, le->prev.lambda);
          reversed = 1;
          if (e_new < mm->error) {
            reversed = 0;
            memcpy(mm->factors, lsf, sizeof(double) * (size_t)n_fac); /* factors only: lambda stays (:130-131) */
            all_gramians(n_modes, modes, R, mm->factors, le->grams);
            mm->error = e_new;
            mm->old_fit = mm->fit;
            mm->fit = 1.0 - fabs(mm->error) / x_norm;
          }
        }
#pragma omp atomic
        rep->ls_performed += extrapolated;
#pragma omp atomic
        rep->ls_failed += reversed;
      }
    }

    /* eviction (cals.cpp:336-358) in registry (= admission) order */
    if (!always_evict_first) {
      for (int64_t m = 0; m < n_models; m++) {
        if (!live[m].live)
          continue;
        cals_oracle_model *mm = &models[m];
        int evict;
        if (!force_max_iter)
          evict = (fabs(mm->old_fit - mm->fit) < tol) || (mm->iters >= max_iter);
        else
          evict = mm->iters >= max_iter;
        if (evict)
          live[m].live = 2; /* marked */
        else
          mm->iters++;
      }
    } else if (n_live > 0 && occupancy[0] != 0) {
      live[occupancy[0] - 1].live = 2; /* leftmost model only (cals.cpp:348-354) */
    }
    for (int64_t m = 0; m < n_models; m++)
      if (live[m].live == 2) {
        live_entry *le = &live[m];
        for (int64_t i = 0; i < buffer_size; i++)
          if (occupancy[i] == m + 1)
            occupancy[i] = 0;
        for (int n = 0; n < n_modes; n++)
          free(le->grams[n]);
        free(le->grams);
        free(le->G);
        free(le->G_last);
        if (le->own_active)
          free(le->active);
        if (ls.enabled) {
          free(le->prev.factors);
          free(le->prev.lambda);
          free(le->backup.factors);
          free(le->backup.lambda);
        }
        le->live = 0;
        n_live--;
      }
    /* compress: stable shift of the live models to the left (multi_ktensor.cpp:188-264) */
    {
      int64_t w = 0;
      for (int64_t i = 0; i < buffer_size; i++)
        if (occupancy[i] != 0) {
          int64_t id = occupancy[i];
          occupancy[i] = 0;
          occupancy[w] = id;
          if (w == 0 || occupancy[w - 1] != id)
            live[id - 1].col = w;
          w++;
        }
    }
  } while (!(next >= n_models && n_live == 0)); /* cals.cpp:379-380 */

  free(occupancy);
  free(live);
  free(x_norms_jk);
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Helpers used by jk_cp_cals (src/cals.cpp:397-446): Ktensor::denormalize (src/ktensor.cpp:101-107) followed by
 * Ktensor::normalize() (src/ktensor.cpp:85-99). */
void cals_oracle_denormalize_normalize(int n_modes, const int64_t *modes, int64_t R, double *factors,
                                       double *lambda) {
  double *F0 = factors;
  for (int64_t c = 0; c < R; c++)
    for (int64_t r = 0; r < modes[0]; r++)
      F0[r + c * modes[0]] *= lambda[c];
  for (int64_t c = 0; c < R; c++)
    lambda[c] = 1.0;
  int64_t off = 0;
  for (int n = 0; n < n_modes; n++) {
    double *F = factors + off;
    for (int64_t c = 0; c < R; c++) {
      double s = 0.0;
      for (int64_t r = 0; r < modes[n]; r++)
        s += F[r + c * modes[n]] * F[r + c * modes[n]];
      double nrm = sqrt(s), inv = 1.0 / nrm;
      for (int64_t r = 0; r < modes[n]; r++)
        F[r + c * modes[n]] *= inv;
      lambda[c] *= nrm;
    }
    off += modes[n] * R;
  }
}
