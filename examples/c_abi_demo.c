/* c_abi_demo -- a plain-C caller of the drop-in boundary (include/cals_b200.h): no C++, no Python.
 * Fits three CP models of ranks 2, 3, 4 to a random 30 x 20 x 10 tensor for 25 ALS iterations each and prints their fits.
 *   gcc -std=c11 -O2 -Iinclude examples/c_abi_demo.c -Lcp-cals_b200 -lcals_b200 -Wl,-rpath,$PWD/cp-cals_b200 -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "cals_b200.h"

static double uniform(unsigned long long *state) { /* xorshift64*, uniform in (-1, 1) */
  *state ^= *state >> 12;
  *state ^= *state << 25;
  *state ^= *state >> 27;
  return (double)((*state * 2685821657736338717ull) >> 11) / 9007199254740992.0 * 2.0 - 1.0;
}

#define CHECK(call)                                                                                                    \
  do {                                                                                                                 \
    if ((call) != 0) {                                                                                                 \
      fprintf(stderr, "%s failed: %s\n", #call, cals_b200_last_error(ctx));                                            \
      return 1;                                                                                                        \
    }                                                                                                                  \
  } while (0)

int main(void) {
  const uint64_t modes[3] = {30, 20, 10};
  const uint64_t ranks[3] = {2, 3, 4};
  unsigned long long seed = 88172645463325252ull;
  cals_b200_ctx *ctx = NULL;
  if (cals_b200_create(&ctx, 0) != 0) {
    fprintf(stderr, "cals_b200_create: %s\n", cals_b200_last_error(NULL));
    return 2; /* no B200: there is no CPU fallback */
  }
  printf("%s\n", cals_b200_version());

  const size_t nX = (size_t)(modes[0] * modes[1] * modes[2]);
  double *X = malloc(nX * sizeof *X);
  for (size_t i = 0; i < nX; i++)
    X[i] = uniform(&seed);
  CHECK(cals_b200_set_tensor(ctx, 3, modes, X));
  CHECK(cals_b200_configure(ctx, /*buffer_cols=*/9, /*max_iterations=*/25, /*tol=*/1e-7, CALS_B200_FORCE_MAX_ITER));

  double *factors[3][3];
  for (int m = 0; m < 3; m++) {
    const double *in[3];
    for (int n = 0; n < 3; n++) {
      const size_t cnt = (size_t)(modes[n] * ranks[m]);
      factors[m][n] = malloc(cnt * sizeof(double));
      for (size_t i = 0; i < cnt; i++)
        factors[m][n][i] = uniform(&seed);
      in[n] = factors[m][n];
    }
    int id = -1;
    CHECK(cals_b200_enqueue_model(ctx, ranks[m], in, /*jk_mode=*/-1, 0, &id));
  }

  cals_b200_report rep;
  CHECK(cals_b200_run(ctx, &rep));
  printf("||X|| = %.6f, %llu CALS iterations, %llu kernel launches, device loop %.3f ms\n", rep.x_norm,
         (unsigned long long)rep.iter, (unsigned long long)rep.kernel_launches, rep.device_ms);
  int bad = 0;
  for (int m = 0; m < 3; m++) {
    double lambda[4];
    cals_b200_model_stats st;
    double *out[3] = {factors[m][0], factors[m][1], factors[m][2]};
    CHECK(cals_b200_fetch_model(ctx, m, out, lambda, &st));
    printf("model %d: rank %llu, %llu iterations, error %.6f, fit %.6f\n", m, (unsigned long long)ranks[m],
           (unsigned long long)st.iters, st.error, st.fit);
    if (st.iters != 25 || !(st.fit > 0.0 && st.fit < 1.0) || fabs(st.fit - (1.0 - st.error / rep.x_norm)) > 1e-12)
      bad = 1;
  }
  cals_b200_destroy(ctx);
  printf(bad ? "UNEXPECTED RESULT\n" : "OK\n");
  return bad;
}
