// Host-side check of the stream-K bookkeeping shared by mttkrp_dmma_kernel and mttkrp_reduce_kernel
// (cp-cals_b200/csrc/mttkrp.cuh, struct StreamK).  Compiled with nvcc, runs on the CPU (no kernel launch).
#include <cstdio>
#include <vector>

#include "../../cp-cals_b200/csrc/mttkrp.cuh"

using calsb200::StreamK;

int main() {
  long long checked = 0;
  for (int G : {1, 2, 7, 148, 296})
    for (int pairs : {1, 2, 3, 5, 9, 45, 125, 400})
      for (int Tp : {1, 2, 3, 13, 25, 125, 1000, 8000}) {
        StreamK sk = StreamK::make(pairs, Tp, G);
        // 1. ranges tile [0,total) in order
        long long prev = 0;
        for (int b = 0; b < G; b++) {
          if (sk.lo(b) != prev || sk.hi(b) < sk.lo(b) || (b < sk.G && sk.hi(b) == sk.lo(b))) {
            printf("FAIL range G=%d pairs=%d Tp=%d b=%d\n", G, pairs, Tp, b);
            return 1;
          }
          prev = sk.hi(b);
        }
        if (prev != sk.total) {
          printf("FAIL cover\n");
          return 1;
        }
        // 2. owner(x) is the CTA whose range holds x (sampled)
        long long step = sk.total > 5000 ? sk.total / 997 : 1;
        for (long long x = 0; x < sk.total; x += step) {
          int b = sk.owner(x);
          if (b < 0 || b >= sk.G || x < sk.lo(b) || x >= sk.hi(b)) {
            printf("FAIL owner G=%d pairs=%d Tp=%d x=%lld -> %d\n", G, pairs, Tp, x, b);
            return 1;
          }
        }
        // 3. slots: every CTA touching a pair gets a distinct slot < kmax; total slots within the allocation bound
        for (int p = 0; p < pairs; p++) {
          int f = sk.first_cta(p), l = sk.last_cta(p);
          if (l - f + 1 > sk.kmax || l < f) {
            printf("FAIL kmax G=%d pairs=%d Tp=%d pair=%d first=%d last=%d kmax=%d\n", G, pairs, Tp, p, f, l, sk.kmax);
            return 1;
          }
          for (int b = f; b <= l; b++) { // each of them really has a chunk of this pair
            long long lo = sk.lo(b) > (long long)p * Tp ? sk.lo(b) : (long long)p * Tp;
            long long hi = sk.hi(b) < (long long)(p + 1) * Tp ? sk.hi(b) : (long long)(p + 1) * Tp;
            if (lo >= hi) {
              printf("FAIL empty slot G=%d pairs=%d Tp=%d pair=%d b=%d\n", G, pairs, Tp, p, b);
              return 1;
            }
          }
        }
        if ((long long)pairs * sk.kmax > 2LL * G + 2LL * pairs + 8) {
          printf("FAIL bound G=%d pairs=%d Tp=%d kmax=%d\n", G, pairs, Tp, sk.kmax);
          return 1;
        }
        checked++;
      }
  printf("OK %lld configurations\n", checked);
  return 0;
}
