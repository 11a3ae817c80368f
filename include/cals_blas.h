// Thread knobs of the reference's BLAS shim (reference include/cals_blas.h:184-186).  The B200 path has no host
// BLAS: the value is only recorded (CalsReport::n_threads) so that callers written against the reference compile
// and behave the same.
#ifndef CALS_B200_CALS_BLAS_H
#define CALS_B200_CALS_BLAS_H

void set_threads(int threads);
int get_threads();

#endif
