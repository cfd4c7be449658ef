/*
 * oracle.c -- builds liborc.so: the CPU restatement of the LiteGS render hot path in fp32 and fp64.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (litegs_b200/) never does.
 * See oracle_core.h for the per-function reference citations and the parity-pinning note.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REAL float
#define SUF(n) n##_f32
#define EXP expf
#define SQRT sqrtf
#define FABS fabsf
#define FMIN fminf
#define FMAX fmaxf
#define CEIL ceilf
#define FLOOR floorf
#include "oracle_core.h"
#undef REAL
#undef SUF
#undef EXP
#undef SQRT
#undef FABS
#undef FMIN
#undef FMAX
#undef CEIL
#undef FLOOR

#define REAL double
#define SUF(n) n##_f64
#define EXP exp
#define SQRT sqrt
#define FABS fabs
#define FMIN fmin
#define FMAX fmax
#define CEIL ceil
#define FLOOR floor
#include "oracle_core.h"

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
