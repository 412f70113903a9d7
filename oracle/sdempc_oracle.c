/*
 * sdempc_oracle.c — builds the float32 and float64 instances of the CPU
 * restatement (see sdempc_oracle_impl.h for the contract and the
 * "parity unpinned" statement).  TEST INFRASTRUCTURE ONLY.
 *
 *   gcc -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp -shared -fPIC sdempc_oracle.c -o _build/libsdempc_oracle.so -lm
 *
 * -ffp-contract=off and no -ffast-math are REQUIRED: the float32 instance is the
 * bit-exact arithmetic specification the CUDA kernels reproduce (SPEC-ARITH).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/sdempc.h"
#include "det_math.h"

#define REAL float
#define SUFFIX f32
#include "sdempc_oracle_impl.h"
#undef REAL
#undef SUFFIX

#define REAL double
#define SUFFIX f64
#define REAL_IS_DOUBLE 1
#include "sdempc_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef REAL_IS_DOUBLE

#ifdef _OPENMP
#include <omp.h>
/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline must use the host cores it is given */
void oracle_set_threads(int n) { omp_set_num_threads(n); }
int oracle_get_max_threads(void) { return omp_get_max_threads(); }
#else
void oracle_set_threads(int n) { (void)n; }
int oracle_get_max_threads(void) { return 1; }
#endif

const char* oracle_version(void) { return "sdempc-oracle 1 (parity unpinned: restates SURVEY.md section 8a [SPEC])"; }
