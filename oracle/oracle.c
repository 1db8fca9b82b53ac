/* oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * C restatement of the arithmetic details of the reference's elementwise path
 * that NumPy cannot express directly.  The reference compiles every kernel with
 * `-ftz=true` (cupy/cuda/compiler.py:667) and default nvcc/NVRTC FMA contraction,
 * so `z = a*x + y` (tests/cupy_tests/core_tests/test_userkernel.py style kernels,
 * BASELINE.json config 2) executes as ONE fused multiply-add with denormals
 * flushed to zero (SASS `FFMA.FTZ`, SURVEY.md compile probe 2).  fmaf() is the
 * correctly rounded single-rounding FMA; the flush is restated explicitly.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg call this.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static float ftz(float v) {
    uint32_t u;
    memcpy(&u, &v, 4);
    if ((u & 0x7f800000u) == 0) u &= 0x80000000u;   /* denormal (or zero) -> signed zero */
    memcpy(&v, &u, 4);
    return v;
}

/* z[i] = fma(a, x[i], y[i]) in float32, flush-to-zero on inputs and result */
void oracle_axpy_f32(float a, const float* x, const float* y, float* z, size_t n) {
    a = ftz(a);
    for (size_t i = 0; i < n; ++i) z[i] = ftz(fmaf(a, ftz(x[i]), ftz(y[i])));
}

/* same without the flush (inputs known to be normal): used for timing the CPU baseline */
void oracle_axpy_f32_noftz(float a, const float* x, const float* y, float* z, size_t n) {
    for (size_t i = 0; i < n; ++i) z[i] = fmaf(a, x[i], y[i]);
}

void oracle_fma_f64(const double* a, const double* b, const double* c, double* z, size_t n) {
    for (size_t i = 0; i < n; ++i) z[i] = fma(a[i], b[i], c[i]);
}

/* elementwise float32 add / multiply with FTZ (IEEE-exact otherwise) */
void oracle_add_f32(const float* x, const float* y, float* z, size_t n) {
    for (size_t i = 0; i < n; ++i) z[i] = ftz(ftz(x[i]) + ftz(y[i]));
}
void oracle_mul_f32(const float* x, const float* y, float* z, size_t n) {
    for (size_t i = 0; i < n; ++i) z[i] = ftz(ftz(x[i]) * ftz(y[i]));
}

/* inclusive scan, int64 (wraps like the device) -- cupy/_core/_routines_math.pyx:702-751 */
void oracle_cumsum_i64(const int64_t* x, int64_t* y, size_t n) {
    uint64_t acc = 0;
    for (size_t i = 0; i < n; ++i) { acc += (uint64_t)x[i]; y[i] = (int64_t)acc; }
}

/* float32 sum with float64 accumulation: the tolerance anchor for float reductions */
double oracle_sum_f32(const float* x, size_t n) {
    double acc = 0.0;
    for (size_t i = 0; i < n; ++i) acc += (double)x[i];
    return acc;
}
