/*
 * det_math.h — the deterministic float32 elementary functions of SPEC-ARITH
 * (DESIGN.md "Arithmetic specification"), CPU statement.  TEST INFRASTRUCTURE.
 *
 * Every function is a fixed sequence of IEEE-754 binary32 add / mul / fma and
 * integer operations, so the CPU oracle and the CUDA kernels (which restate the
 * same sequences with __fmaf_rn, compiled with -fmad=false) produce
 * bit-identical results.  Compile with -ffp-contract=off and without
 * -ffast-math.
 */
#ifndef ORACLE_DET_MATH_H_
#define ORACLE_DET_MATH_H_
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint32_t det_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float det_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* round to nearest even integer, |t| < 2^22 */
static inline float det_rint(float t) { return (t + 12582912.0f) - 12582912.0f; }

/* exp(x) for x <= 0 (clamped at -80) : Cody-Waite reduction + degree-6 Taylor, |rel err| ~ 2e-7 */
static inline float det_exp_nonpos(float x) {
    x = x < -80.0f ? -80.0f : x;
    float n = det_rint(x * 1.44269504f);
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.38888889e-3f;
    p = fmaf(p, r, 8.33333333e-3f);
    p = fmaf(p, r, 4.16666667e-2f);
    p = fmaf(p, r, 1.66666667e-1f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    return det_float(det_bits(p) + ((uint32_t)((int32_t)n) << 23));
}

/* 1/d for d in [1, 2]: linear seed + 3 Newton steps */
static inline float det_recip12(float d) {
    float r = fmaf(-0.470588235f, d, 1.41176471f);
    float e = fmaf(-d, r, 1.0f); r = fmaf(r, e, r);
    e = fmaf(-d, r, 1.0f); r = fmaf(r, e, r);
    e = fmaf(-d, r, 1.0f); r = fmaf(r, e, r);
    return r;
}

static inline float det_tanh(float x) {
    float a = fabsf(x);
    a = a > 10.0f ? 10.0f : a;
    float e = det_exp_nonpos(-2.0f * a);
    float t = (1.0f - e) * det_recip12(1.0f + e);
    return copysignf(t, x);
}

/* cephes logf kernel: log(m) for m in [sqrt(1/2), sqrt(2)], f = m - 1 */
static inline float det_log_kernel(float f) {
    float p = 7.0376836292e-2f;
    p = fmaf(p, f, -1.1514610310e-1f);
    p = fmaf(p, f, 1.1676998740e-1f);
    p = fmaf(p, f, -1.2420140846e-1f);
    p = fmaf(p, f, 1.4249322787e-1f);
    p = fmaf(p, f, -1.6668057665e-1f);
    p = fmaf(p, f, 2.0000714765e-1f);
    p = fmaf(p, f, -2.4999993993e-1f);
    p = fmaf(p, f, 3.3333331174e-1f);
    float z = f * f;
    float y = (f * z) * p;
    y = fmaf(-0.5f, z, y);
    return f + y;
}

/* log(1 + e) for e in [0, 1] */
static inline float det_log1p01(float e) {
    float d = 1.0f + e;
    int big = d > 1.41421356f;
    float m = big ? d * 0.5f : d;
    float y = det_log_kernel(m - 1.0f);
    return big ? y + 0.693147181f : y;
}

static inline float det_softplus(float s) {
    float e = det_exp_nonpos(-fabsf(s));
    return (s > 0.0f ? s : 0.0f) + det_log1p01(e);
}

static inline float det_sigmoid(float s) {
    float e = det_exp_nonpos(-fabsf(s));
    float r = det_recip12(1.0f + e);
    return s >= 0.0f ? r : e * r;
}

/* log(u) for normal u in (0, 1) */
static inline float det_log(float u) {
    uint32_t b = det_bits(u);
    int32_t ex = (int32_t)((b >> 23) & 0xffu) - 127;
    float m = det_float((b & 0x007fffffu) | 0x3f800000u);
    if (m > 1.41421356f) { m = m * 0.5f; ex += 1; }
    float y = det_log_kernel(m - 1.0f);
    float ef = (float)ex;
    y = fmaf(ef, -2.12194440e-4f, y);
    return fmaf(ef, 0.693359375f, y);
}

/* sin and cos of 2*pi*u, u in [0, 1) */
static inline void det_sincos2pi(float u, float* s_out, float* c_out) {
    float q = u * 4.0f;
    float kf = det_rint(q);
    float a = (q - kf) * 1.57079633f;
    float z = a * a;
    float sp = -1.9515295891e-4f;
    sp = fmaf(sp, z, 8.3321608736e-3f);
    sp = fmaf(sp, z, -1.6666654611e-1f);
    float s = fmaf(a * z, sp, a);
    float cp = 2.443315711809948e-5f;
    cp = fmaf(cp, z, -1.388731625493765e-3f);
    cp = fmaf(cp, z, 4.166664568298827e-2f);
    float c = fmaf(z * z, cp, fmaf(-0.5f, z, 1.0f));
    int k = ((int)kf) & 3;
    *s_out = (k == 0) ? s : (k == 1) ? c : (k == 2) ? -s : -c;
    *c_out = (k == 0) ? c : (k == 1) ? -s : (k == 2) ? -c : s;
}

/* 1/sqrt(n2) for n2 near 1 (quaternion renormalisation inside the EM step) */
static inline float det_rsqrt_near1(float n2) {
    float r = fmaf(-0.5f, n2 - 1.0f, 1.0f);
    float h = 0.5f * n2;
    float t = fmaf(-h, r * r, 1.5f); r = r * t;
    t = fmaf(-h, r * r, 1.5f); r = r * t;
    t = fmaf(-h, r * r, 1.5f); r = r * t;
    return r;
}

#endif
