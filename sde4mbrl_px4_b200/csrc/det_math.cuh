// det_math.cuh — device statement of the SPEC-ARITH elementary functions
// (DESIGN.md "Arithmetic specification").  Each function is a fixed sequence of
// IEEE-754 binary32 add / mul / fma (explicit __fmaf_rn / __ffma2_rn; the file is
// compiled with -fmad=false so nothing else is contracted) and integer ops.  The
// CPU oracle states the same sequences independently (oracle/det_math.h); the
// parity tests require bit-identical results.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdempc {

__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float2 fma2_(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2_(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2_(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 splat(float a) { return make_float2(a, a); }

__device__ __forceinline__ float det_rint(float t) { return __fadd_rn(__fadd_rn(t, 12582912.0f), -12582912.0f); }

// exp(x), x <= 0 (clamped at -80)
__device__ __forceinline__ float det_exp_nonpos(float x) {
    x = x < -80.0f ? -80.0f : x;
    float n = det_rint(__fmul_rn(x, 1.44269504f));
    float r = fma_(n, -0.693359375f, x);
    r = fma_(n, 2.12194440e-4f, r);
    float p = 1.38888889e-3f;
    p = fma_(p, r, 8.33333333e-3f);
    p = fma_(p, r, 4.16666667e-2f);
    p = fma_(p, r, 1.66666667e-1f);
    p = fma_(p, r, 0.5f);
    p = fma_(p, r, 1.0f);
    p = fma_(p, r, 1.0f);
    return __uint_as_float(__float_as_uint(p) + (((uint32_t)__float2int_rz(n)) << 23));
}

// 1/d, d in [1, 2]
__device__ __forceinline__ float det_recip12(float d) {
    float r = fma_(-0.470588235f, d, 1.41176471f);
    float e = fma_(-d, r, 1.0f); r = fma_(r, e, r);
    e = fma_(-d, r, 1.0f); r = fma_(r, e, r);
    e = fma_(-d, r, 1.0f); r = fma_(r, e, r);
    return r;
}

__device__ __forceinline__ float det_tanh(float x) {
    float a = fabsf(x);
    a = a > 10.0f ? 10.0f : a;
    float e = det_exp_nonpos(__fmul_rn(-2.0f, a));
    float t = __fmul_rn(__fadd_rn(1.0f, -e), det_recip12(__fadd_rn(1.0f, e)));
    return copysignf(t, x);
}

// tanh of both halves with packed f32x2 arithmetic (same per-component sequence as det_tanh)
__device__ __forceinline__ float2 det_tanh2(float2 x) {
    float2 a = make_float2(fabsf(x.x), fabsf(x.y));
    a.x = a.x > 10.0f ? 10.0f : a.x;
    a.y = a.y > 10.0f ? 10.0f : a.y;
    float2 arg = mul2_(splat(-2.0f), a);  // >= -20: the -80 clamp of det_exp_nonpos is a no-op
    float2 t = mul2_(arg, splat(1.44269504f));
    float2 n = add2_(add2_(t, splat(12582912.0f)), splat(-12582912.0f));
    float2 r = fma2_(n, splat(-0.693359375f), arg);
    r = fma2_(n, splat(2.12194440e-4f), r);
    float2 p = splat(1.38888889e-3f);
    p = fma2_(p, r, splat(8.33333333e-3f));
    p = fma2_(p, r, splat(4.16666667e-2f));
    p = fma2_(p, r, splat(1.66666667e-1f));
    p = fma2_(p, r, splat(0.5f));
    p = fma2_(p, r, splat(1.0f));
    p = fma2_(p, r, splat(1.0f));
    float2 e;
    e.x = __uint_as_float(__float_as_uint(p.x) + (((uint32_t)__float2int_rz(n.x)) << 23));
    e.y = __uint_as_float(__float_as_uint(p.y) + (((uint32_t)__float2int_rz(n.y)) << 23));
    float2 num = add2_(splat(1.0f), make_float2(-e.x, -e.y));
    float2 d = add2_(splat(1.0f), e);
    float2 q = fma2_(splat(-0.470588235f), d, splat(1.41176471f));
    float2 nd = make_float2(-d.x, -d.y);
    float2 err = fma2_(nd, q, splat(1.0f)); q = fma2_(q, err, q);
    err = fma2_(nd, q, splat(1.0f)); q = fma2_(q, err, q);
    err = fma2_(nd, q, splat(1.0f)); q = fma2_(q, err, q);
    float2 th = mul2_(num, q);
    return make_float2(copysignf(th.x, x.x), copysignf(th.y, x.y));
}

// exp of both halves, x <= 0 (same per-component sequence as det_exp_nonpos)
__device__ __forceinline__ float2 det_exp_nonpos2(float2 x) {
    x.x = x.x < -80.0f ? -80.0f : x.x;
    x.y = x.y < -80.0f ? -80.0f : x.y;
    float2 t = mul2_(x, splat(1.44269504f));
    float2 n = add2_(add2_(t, splat(12582912.0f)), splat(-12582912.0f));
    float2 r = fma2_(n, splat(-0.693359375f), x);
    r = fma2_(n, splat(2.12194440e-4f), r);
    float2 p = splat(1.38888889e-3f);
    p = fma2_(p, r, splat(8.33333333e-3f));
    p = fma2_(p, r, splat(4.16666667e-2f));
    p = fma2_(p, r, splat(1.66666667e-1f));
    p = fma2_(p, r, splat(0.5f));
    p = fma2_(p, r, splat(1.0f));
    p = fma2_(p, r, splat(1.0f));
    float2 e;
    e.x = __uint_as_float(__float_as_uint(p.x) + (((uint32_t)__float2int_rz(n.x)) << 23));
    e.y = __uint_as_float(__float_as_uint(p.y) + (((uint32_t)__float2int_rz(n.y)) << 23));
    return e;
}

// softplus and (optionally) sigmoid of both halves; per component the same sequence as det_softplus_sigmoid
__device__ __forceinline__ void det_softplus_sigmoid2(float2 s, float2& sp, float2& sg, bool want_sg) {
    const float2 e = det_exp_nonpos2(make_float2(-fabsf(s.x), -fabsf(s.y)));
    const float2 d = add2_(splat(1.0f), e);
    const bool bx = d.x > 1.41421356f, by = d.y > 1.41421356f;
    float2 m;
    m.x = bx ? __fmul_rn(d.x, 0.5f) : d.x;
    m.y = by ? __fmul_rn(d.y, 0.5f) : d.y;
    const float2 f = add2_(m, splat(-1.0f));
    float2 p = splat(7.0376836292e-2f);
    p = fma2_(p, f, splat(-1.1514610310e-1f));
    p = fma2_(p, f, splat(1.1676998740e-1f));
    p = fma2_(p, f, splat(-1.2420140846e-1f));
    p = fma2_(p, f, splat(1.4249322787e-1f));
    p = fma2_(p, f, splat(-1.6668057665e-1f));
    p = fma2_(p, f, splat(2.0000714765e-1f));
    p = fma2_(p, f, splat(-2.4999993993e-1f));
    p = fma2_(p, f, splat(3.3333331174e-1f));
    const float2 z = mul2_(f, f);
    float2 y = mul2_(mul2_(f, z), p);
    y = fma2_(splat(-0.5f), z, y);
    y = add2_(f, y);
    y.x = bx ? __fadd_rn(y.x, 0.693147181f) : y.x;
    y.y = by ? __fadd_rn(y.y, 0.693147181f) : y.y;
    sp = add2_(make_float2(s.x > 0.0f ? s.x : 0.0f, s.y > 0.0f ? s.y : 0.0f), y);
    if (want_sg) {
        float2 q = fma2_(splat(-0.470588235f), d, splat(1.41176471f));
        const float2 nd = make_float2(-d.x, -d.y);
        float2 err = fma2_(nd, q, splat(1.0f)); q = fma2_(q, err, q);
        err = fma2_(nd, q, splat(1.0f)); q = fma2_(q, err, q);
        err = fma2_(nd, q, splat(1.0f)); q = fma2_(q, err, q);
        const float2 eq = mul2_(e, q);
        sg.x = s.x >= 0.0f ? q.x : eq.x;
        sg.y = s.y >= 0.0f ? q.y : eq.y;
    } else {
        sg = splat(0.f);
    }
}

// cephes logf kernel on f = m - 1, m in [sqrt(1/2), sqrt(2)]
__device__ __forceinline__ float det_log_kernel(float f) {
    float p = 7.0376836292e-2f;
    p = fma_(p, f, -1.1514610310e-1f);
    p = fma_(p, f, 1.1676998740e-1f);
    p = fma_(p, f, -1.2420140846e-1f);
    p = fma_(p, f, 1.4249322787e-1f);
    p = fma_(p, f, -1.6668057665e-1f);
    p = fma_(p, f, 2.0000714765e-1f);
    p = fma_(p, f, -2.4999993993e-1f);
    p = fma_(p, f, 3.3333331174e-1f);
    float z = __fmul_rn(f, f);
    float y = __fmul_rn(__fmul_rn(f, z), p);
    y = fma_(-0.5f, z, y);
    return __fadd_rn(f, y);
}

__device__ __forceinline__ float det_log1p01(float e) {
    float d = __fadd_rn(1.0f, e);
    bool big = d > 1.41421356f;
    float m = big ? __fmul_rn(d, 0.5f) : d;
    float y = det_log_kernel(__fadd_rn(m, -1.0f));
    return big ? __fadd_rn(y, 0.693147181f) : y;
}

// softplus(s) and sigmoid(s) share exp(-|s|)
__device__ __forceinline__ void det_softplus_sigmoid(float s, float& sp, float& sg) {
    float e = det_exp_nonpos(-fabsf(s));
    sp = __fadd_rn(s > 0.0f ? s : 0.0f, det_log1p01(e));
    float r = det_recip12(__fadd_rn(1.0f, e));
    sg = s >= 0.0f ? r : __fmul_rn(e, r);
}

// same sequences; the sigmoid is skipped when not wanted (sg = 0)
__device__ __forceinline__ void det_softplus_sigmoid_opt(float s, float& sp, float& sg, bool want_sg) {
    float e = det_exp_nonpos(-fabsf(s));
    sp = __fadd_rn(s > 0.0f ? s : 0.0f, det_log1p01(e));
    sg = 0.f;
    if (want_sg) {
        float r = det_recip12(__fadd_rn(1.0f, e));
        sg = s >= 0.0f ? r : __fmul_rn(e, r);
    }
}

__device__ __forceinline__ float det_log(float u) {
    uint32_t b = __float_as_uint(u);
    int ex = (int)((b >> 23) & 0xffu) - 127;
    float m = __uint_as_float((b & 0x007fffffu) | 0x3f800000u);
    if (m > 1.41421356f) { m = __fmul_rn(m, 0.5f); ex += 1; }
    float y = det_log_kernel(__fadd_rn(m, -1.0f));
    float ef = (float)ex;
    y = fma_(ef, -2.12194440e-4f, y);
    return fma_(ef, 0.693359375f, y);
}

__device__ __forceinline__ void det_sincos2pi(float u, float& s_out, float& c_out) {
    float q = __fmul_rn(u, 4.0f);
    float kf = det_rint(q);
    float a = __fmul_rn(__fadd_rn(q, -kf), 1.57079633f);
    float z = __fmul_rn(a, a);
    float sp = -1.9515295891e-4f;
    sp = fma_(sp, z, 8.3321608736e-3f);
    sp = fma_(sp, z, -1.6666654611e-1f);
    float s = fma_(__fmul_rn(a, z), sp, a);
    float cp = 2.443315711809948e-5f;
    cp = fma_(cp, z, -1.388731625493765e-3f);
    cp = fma_(cp, z, 4.166664568298827e-2f);
    float c = fma_(__fmul_rn(z, z), cp, fma_(-0.5f, z, 1.0f));
    int k = __float2int_rz(kf) & 3;
    s_out = (k == 0) ? s : (k == 1) ? c : (k == 2) ? -s : -c;
    c_out = (k == 0) ? c : (k == 1) ? -s : (k == 2) ? -c : s;
}

__device__ __forceinline__ float det_rsqrt_near1(float n2) {
    float r = fma_(-0.5f, __fadd_rn(n2, -1.0f), 1.0f);
    float h = __fmul_rn(0.5f, n2);
    float t = fma_(-h, __fmul_rn(r, r), 1.5f); r = __fmul_rn(r, t);
    t = fma_(-h, __fmul_rn(r, r), 1.5f); r = __fmul_rn(r, t);
    t = fma_(-h, __fmul_rn(r, r), 1.5f); r = __fmul_rn(r, t);
    return r;
}

// Philox4x32-10
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    float u1 = __fmul_rn(__fadd_rn((float)(a >> 8), 0.5f), 5.9604644775390625e-08f);
    float u2 = __fmul_rn(__fadd_rn((float)(b >> 8), 0.5f), 5.9604644775390625e-08f);
    float rad = __fsqrt_rn(__fmul_rn(-2.0f, det_log(u1)));
    float sn, cs;
    det_sincos2pi(u2, sn, cs);
    n0 = __fmul_rn(rad, cs);
    n1 = __fmul_rn(rad, sn);
}

}  // namespace sdempc
