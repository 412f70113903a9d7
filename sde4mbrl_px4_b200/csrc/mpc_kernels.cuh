// mpc_kernels.cuh — sm_100a kernels of the neural-SDE MPC solve.
//
// Replaces, for the path of BASELINE.json's north_star, the arithmetic behind the
// reference's m_mpc(x, rng, opt_state, curr_t=, xdes=) call
// (/root/reference sde4mbrl_px4/mpc_controller/sde_control.py:400-416):
//   (a) Euler-Maruyama particle rollout of the drift/diffusion MLPs (R6),
//   (b) its reverse-mode adjoint wrt the control sequence fused with the tracking
//       cost and the uncertainty penalty (R7, R8),
//   (c) the accelerated projected-gradient loop with Armijo line search (R9),
// all inside ONE launch per batch of problems: no host round trip per iteration.
//
// Mapping (DESIGN.md "Kernel design"):
//   * one warp integrates one particle of one problem; lane j owns hidden unit j
//     (and j+32 when W = 64) of BOTH networks, packed as float2 = (drift, diffusion)
//     so every hidden-layer multiply-add is one FFMA2 (fma.rn.f32x2);
//   * the 13-state, its adjoint and the rigid-body algebra are replicated in the
//     registers of all 32 lanes (no communication on the dependent chain);
//   * hidden activations cross lanes through a 256-byte shared-memory line read
//     back with broadcast LDS.128; transposed weights for the adjoint are staged
//     once per CTA into shared memory with a TMA bulk copy (cp.async.bulk);
//   * a team of P warps (P = particles) shares one problem; the APG state is
//     replicated per warp so the only cross-warp traffic is the particle mean.
// Arithmetic order follows SPEC-ARITH exactly (bit-identical to the CPU oracle).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sdempc.h"
#include "det_math.cuh"

namespace sdempc {

constexpr int NX = SDEMPC_NX;

// Everything a kernel needs, passed by value as a __grid_constant__ parameter.
struct KParams {
    // solver configuration
    int H, P, max_iter, max_no_improve, maxls, reset_option;
    unsigned flags;
    float dt[SDEMPC_MAX_H], sdt[SDEMPC_MAX_H];
    float discount;
    float u_lo[SDEMPC_MAX_NU], u_hi[SDEMPC_MAX_NU], uref[SDEMPC_MAX_NU];
    float uerr, perr[3], verr[3], qerr[3], werr[3], res_mult, slew;
    float slewc, slew_lo[SDEMPC_MAX_NU], slew_hi[SDEMPC_MAX_NU];   // soft rate constraint (0: off)
    float rate_w[SDEMPC_MAX_H];                                    // slewc * discount^t (float products, host)
    float init_step, max_step, coef, dec_f, inc_f, atol, rtol;
    const float* beta_tab;                                         // momentum table [max_iter + 2] (device)
    int beta_adaptive;                                             // moment_scale given
    // rigid-body model
    float inv_m, grav, kT, kT2, J[3], Jinv[3], Jd[3], mixer[3][SDEMPC_MAX_NU], sig0[6];
    // device data
    const float* wimg;   // packed weight image (Layout)
    const float* traj;   // [T][14] internal frame, or nullptr
    int T;
    float2* mtape_g;     // global MLP tape scratch (W = 64), per resident warp
    // per-warp shared-memory layout (float offsets), computed on the host
    int ws_stride, o_xk, o_yk, o_g, o_xp, o_uprev, o_xref, o_xi, o_xtape, o_stape, o_mtape, o_bufA, o_bufB,
        o_act3, o_lz, o_red, o_zb, o_lob, o_g2;
    int gx_stride;       // group kernel: floats of per-warp exchange buffers (bufA | bufB | act3)
    int team_stride;     // per-team scratch (P > 1): floats
    // batch I/O (device pointers)
    int B;
    const float* x;
    const float* curr_t;
    const float* xdes;
    const float* xref_win;
    const unsigned long long* rng;
    const float* xi_override;
    float* u_plan;       // in  [B][H][nu] (staged plan; never written, so a launch can be repeated)
    sdempc_info* info;   // in  [B] (carried step size)
    float* u_plan_out;   // out [B][H][nu]
    sdempc_info* info_out;
    float* x_evol;       // out [B][H+1][13]
    float* trace;        // nullptr or [B][max_iter][8]
    // rollout-only mode (sdempc_rollout)
    const float* u_in;   // [B][H][nu]
    const float* uprev_in;  // [B][nu]
    float* cost_out;     // [B]
    float* grad_out;     // [B][H][nu] or nullptr
    // tensor-core solve (mpc_tcsolve.cuh): per-CTA global workspace, problems per CTA, row stride of the workspace arrays
    float* tcs_ws;
    int tcs_ppc, tcs_rs, tcs_sms;
    // closed loop
    int ticks;
    const float* t0;
    float* x_hist;
    float* u_hist;
    float* stats;
};

// Packed weight image.  Part A is always staged in shared memory (transposed /
// row copies used by the adjoint and by the 12-lane output layer); part B holds
// the forward rows, kept in registers when W == 32 and in shared memory otherwise.
// Row strides are multiples of 4 floats congruent to 4 mod 8, which makes
// per-lane-row LDS.128 conflict free.
template <int NU, int W>
struct Layout {
    static constexpr int NIN = 6 + NU;
    static constexpr int UPL = W / 32;
    static constexpr bool WREG = (W == 32);
    static constexpr int PAIR_STRIDE = 2 * W + 4;                    // W pairs per row
    static constexpr int W2T = 0;                                     // [W rows k][W pairs j]
    static constexpr int W1T = W2T + W * PAIR_STRIDE;                 // [NIN rows i][W pairs j]
    static constexpr int W3R_STRIDE = W + 4;
    static constexpr int W3R = W1T + NIN * PAIR_STRIDE;               // [12 rows][W]
    static constexpr int B3 = W3R + 12 * W3R_STRIDE;                  // [12] + pad
    static constexpr int PART_A = B3 + 16;
    static constexpr int W1P_STRIDE = (2 * NIN) % 8 == 4 ? 2 * NIN : 2 * NIN + 4;
    static constexpr int W1P = PART_A;                                // [W rows j][NIN pairs k]
    static constexpr int W2P = W1P + W * W1P_STRIDE;                  // [W rows j][W pairs k]
    static constexpr int W3C = W2P + W * PAIR_STRIDE;                 // [W rows j][6 pairs o]
    static constexpr int B1 = W3C + W * 12;                           // [W] pairs
    static constexpr int B2 = B1 + 2 * W;                             // [W] pairs
    static constexpr int TOTAL = B2 + 2 * W;
    static constexpr int SMEM_FLOATS = WREG ? PART_A : TOTAL;
};

#ifdef SDEMPC_NO_RATE   // experiment builds only: compile the soft rate constraint out
#define SDEMPC_RATE_ON(P) false
#else
#define SDEMPC_RATE_ON(P) ((P).slewc != 0.f)
#endif
#ifndef SDEMPC_L3_OFF
#define SDEMPC_L3_OFF 16
#endif
#ifndef SDEMPC_LZ_OFF
#define SDEMPC_LZ_OFF 16
#endif
constexpr int L3_OFF = SDEMPC_L3_OFF, LZ_OFF = SDEMPC_LZ_OFF;   // first lane of the second problem in the split layers

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 lds2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 xy(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 zw(float4 v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float clipf(float v, float lo, float hi) { v = v < lo ? lo : v; return v > hi ? hi : v; }

// momentum of the k-th consecutive accepted step ([SPEC] "APG" step 4; include/sdempc.h `moment_scale`).  Two equivalent
// forms: a table built on the host with the oracle's float operations (k / (k + 3), or beta_init / mu^(k-1) capped at 1;
// always present), and the classical rule computed in place.  Which one a kernel uses is a code-generation choice: both
// register-bound solve kernels are sensitive to it (measured on the same build: the throughput kernel 37.5 ms with the
// in-place form / 39.9 ms with the table; the latency kernel 7.18 ms in place / 6.64 ms with the table).
__device__ __forceinline__ float apg_momentum_tab(const KParams& P, int k) { return __ldg(P.beta_tab + k); }
__device__ __forceinline__ float apg_momentum(const KParams& P, int k) {
    if (!P.beta_adaptive) return __fdiv_rn((float)k, (float)(k + 3));
    return __ldg(P.beta_tab + k);
}

// warp-shaped sum of SPEC-ARITH: per-lane strided partial, then xor butterfly 16,8,4,2,1
__device__ __forceinline__ float warp_butterfly(float p) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) p = p + __shfl_xor_sync(0xffffffffu, p, off);
    return p;
}

// enu2ned (involution): see oracle enu_ned / SURVEY 8a "State / frames"
__device__ __forceinline__ void enu_ned(const float* x, float* o) {
    const float s = 0.70710678118654752440f;
    float c0 = s * (x[6] + x[9]), c1 = s * (x[7] + x[8]), c2 = s * (x[7] - x[8]), c3 = s * (x[6] - x[9]);
    if (c0 < 0.f) { c0 = -c0; c1 = -c1; c2 = -c2; c3 = -c3; }
    float t0 = x[0], t3 = x[3];
    o[0] = x[1]; o[1] = t0; o[2] = -x[2];
    o[3] = x[4]; o[4] = t3; o[5] = -x[5];
    o[6] = c0; o[7] = c1; o[8] = c2; o[9] = c3;
    o[10] = x[10]; o[11] = -x[11]; o[12] = -x[12];
}

__device__ __forceinline__ void quat_renorm(float* q) {
    const float n2 = fma_(q[3], q[3], fma_(q[2], q[2], fma_(q[1], q[1], q[0] * q[0])));
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(n2));
    q[0] = q[0] * inv; q[1] = q[1] * inv; q[2] = q[2] * inv; q[3] = q[3] * inv;
}

// state_from_traj on the internal-frame table (clamped linear interpolation)
__device__ __forceinline__ void traj_interp(const float* __restrict__ tab, int T, float t, float* out) {
    int lo = 0, hi = T - 1;
    if (t <= __ldg(tab)) {
#pragma unroll
        for (int i = 0; i < NX; ++i) out[i] = __ldg(tab + 1 + i);
    } else if (t >= __ldg(tab + (size_t)hi * 14)) {
#pragma unroll
        for (int i = 0; i < NX; ++i) out[i] = __ldg(tab + (size_t)hi * 14 + 1 + i);
    } else {
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (__ldg(tab + (size_t)mid * 14) <= t) lo = mid; else hi = mid;
        }
        const float* a = tab + (size_t)lo * 14;
        const float* b = a + 14;
        const float ta = __ldg(a), tb = __ldg(b);
        const float al = __fdiv_rn(t - ta, tb - ta);
#pragma unroll
        for (int i = 0; i < NX; ++i) { const float av = __ldg(a + 1 + i); out[i] = fma_(al, __ldg(b + 1 + i) - av, av); }
    }
    quat_renorm(out + 6);
}

__device__ __forceinline__ void rotmat(const float* q, float (&R)[3][3]) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float xx = x * x, yy = y * y, zz = z * z, xy_ = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
    R[0][0] = fma_(-2.f, yy + zz, 1.f); R[0][1] = 2.f * (xy_ - wz); R[0][2] = 2.f * (xz + wy);
    R[1][0] = 2.f * (xy_ + wz); R[1][1] = fma_(-2.f, xx + zz, 1.f); R[1][2] = 2.f * (yz - wx);
    R[2][0] = 2.f * (xz - wy); R[2][1] = 2.f * (yz + wx); R[2][2] = fma_(-2.f, xx + yy, 1.f);
}

__device__ __forceinline__ void quat_err(const float* rq, const float* qq, float (&e)[3]) {
    e[0] = fma_(rq[3], qq[2], fma_(-rq[2], qq[3], fma_(-rq[1], qq[0], rq[0] * qq[1])));
    e[1] = fma_(-rq[3], qq[1], fma_(-rq[2], qq[0], fma_(rq[1], qq[3], rq[0] * qq[2])));
    e[2] = fma_(-rq[3], qq[0], fma_(rq[2], qq[1], fma_(-rq[1], qq[2], rq[0] * qq[3])));
}

// Per-warp view of shared memory + register-resident forward weights.
template <int NU, int W>
struct Warp {
    using L = Layout<NU, W>;
    static constexpr int NIN = L::NIN, UPL = L::UPL;
    int lane;
    const float* ws;   // CTA weight image in shared memory
    float *xk, *yk, *g, *g2, *xp, *uprev, *xref, *xi, *xtape, *stape, *bufA, *bufB, *act3, *lz, *red;
    float2* mtape;     // [H][2][W] (shared or global)
    // W == 32: forward rows in registers
    float2 w1[L::WREG ? UPL : 1][L::WREG ? NIN : 1];
    float2 w2[L::WREG ? UPL : 1][L::WREG ? W : 1];
    float2 w3c[L::WREG ? UPL : 1][L::WREG ? 6 : 1];
    float2 b1[L::WREG ? UPL : 1], b2[L::WREG ? UPL : 1];

    __device__ __forceinline__ float2 W1P(int uu, int k) const {
        if constexpr (L::WREG) return w1[uu][k];
        else return lds2(ws + L::W1P + (lane + 32 * uu) * L::W1P_STRIDE + 2 * k);
    }
    __device__ __forceinline__ float2 W3C(int uu, int o) const {
        if constexpr (L::WREG) return w3c[uu][o];
        else return lds2(ws + L::W3C + (lane + 32 * uu) * 12 + 2 * o);
    }
    __device__ __forceinline__ float2 B1(int uu) const {
        if constexpr (L::WREG) return b1[uu];
        else return lds2(ws + L::B1 + 2 * (lane + 32 * uu));
    }
    __device__ __forceinline__ float2 B2(int uu) const {
        if constexpr (L::WREG) return b2[uu];
        else return lds2(ws + L::B2 + 2 * (lane + 32 * uu));
    }
    __device__ __forceinline__ void load_regs(const float* __restrict__ wimg) {
        if constexpr (L::WREG) {
#pragma unroll
            for (int uu = 0; uu < UPL; ++uu) {
                const int j = lane + 32 * uu;
#pragma unroll
                for (int k = 0; k < NIN; ++k) w1[uu][k] = __ldg(reinterpret_cast<const float2*>(wimg + L::W1P + j * L::W1P_STRIDE) + k);
#pragma unroll
                for (int k = 0; k < W; ++k) w2[uu][k] = __ldg(reinterpret_cast<const float2*>(wimg + L::W2P + j * L::PAIR_STRIDE) + k);
#pragma unroll
                for (int o = 0; o < 6; ++o) w3c[uu][o] = __ldg(reinterpret_cast<const float2*>(wimg + L::W3C + j * 12) + o);
                b1[uu] = __ldg(reinterpret_cast<const float2*>(wimg + L::B1) + j);
                b2[uu] = __ldg(reinterpret_cast<const float2*>(wimg + L::B2) + j);
            }
        }
    }
};

template <int NU>
__device__ __forceinline__ void load_u(const float* u, int t, float (&o)[NU]) {
    if constexpr (NU % 4 == 0) {
#pragma unroll
        for (int i = 0; i < NU; i += 4) { float4 v = lds4(u + t * NU + i); o[i] = v.x; o[i + 1] = v.y; o[i + 2] = v.z; o[i + 3] = v.w; }
    } else if constexpr (NU % 2 == 0) {
#pragma unroll
        for (int i = 0; i < NU; i += 2) { float2 v = lds2(u + t * NU + i); o[i] = v.x; o[i + 1] = v.y; }
    } else {
#pragma unroll
        for (int i = 0; i < NU; ++i) o[i] = u[t * NU + i];
    }
}

__device__ __forceinline__ void load13(const float* p, float (&o)[NX]) {
    float4 a = lds4(p), b = lds4(p + 4), c = lds4(p + 8);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    o[8] = c.x; o[9] = c.y; o[10] = c.z; o[11] = c.w; o[12] = p[12];
}

__device__ __forceinline__ void store13_lane0(float* p, const float (&x)[NX], int lane) {
    if (lane == 0) {
        *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(x[4], x[5], x[6], x[7]);
        *reinterpret_cast<float4*>(p + 8) = make_float4(x[8], x[9], x[10], x[11]);
        p[12] = x[12];
    }
}

// ---------------------------------------------------------------------------------
// Register-level pieces of one step (no memory access), shared by the per-warp
// kernel (state replicated in all lanes) and the group kernel (lane = problem).
// ---------------------------------------------------------------------------------
template <int NU>
__device__ __forceinline__ void phys_features(const float (&x)[NX], const float (&u)[NU], float (&z)[6 + NU]) {
    const float* v = x + 3;
    float R[3][3];
    rotmat(x + 6, R);
#pragma unroll
    for (int i = 0; i < 3; ++i) z[i] = fma_(R[2][i], v[2], fma_(R[1][i], v[1], R[0][i] * v[0]));
#pragma unroll
    for (int i = 0; i < 3; ++i) z[3 + i] = x[10 + i];
#pragma unroll
    for (int i = 0; i < NU; ++i) z[6 + i] = u[i];
}

// rigid body + Euler-Maruyama update + stage cost; returns the undiscounted stage cost
template <int NU>
__device__ __forceinline__ float phys_step(const KParams& P, int t, const float (&x)[NX], const float (&u)[NU],
                                           const float (&up)[NU], const float (&r)[6], const float (&sig)[6],
                                           const float (&xi)[6], const float (&xr)[NX], float (&xn)[NX], float& rn) {
    const float* v = x + 3;
    const float* q = x + 6;
    const float* w = x + 10;
    float R[3][3];
    rotmat(q, R);
    float Tsum = 0.f, Mb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) Mb[k] = P.J[k] * r[3 + k];
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        const float T = P.kT * (u[i] * u[i]);
        Tsum = (i == 0) ? T : Tsum + T;
#pragma unroll
        for (int k = 0; k < 3; ++k) Mb[k] = fma_(P.mixer[k][i], T, Mb[k]);
    }
    const float fb0 = r[0], fb1 = r[1], fb2 = fma_(-Tsum, P.inv_m, r[2]);
    float acc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) acc[i] = fma_(R[i][2], fb2, fma_(R[i][1], fb1, R[i][0] * fb0));
    acc[2] = acc[2] + P.grav;
    const float qd0 = -0.5f * fma_(q[3], w[2], fma_(q[2], w[1], q[1] * w[0]));
    const float qd1 = 0.5f * fma_(-q[3], w[1], fma_(q[2], w[2], q[0] * w[0]));
    const float qd2 = 0.5f * fma_(q[3], w[0], fma_(-q[1], w[2], q[0] * w[1]));
    const float qd3 = 0.5f * fma_(-q[2], w[0], fma_(q[1], w[1], q[0] * w[2]));
    const float gy0 = (P.Jd[0] * w[1]) * w[2], gy1 = (P.Jd[1] * w[2]) * w[0], gy2 = (P.Jd[2] * w[0]) * w[1];
    const float wd0 = P.Jinv[0] * (Mb[0] - gy0), wd1 = P.Jinv[1] * (Mb[1] - gy1), wd2 = P.Jinv[2] * (Mb[2] - gy2);
    float sig2 = sig[0] * sig[0];
#pragma unroll
    for (int i = 1; i < 6; ++i) sig2 = fma_(sig[i], sig[i], sig2);
    const float dt = P.dt[t], sdt = P.sdt[t];
#pragma unroll
    for (int i = 0; i < 3; ++i) xn[i] = fma_(v[i], dt, x[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) xn[3 + i] = fma_(sig[i] * xi[i], sdt, fma_(acc[i], dt, v[i]));
    const float qt0 = fma_(qd0, dt, q[0]), qt1 = fma_(qd1, dt, q[1]), qt2 = fma_(qd2, dt, q[2]), qt3 = fma_(qd3, dt, q[3]);
    const float n2 = fma_(qt3, qt3, fma_(qt2, qt2, fma_(qt1, qt1, qt0 * qt0)));
    rn = det_rsqrt_near1(n2);
    xn[6] = qt0 * rn; xn[7] = qt1 * rn; xn[8] = qt2 * rn; xn[9] = qt3 * rn;
    xn[10] = fma_(sig[3] * xi[3], sdt, fma_(wd0, dt, w[0]));
    xn[11] = fma_(sig[4] * xi[4], sdt, fma_(wd1, dt, w[1]));
    xn[12] = fma_(sig[5] * xi[5], sdt, fma_(wd2, dt, w[2]));
    // ---- stage cost on (x_{t+1}, u_t) ----
    float l = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) { const float e = xn[i] - xr[i]; l = fma_(P.perr[i] * e, e, l); }
#pragma unroll
    for (int i = 0; i < 3; ++i) { const float e = xn[3 + i] - xr[3 + i]; l = fma_(P.verr[i] * e, e, l); }
    float eq[3];
    quat_err(xr + 6, xn + 6, eq);
#pragma unroll
    for (int i = 0; i < 3; ++i) l = fma_(P.qerr[i] * eq[i], eq[i], l);
#pragma unroll
    for (int i = 0; i < 3; ++i) { const float e = xn[10 + i] - xr[10 + i]; l = fma_(P.werr[i] * e, e, l); }
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        const float du = u[i] - P.uref[i], ds = u[i] - up[i];
        l = fma_(P.uerr * du, du, l);
        l = fma_(P.slew * ds, ds, l);
    }
    l = fma_(P.res_mult, sig2, l);
    return l;
}

// ---------------------------------------------------------------------------------
// Soft input-rate constraint (include/sdempc.h; iris_sitl_posctrl_mpc.yaml:40-41).  It depends on the control
// sequence only, so it is evaluated once per rollout as a warp-shaped vector operation instead of inside every
// step: J += sum_i w_t e_i^2 (32 strided partials + butterfly), g_i += 2 w_t e_i - 2 w_{t+1} e_{i+NU}, with
// e_i the violation of [lo, hi] by u_t[k] - u_{t-1}[k] and w_t = rate_w[t].  All lanes of the warp take part.
// ---------------------------------------------------------------------------------
template <int NU>
__device__ __forceinline__ float rate_violation(const KParams& P, const float* useq, const float* uprev, int t, int k) {
    const float ds = useq[t * NU + k] - (t == 0 ? uprev[k] : useq[(t - 1) * NU + k]);
    const float hi = P.slew_hi[k], lo = P.slew_lo[k];
    return ds > hi ? ds - hi : (ds < lo ? ds - lo : 0.f);
}

template <int NU>
__device__ __forceinline__ float rate_cost(const KParams& P, int lane, const float* useq, const float* uprev) {
    const int n = P.H * NU;
    float part = 0.f;
    for (int i = lane; i < n; i += 32) {
        const int t = i / NU, k = i % NU;
        const float e = rate_violation<NU>(P, useq, uprev, t, k);
        part = fma_(P.rate_w[t] * e, e, part);
    }
    return warp_butterfly(part);
}

template <int NU>
__device__ __forceinline__ void rate_grad_add(const KParams& P, int lane, const float* useq, const float* uprev, float* g) {
    const int n = P.H * NU;
    for (int i = lane; i < n; i += 32) {
        const int t = i / NU, k = i % NU;
        float a = (2.f * P.rate_w[t]) * rate_violation<NU>(P, useq, uprev, t, k);
        if (t + 1 < P.H) a = a - (2.f * P.rate_w[t + 1]) * rate_violation<NU>(P, useq, uprev, t + 1, k);
        g[i] = g[i] + a;
    }
}

// Both networks for NP problems at once (lane j = hidden unit j; NP independent dependency chains are
// interleaved for instruction-level parallelism): z (replicated registers) -> ob[0..5] = residual outputs,
// ob[6..11] = sigma, ob[12..17] = d sigma / d s (MODE 1).  mt: activation tape of each (problem, step)
// (MODE 1).  Problem p uses the exchange buffers at c.bufA/c.act3 + p * xstride.  Ends with a __syncwarp
// after which every ob is readable by every lane.
// MODE < 0: whether to record the tape is the run-time flag `rec` (one copy of the code for all rollouts).
template <int NU, int W, int MODE, int NP>
__device__ __forceinline__ void mlp_forward_n(const KParams& P, Warp<NU, W>& c, const float (&z)[NP][6 + NU], float2* const (&mt)[NP],
                                              float* const (&ob)[NP], int xstride, bool rec = false) {
    using L = Layout<NU, W>;
    constexpr int NIN = L::NIN, UPL = L::UPL;
    const int lane = c.lane;
    const bool tape = (MODE == 1) || (MODE < 0 && rec);
    // ---- layer 1: lane j, both nets ----
#pragma unroll
    for (int uu = 0; uu < UPL; ++uu) {
        float2 a[NP][4];
#pragma unroll
        for (int p = 0; p < NP; ++p) { a[p][0] = c.B1(uu); a[p][1] = a[p][2] = a[p][3] = make_float2(0.f, 0.f); }
#pragma unroll
        for (int k = 0; k < NIN; ++k) {
            const float2 wv = c.W1P(uu, k);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if constexpr (NP == 1) {   // latency path: two scalar chains are a little shorter than one packed chain
                    a[p][k & 3].x = fma_(wv.x, z[p][k], a[p][k & 3].x);
                    a[p][k & 3].y = fma_(wv.y, z[p][k], a[p][k & 3].y);
                } else {
                    a[p][k & 3] = fma2_(wv, splat(z[p][k]), a[p][k & 3]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float2 h = det_tanh2(add2_(add2_(a[p][0], a[p][1]), add2_(a[p][2], a[p][3])));
            reinterpret_cast<float2*>(c.bufA + p * xstride)[lane + 32 * uu] = h;
            if (tape) mt[p][lane + 32 * uu] = h;
        }
    }
    __syncwarp();
    // ---- layer 2 ----
    {
        float2 a[NP][UPL][4];
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
            for (int uu = 0; uu < UPL; ++uu) { a[p][uu][0] = c.B2(uu); a[p][uu][1] = a[p][uu][2] = a[p][uu][3] = make_float2(0.f, 0.f); }
#pragma unroll
        for (int k = 0; k < W; k += 2) {
            float4 hv[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) hv[p] = lds4(c.bufA + p * xstride + 2 * k);
#pragma unroll
            for (int uu = 0; uu < UPL; ++uu) {
                float2 w0, w1;
                if constexpr (L::WREG) { w0 = c.w2[uu][k]; w1 = c.w2[uu][k + 1]; }
                else { const float4 wv = lds4(c.ws + L::W2P + (lane + 32 * uu) * L::PAIR_STRIDE + 2 * k); w0 = xy(wv); w1 = zw(wv); }
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    a[p][uu][k & 3] = fma2_(w0, xy(hv[p]), a[p][uu][k & 3]);
                    a[p][uu][(k + 1) & 3] = fma2_(w1, zw(hv[p]), a[p][uu][(k + 1) & 3]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
            for (int uu = 0; uu < UPL; ++uu) {
                const float2 h = det_tanh2(add2_(add2_(a[p][uu][0], a[p][uu][1]), add2_(a[p][uu][2], a[p][uu][3])));
                c.act3[p * xstride + lane + 32 * uu] = h.x;
                c.act3[p * xstride + W + 4 + lane + 32 * uu] = h.y;
                if (tape) mt[p][W + lane + 32 * uu] = h;
            }
    }
    __syncwarp();
    // ---- output layer: lanes 0..5 drift rows, 6..11 diffusion rows ----
    if constexpr (NP == 2) {
        // lanes 0..11 serve problem 0, lanes 16..27 problem 1 (the others shadow row 11): one pass of weight-row
        // loads and half the activation loads of the sequential form; each quarter warp reads rows whose
        // 16-byte chunks fall in distinct banks
        const int p = lane >= L3_OFF ? 1 : 0;
        const int ll = lane - L3_OFF * p;
        const int o = ll < 12 ? ll : 11;
        const float* row = c.ws + L::W3R + o * L::W3R_STRIDE;
        const float* act = c.act3 + p * xstride + (o >= 6 ? W + 4 : 0);
        float2 aA = make_float2(c.ws[L::B3 + o], 0.f), aB = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < W; k += 4) {
            const float4 wv = lds4(row + k);
            const float4 hv = lds4(act + k);
            aA = fma2_(xy(wv), xy(hv), aA);
            aB = fma2_(zw(wv), zw(hv), aB);
        }
        const float out = (aA.x + aA.y) + (aB.x + aB.y);
        const float s0 = P.sig0[o >= 6 ? o - 6 : 0];
        float sp, sg;
        det_softplus_sigmoid_opt(out, sp, sg, tape);
        float* obp = p ? ob[1] : ob[0];
        if (ll < 12) {
            if (o < 6) obp[o] = out;
            else {
                obp[o] = s0 * sp;
                if (tape) obp[o + 6] = s0 * sg;
            }
        }
    } else {
        const int o = lane < 12 ? lane : 11;
        const float* row = c.ws + L::W3R + o * L::W3R_STRIDE;
        // the four partial sums of SPEC-ARITH packed two per FFMA2: (a0, a1) and (a2, a3)
        float2 aA[NP], aB[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) { aA[p] = make_float2(c.ws[L::B3 + o], 0.f); aB[p] = make_float2(0.f, 0.f); }
#pragma unroll
        for (int k = 0; k < W; k += 4) {
            const float4 wv = lds4(row + k);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const float4 hv = lds4(c.act3 + p * xstride + (o >= 6 ? W + 4 : 0) + k);
                aA[p] = fma2_(xy(wv), xy(hv), aA[p]);
                aB[p] = fma2_(zw(wv), zw(hv), aB[p]);
            }
        }
        const float s0 = P.sig0[o >= 6 ? o - 6 : 0];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float out = (aA[p].x + aA[p].y) + (aB[p].x + aB[p].y);
            float sp, sg;
            det_softplus_sigmoid(out, sp, sg);
            if (lane < 6) ob[p][lane] = out;
            else if (lane < 12) {
                ob[p][lane] = s0 * sp;
                if (tape) ob[p][lane + 6] = s0 * sg;
            }
        }
    }
    __syncwarp();
}

template <int NU, int W, int MODE>
__device__ __forceinline__ void mlp_forward(const KParams& P, Warp<NU, W>& c, const float (&z)[6 + NU], float2* mt, float* ob) {
    float z1[1][6 + NU];
#pragma unroll
    for (int i = 0; i < 6 + NU; ++i) z1[0][i] = z[i];
    float2* const mt1[1] = {mt};
    float* const ob1[1] = {ob};
    mlp_forward_n<NU, W, MODE, 1>(P, c, z1, mt1, ob1, 0);
}

__device__ __forceinline__ void load_r_sig(const float* ob, float (&r)[6], float (&sig)[6]) {
    const float4 a = lds4(ob), b = lds4(ob + 4), d = lds4(ob + 8);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y;
    sig[0] = b.z; sig[1] = b.w; sig[2] = d.x; sig[3] = d.y; sig[4] = d.z; sig[5] = d.w;
}

__device__ __forceinline__ void load6(const float* p, float (&o)[6]) {
    const float4 a = lds4(p);
    const float2 b = lds2(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y;
}

// ---------------------------------------------------------------------------------
// One Euler-Maruyama step + stage cost of the per-warp kernel.  MODE 0: cost only;
// 1: record the full tape for the adjoint; 2: record the state tape only (final
// x_evol pass).  x is advanced in place; returns the undiscounted stage cost.
// ---------------------------------------------------------------------------------
template <int NU, int W, int MODE>
__device__ __forceinline__ float fwd_step(const KParams& P, Warp<NU, W>& c, int t, float disc, float (&x)[NX],
                                          const float (&u)[NU], const float (&up)[NU]) {
    const int lane = c.lane;
    float z[6 + NU];
    phys_features<NU>(x, u, z);
    float* ob = (MODE == 1) ? (c.stape + t * 20) : c.lz;   // MODE != 1: 20-float scratch (lz has 24)
    mlp_forward<NU, W, MODE>(P, c, z, c.mtape + (size_t)t * 2 * W, ob);
    float r[6], sig[6], xi[6], xr[NX], xn[NX], rn;
    load_r_sig(ob, r, sig);
    load6(c.xi + t * 8, xi);
    load13(c.xref + (t + 1) * 16, xr);
    const float l = phys_step<NU>(P, t, x, u, up, r, sig, xi, xr, xn, rn);
    if constexpr (MODE == 1) { if (lane == 0) { ob[18] = rn; ob[19] = disc; } }
    if constexpr (MODE != 0) store13_lane0(c.xtape + (t + 1) * 16, xn, lane);
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = xn[i];
    return l;
}

// Forward rollout of this warp's particle at the control sequence `useq` (shared
// memory, [H][NU]).  Returns the particle cost J_p (identical in all lanes).
template <int NU, int W, int MODE>
__device__ __forceinline__ float rollout_fwd(const KParams& P, Warp<NU, W>& c, const float* useq, const float (&x0)[NX]) {
    float x[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = x0[i];
    if constexpr (MODE != 0) store13_lane0(c.xtape, x, c.lane);
    float up[NU], u[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) up[i] = c.uprev[i];
    float Jp = 0.f, disc = 1.f;
    for (int t = 0; t < P.H; ++t) {
        load_u<NU>(useq, t, u);
        const float l = fwd_step<NU, W, MODE>(P, c, t, disc, x, u, up);
        Jp = fma_(disc, l, Jp);
        disc = disc * P.discount;
#pragma unroll
        for (int i = 0; i < NU; ++i) up[i] = u[i];
    }
    if constexpr (MODE != 0) __syncwarp();   // lane 0's last tape writes are visible to every lane that reads the tape next
    if (SDEMPC_RATE_ON(P)) Jp = Jp + rate_cost<NU>(P, c.lane, useq, c.uprev);
    return Jp;
}

// ---------------------------------------------------------------------------------
// Adjoint of one step, in three pieces: bwd_pre (cost, normalisation, EM and rigid-body
// adjoints down to the network outputs), mlp_backward (both networks, transposed),
// bwd_post (feature / rotation adjoints, direct input terms).
// ---------------------------------------------------------------------------------
struct BwdMid {
    float lp[3], lv[3], lw[3], lq[4], la[3], fb[3], g2;
};

// lam: dJ/dx_{t+1} WITHOUT this stage's direct cost.  Outputs: mid, lo[6] = (drift, diffusion) output adjoints,
// gu[NU] = thrust part of dJ/du_t.
template <int NU>
__device__ __forceinline__ void bwd_pre(const KParams& P, int t, const float (&x)[NX], const float (&xn)[NX], const float (&xr)[NX],
                                        const float (&u)[NU], const float (&r012)[3], const float (&sig)[6], const float (&dsg)[6],
                                        float rn, float disc, const float (&xi)[6], float (&lam)[NX], BwdMid& m, float2 (&lo)[6],
                                        float (&gu)[NU]) {
    const float dt = P.dt[t], sdt = P.sdt[t];
    const float* q = x + 6;
    const float* w = x + 10;
    const float g2 = 2.f * disc;
    m.g2 = g2;
    // direct cost on x_{t+1}
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        lam[i] = fma_(g2 * P.perr[i], xn[i] - xr[i], lam[i]);
        lam[3 + i] = fma_(g2 * P.verr[i], xn[3 + i] - xr[3 + i], lam[3 + i]);
        lam[10 + i] = fma_(g2 * P.werr[i], xn[10 + i] - xr[10 + i], lam[10 + i]);
    }
    {
        const float* rq = xr + 6;
        float e[3];
        quat_err(rq, xn + 6, e);
        const float k0 = (g2 * P.qerr[0]) * e[0], k1 = (g2 * P.qerr[1]) * e[1], k2 = (g2 * P.qerr[2]) * e[2];
        lam[6] = lam[6] - fma_(rq[3], k2, fma_(rq[2], k1, rq[1] * k0));
        lam[7] = lam[7] + fma_(rq[2], k2, fma_(-rq[3], k1, rq[0] * k0));
        lam[8] = lam[8] + fma_(-rq[1], k2, fma_(rq[0], k1, rq[3] * k0));
        lam[9] = lam[9] + fma_(rq[0], k2, fma_(rq[1], k1, (-rq[2]) * k0));
    }
    float lqt[4];
    {
        const float dot = fma_(xn[9], lam[9], fma_(xn[8], lam[8], fma_(xn[7], lam[7], xn[6] * lam[6])));
#pragma unroll
        for (int i = 0; i < 4; ++i) lqt[i] = fma_(-xn[6 + i], dot, lam[6 + i]) * rn;
    }
    float lwd[3], lsig[6];
    const float rs2 = g2 * P.res_mult;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        m.lp[i] = lam[i];
        m.lv[i] = fma_(dt, lam[i], lam[3 + i]);
        m.la[i] = dt * lam[3 + i];
        m.lw[i] = lam[10 + i];
        lwd[i] = dt * lam[10 + i];
        lsig[i] = fma_(rs2, sig[i], (lam[3 + i] * xi[i]) * sdt);
        lsig[3 + i] = fma_(rs2, sig[3 + i], (lam[10 + i] * xi[3 + i]) * sdt);
    }
    {
        const float l0 = dt * lqt[0], l1 = dt * lqt[1], l2 = dt * lqt[2], l3 = dt * lqt[3];
        m.lq[0] = fma_(0.5f, fma_(w[2], l3, fma_(w[1], l2, w[0] * l1)), lqt[0]);
        m.lq[1] = fma_(0.5f, fma_(w[1], l3, fma_(-w[2], l2, (-w[0]) * l0)), lqt[1]);
        m.lq[2] = fma_(0.5f, fma_(-w[0], l3, fma_(w[2], l1, (-w[1]) * l0)), lqt[2]);
        m.lq[3] = fma_(0.5f, fma_(w[0], l2, fma_(-w[1], l1, (-w[2]) * l0)), lqt[3]);
        m.lw[0] = fma_(0.5f, fma_(-q[2], l3, fma_(q[3], l2, fma_(q[0], l1, (-q[1]) * l0))), m.lw[0]);
        m.lw[1] = fma_(0.5f, fma_(q[1], l3, fma_(q[0], l2, fma_(-q[3], l1, (-q[2]) * l0))), m.lw[1]);
        m.lw[2] = fma_(0.5f, fma_(q[0], l3, fma_(-q[1], l2, fma_(q[2], l1, (-q[3]) * l0))), m.lw[2]);
    }
    const float lMb[3] = {P.Jinv[0] * lwd[0], P.Jinv[1] * lwd[1], P.Jinv[2] * lwd[2]};
    m.lw[0] = m.lw[0] - fma_(P.Jd[2] * w[1], lMb[2], (P.Jd[1] * w[2]) * lMb[1]);
    m.lw[1] = m.lw[1] - fma_(P.Jd[2] * w[0], lMb[2], (P.Jd[0] * w[2]) * lMb[0]);
    m.lw[2] = m.lw[2] - fma_(P.Jd[1] * w[0], lMb[1], (P.Jd[0] * w[1]) * lMb[0]);
    float R[3][3];
    rotmat(q, R);
    float Tsum = 0.f;
#pragma unroll
    for (int i = 0; i < NU; ++i) { const float T = P.kT * (u[i] * u[i]); Tsum = (i == 0) ? T : Tsum + T; }
    m.fb[0] = r012[0]; m.fb[1] = r012[1]; m.fb[2] = fma_(-Tsum, P.inv_m, r012[2]);
    float lfb[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) lfb[j] = fma_(R[2][j], m.la[2], fma_(R[1][j], m.la[1], R[0][j] * m.la[0]));
#pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k].x = lfb[k]; lo[3 + k].x = P.J[k] * lMb[k]; }
#pragma unroll
    for (int i = 0; i < 6; ++i) lo[i].y = lsig[i] * dsg[i];
    const float lTsum = -(lfb[2] * P.inv_m);
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        const float lT = fma_(P.mixer[2][i], lMb[2], fma_(P.mixer[1][i], lMb[1], fma_(P.mixer[0][i], lMb[0], lTsum)));
        gu[i] = (P.kT2 * u[i]) * lT;
    }
}

// lz[NIN] (feature adjoints) -> lzbuf (shared) for NP problems at once; lo replicated in registers; mt =
// activation tape of each (problem, step); problem p uses c.bufA/c.bufB + p * xstride.  Ends with a
// __syncwarp after which every lzbuf is readable by every lane.
template <int NU, int W, int NP>
__device__ __forceinline__ void mlp_backward_n(Warp<NU, W>& c, const float2 (&lo)[NP][6], const float2* const (&mt)[NP],
                                               float* const (&lzbuf)[NP], int xstride) {
    using L = Layout<NU, W>;
    constexpr int NIN = L::NIN, UPL = L::UPL;
    const int lane = c.lane;
    const float2 one = make_float2(1.f, 1.f);
#pragma unroll
    for (int uu = 0; uu < UPL; ++uu) {
        float2 w3[6];
#pragma unroll
        for (int o = 0; o < 6; ++o) w3[o] = c.W3C(uu, o);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            float2 a0 = fma2_(w3[0], lo[p][0], make_float2(0.f, 0.f));
            float2 a1 = fma2_(w3[1], lo[p][1], make_float2(0.f, 0.f));
            const float2 a2 = fma2_(w3[2], lo[p][2], make_float2(0.f, 0.f));
            const float2 a3 = fma2_(w3[3], lo[p][3], make_float2(0.f, 0.f));
            a0 = fma2_(w3[4], lo[p][4], a0);
            a1 = fma2_(w3[5], lo[p][5], a1);
            const float2 h = mt[p][W + lane + 32 * uu];
            const float2 d = mul2_(add2_(add2_(a0, a1), add2_(a2, a3)), fma2_(make_float2(-h.x, -h.y), h, one));
            reinterpret_cast<float2*>(c.bufA + p * xstride)[lane + 32 * uu] = d;
        }
    }
    __syncwarp();
#pragma unroll
    for (int uu = 0; uu < UPL; ++uu) {
        const float* row = c.ws + L::W2T + (lane + 32 * uu) * L::PAIR_STRIDE;
        float2 a[NP][4];
#pragma unroll
        for (int p = 0; p < NP; ++p) a[p][0] = a[p][1] = a[p][2] = a[p][3] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < W; j += 2) {
            const float4 wv = lds4(row + 2 * j);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const float4 dv = lds4(c.bufA + p * xstride + 2 * j);
                a[p][j & 3] = fma2_(xy(wv), xy(dv), a[p][j & 3]);
                a[p][(j + 1) & 3] = fma2_(zw(wv), zw(dv), a[p][(j + 1) & 3]);
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float2 h = mt[p][lane + 32 * uu];
            const float2 d = mul2_(add2_(add2_(a[p][0], a[p][1]), add2_(a[p][2], a[p][3])), fma2_(make_float2(-h.x, -h.y), h, one));
            reinterpret_cast<float2*>(c.bufB + p * xstride)[lane + 32 * uu] = d;
        }
    }
    __syncwarp();
    if constexpr (NP == 2) {
        // lanes 0..NIN-1 serve problem 0, lanes 16..16+NIN-1 problem 1 (the others shadow the last row)
        const int p = lane >= LZ_OFF ? 1 : 0;
        const int ll = lane - LZ_OFF * p;
        const int i = ll < NIN ? ll : NIN - 1;
        const float* row = c.ws + L::W1T + i * L::PAIR_STRIDE;
        const float* db = c.bufB + p * xstride;
        float2 a[4];
        a[0] = a[1] = a[2] = a[3] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < W; j += 2) {
            const float4 wv = lds4(row + 2 * j);
            const float4 dv = lds4(db + 2 * j);
            a[j & 3] = fma2_(xy(wv), xy(dv), a[j & 3]);
            a[(j + 1) & 3] = fma2_(zw(wv), zw(dv), a[(j + 1) & 3]);
        }
        const float2 s = add2_(add2_(a[0], a[1]), add2_(a[2], a[3]));
        float* lzp = p ? lzbuf[1] : lzbuf[0];
        if (ll < NIN) lzp[i] = s.x + s.y;
    } else {
        const int i = lane < NIN ? lane : NIN - 1;
        const float* row = c.ws + L::W1T + i * L::PAIR_STRIDE;
        float2 a[NP][4];
#pragma unroll
        for (int p = 0; p < NP; ++p) a[p][0] = a[p][1] = a[p][2] = a[p][3] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < W; j += 2) {
            const float4 wv = lds4(row + 2 * j);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const float4 dv = lds4(c.bufB + p * xstride + 2 * j);
                a[p][j & 3] = fma2_(xy(wv), xy(dv), a[p][j & 3]);
                a[p][(j + 1) & 3] = fma2_(zw(wv), zw(dv), a[p][(j + 1) & 3]);
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float2 s = add2_(add2_(a[p][0], a[p][1]), add2_(a[p][2], a[p][3]));
            if (lane < NIN) lzbuf[p][lane] = s.x + s.y;
        }
    }
    __syncwarp();
}

template <int NU, int W>
__device__ __forceinline__ void mlp_backward(Warp<NU, W>& c, const float2 (&lo)[6], const float2* mt, float* lzbuf) {
    float2 lo1[1][6];
#pragma unroll
    for (int i = 0; i < 6; ++i) lo1[0][i] = lo[i];
    const float2* const mt1[1] = {mt};
    float* const lz1[1] = {lzbuf};
    mlp_backward_n<NU, W, 1>(c, lo1, mt1, lz1, 0);
}

// gp: slew term of step t+1 w.r.t. u_t on entry, of step t w.r.t. u_{t-1} on exit; gu: dJ/du_t on exit.
template <int NU>
__device__ __forceinline__ void bwd_post(const KParams& P, const float (&x)[NX], const float (&u)[NU], const float (&up)[NU],
                                         BwdMid& m, const float (&lz)[6 + NU], float (&gu)[NU], float (&gp)[NU], float (&lam)[NX]) {
    const float* v = x + 3;
    const float* q = x + 6;
    float R[3][3];
    rotmat(q, R);
#pragma unroll
    for (int i = 0; i < 3; ++i) m.lw[i] = m.lw[i] + lz[3 + i];
#pragma unroll
    for (int i = 0; i < NU; ++i) gu[i] = gu[i] + lz[6 + i];
    {
        float M[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) M[i][j] = fma_(v[i], lz[j], m.la[i] * m.fb[j]);
        const float qw = q[0], qx = q[1], qy = q[2], qz = q[3];
        const float m2x = -2.f * qx, m2y = -2.f * qy, m2z = -2.f * qz;
        const float d0 = fma_(qx, M[2][1], fma_(-qy, M[2][0], fma_(-qx, M[1][2], fma_(qz, M[1][0], fma_(qy, M[0][2], (-qz) * M[0][1])))));
        const float d1 = fma_(m2x, M[2][2], fma_(qw, M[2][1], fma_(qz, M[2][0], fma_(-qw, M[1][2], fma_(m2x, M[1][1], fma_(qy, M[1][0], fma_(qz, M[0][2], qy * M[0][1])))))));
        const float d2 = fma_(m2y, M[2][2], fma_(qz, M[2][1], fma_(-qw, M[2][0], fma_(qz, M[1][2], fma_(qx, M[1][0], fma_(qw, M[0][2], fma_(qx, M[0][1], m2y * M[0][0])))))));
        const float d3 = fma_(qy, M[2][1], fma_(qx, M[2][0], fma_(qy, M[1][2], fma_(m2z, M[1][1], fma_(qw, M[1][0], fma_(qx, M[0][2], fma_(-qw, M[0][1], m2z * M[0][0])))))));
        m.lq[0] = fma_(2.f, d0, m.lq[0]); m.lq[1] = fma_(2.f, d1, m.lq[1]);
        m.lq[2] = fma_(2.f, d2, m.lq[2]); m.lq[3] = fma_(2.f, d3, m.lq[3]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) m.lv[i] = fma_(R[i][2], lz[2], fma_(R[i][1], lz[1], fma_(R[i][0], lz[0], m.lv[i])));
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        const float ds = (m.g2 * P.slew) * (u[i] - up[i]);
        gu[i] = fma_(m.g2 * P.uerr, u[i] - P.uref[i], gu[i]) + ds;
        const float gt = gu[i] + gp[i];
        gp[i] = -ds;
        gu[i] = gt;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { lam[i] = m.lp[i]; lam[3 + i] = m.lv[i]; lam[10 + i] = m.lw[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) lam[6 + i] = m.lq[i];
}

// per-step tape fields written by the forward pass (MODE 1)
__device__ __forceinline__ void load_step_tape(const float* ob, float (&r012)[3], float (&sig)[6], float (&dsg)[6], float& rn, float& disc) {
    const float4 a = lds4(ob), b = lds4(ob + 4), d = lds4(ob + 8), e = lds4(ob + 12), f = lds4(ob + 16);
    r012[0] = a.x; r012[1] = a.y; r012[2] = a.z;
    sig[0] = b.z; sig[1] = b.w; sig[2] = d.x; sig[3] = d.y; sig[4] = d.z; sig[5] = d.w;
    dsg[0] = e.x; dsg[1] = e.y; dsg[2] = e.z; dsg[3] = e.w; dsg[4] = f.x; dsg[5] = f.y;
    rn = f.z; disc = f.w;
}

// ---------------------------------------------------------------------------------
// Adjoint sweep of the per-warp kernel (after rollout_fwd<MODE=1> at the same useq):
// writes this particle's gradient to c.g[H][NU].
// ---------------------------------------------------------------------------------
template <int NU, int W>
__device__ __forceinline__ void rollout_bwd(const KParams& P, Warp<NU, W>& c, const float* useq) {
    constexpr int NIN = 6 + NU;
    const int lane = c.lane;
    float lam[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) lam[i] = 0.f;
    float gp[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) gp[i] = 0.f;
    for (int t = P.H - 1; t >= 0; --t) {
        float x[NX], xn[NX], xr[NX], u[NU], up[NU];
        load13(c.xtape + t * 16, x);
        load13(c.xtape + (t + 1) * 16, xn);
        load13(c.xref + (t + 1) * 16, xr);
        load_u<NU>(useq, t, u);
        if (t == 0) {
#pragma unroll
            for (int i = 0; i < NU; ++i) up[i] = c.uprev[i];
        } else load_u<NU>(useq, t - 1, up);
        float r012[3], sig[6], dsg[6], rn, disc, xi[6];
        load_step_tape(c.stape + t * 20, r012, sig, dsg, rn, disc);
        load6(c.xi + t * 8, xi);
        BwdMid mid;
        float2 lo[6];
        float gu[NU];
        bwd_pre<NU>(P, t, x, xn, xr, u, r012, sig, dsg, rn, disc, xi, lam, mid, lo, gu);
        mlp_backward<NU, W>(c, lo, c.mtape + (size_t)t * 2 * W, c.lz);
        float lz[NIN];
#pragma unroll
        for (int i = 0; i < NIN; i += 2) { const float2 a = lds2(c.lz + i); lz[i] = a.x; lz[i + 1] = a.y; }
        bwd_post<NU>(P, x, u, up, mid, lz, gu, gp, lam);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < NU; ++i) c.g[t * NU + i] = gu[i];
        }
    }
    __syncwarp();
    if (SDEMPC_RATE_ON(P)) {
        rate_grad_add<NU>(P, lane, useq, c.uprev, c.g);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Team (= P warps, one problem) helpers.  P == 1: everything is warp local.
// ---------------------------------------------------------------------------------
template <int PP>
struct Team {
    int warp_in_team;   // particle index
    int bar_id;         // named barrier (P > 1)
    float* scratch;     // team scratch in shared memory: [PP] costs
    float* warp0_base;  // per-warp region of the team's first warp (of this replica when LSW > 1)
    float* team0_base;  // per-warp region of the first warp of the whole team (all replicas)
    float* spec0_base;  // per-warp region of the team's first speculation warp (possibly in the sibling CTA)
    int ws_stride;
    int ls_index;       // speculative line search: this warp evaluates trial `ls_index` (0 when LSW == 1)
    int ls_bar_id;      // named barrier of the LSW sibling warps
    __device__ __forceinline__ void sync() const {
        if constexpr (PP == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(PP * 32) : "memory");
    }
};

// mean over particles of the cost (sequential sum in particle order, times 1/P)
template <int PP>
__device__ __forceinline__ float team_mean_cost(const Team<PP>& tm, float Jp, int lane, float invP) {
    if constexpr (PP == 1) return Jp * invP;
    else {
        tm.sync();   // previous readers done
        if (lane == 0) tm.scratch[tm.warp_in_team] = Jp;
        tm.sync();
        float J = tm.scratch[0];
#pragma unroll
        for (int p = 1; p < PP; ++p) J = J + tm.scratch[p];
        return J * invP;
    }
}

// mean over particles of the per-warp gradient buffers -> every warp's c.g
template <int NU, int W, int PP>
__device__ __forceinline__ void team_mean_grad(const KParams& P, const Team<PP>& tm, Warp<NU, W>& c, int n, float invP) {
    if constexpr (PP == 1) {
        for (int i = c.lane; i < n; i += 32) c.g[i] = c.g[i] * invP;
        __syncwarp();
    } else {
        tm.sync();
        float acc[(SDEMPC_MAX_H * SDEMPC_MAX_NU + 31) / 32];
        int m = 0;
        for (int i = c.lane; i < n; i += 32, ++m) {
            float s = tm.warp0_base[P.o_g + i];
#pragma unroll
            for (int p = 1; p < PP; ++p) s = s + tm.warp0_base[p * tm.ws_stride + P.o_g + i];
            acc[m] = s * invP;
        }
        tm.sync();
        m = 0;
        for (int i = c.lane; i < n; i += 32, ++m) c.g[i] = acc[m];
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// The APG solve of one problem (SPEC "APG").  On entry: c.xk holds the (shifted,
// clipped) plan, c.uprev the slew reference, c.xref / c.xi the window and noise,
// x0 the internal-frame state; s the carried step size.  On exit c.xk = u*, the
// state tape holds this warp's particle trajectory at u*, and `inf` the telemetry.
// ---------------------------------------------------------------------------------
// (The latency variant with concurrent line-search trials and speculative gradients is apg_solve_latency.)
template <int NU, int W, int PP, int LSW>
__device__ __forceinline__ void apg_solve(const KParams& P, Warp<NU, W>& c, const Team<PP>& tm, const float (&x0)[NX],
                                          float s, sdempc_info& inf, float* trace) {
    static_assert(LSW == 1, "the latency variant is apg_solve_latency");
    const int lane = c.lane;
    const int n = P.H * NU;
    const float invP = __fdiv_rn(1.0f, (float)PP);
    for (int i = lane; i < n; i += 32) c.yk[i] = c.xk[i];
    __syncwarp();
    float Jx = 0.f, Jp = 0.f, fy = 0.f, gsq = 0.f, sum_ls = 0.f, sum_s = 0.f, init_cost = 0.f;
    int k = 1, no_improve = 0, it = 0;
    for (;;) {
        ++it;
        {
            const float Jw = rollout_fwd<NU, W, 1>(P, c, c.yk, x0);
            rollout_bwd<NU, W>(P, c, c.yk);
            fy = team_mean_cost<PP>(tm, Jw, lane, invP);
            team_mean_grad<NU, W, PP>(P, tm, c, n, invP);
        }
        if (it == 1) { Jx = fy; init_cost = fy; }
        {
            float part = 0.f;
            for (int i = lane; i < n; i += 32) { const float gi = c.g[i]; part = fma_(gi, gi, part); }
            gsq = warp_butterfly(part);
        }
        if (P.reset_option == 1) { s = s * P.inc_f; s = s > P.max_step ? P.max_step : s; }
        bool ok = false;
        int n_ls = 0;
        if constexpr (LSW == 1) {
            for (int j = 0; j <= P.maxls; ++j) {
                float part = 0.f;
                for (int i = lane; i < n; i += 32) {
                    const int ii = i % NU;
                    const float gi = c.g[i], yi = c.yk[i];
                    const float xv = clipf(fma_(-s, gi, yi), P.u_lo[ii], P.u_hi[ii]);
                    c.xp[i] = xv;
                    part = fma_(gi, xv - yi, part);
                }
                const float dec = warp_butterfly(part);
                __syncwarp();
                const float Jw = rollout_fwd<NU, W, 0>(P, c, c.xp, x0);
                Jp = team_mean_cost<PP>(tm, Jw, lane, invP);
                n_ls = j + 1;
                ok = (Jp <= fma_(P.coef, dec, fy));
                if (ok) break;
                if (j < P.maxls) s = s * P.dec_f;
            }
        }
        sum_ls = sum_ls + (float)n_ls;
        sum_s = sum_s + s;
        const bool accept = ok && (Jp <= Jx);
        bool converged = false;
        if (accept) {
            const float beta = apg_momentum(P, k);
            for (int i = lane; i < n; i += 32) {
                const int ii = i % NU;
                const float xv = c.xp[i];
                c.yk[i] = clipf(fma_(beta, xv - c.xk[i], xv), P.u_lo[ii], P.u_hi[ii]);
                c.xk[i] = xv;
            }
            const float Jprev = Jx;
            Jx = Jp; ++k; no_improve = 0;
            const float tol = P.atol + P.rtol * fabsf(Jprev);
            converged = (fabsf(Jprev - Jx) <= tol) || (Jx <= P.atol);
        } else {
            for (int i = lane; i < n; i += 32) c.yk[i] = c.xk[i];
            k = 1; ++no_improve;
        }
        __syncwarp();
        if (trace != nullptr && tm.warp_in_team == 0 && tm.ls_index == 0 && lane == 0) {
            float* tr = trace + (size_t)(it - 1) * SDEMPC_TRACE_W;
            tr[0] = fy; tr[1] = Jp; tr[2] = s; tr[3] = (float)n_ls; tr[4] = accept ? 1.f : 0.f; tr[5] = Jx; tr[6] = gsq; tr[7] = (float)k;
        }
        if (it >= P.max_iter || no_improve >= P.max_no_improve || converged || !(fy == fy)) break;
    }
    (void)rollout_fwd<NU, W, 2>(P, c, c.xk, x0);
    __syncwarp();
    inf.avg_linesearch = __fdiv_rn(sum_ls, (float)it);
    inf.stepsize = s;
    inf.num_steps = (float)it;
    inf.grad_sqr = gsq;
    inf.avg_stepsize = __fdiv_rn(sum_s, (float)it);
    inf.init_cost = init_cost;
    inf.opt_cost = (Jx == Jx) ? Jx : __int_as_float(0x7f800000);
    inf.solve_time_us = 0.f;
}

// ---------------------------------------------------------------------------------
// Latency-mode APG solve (P = 1): LSW line-search warps + SGW speculative-gradient warps run the same solve
// as replicas.  Per iteration, while LS warp l evaluates trial l (step s*dec^l), speculation warp c already
// computes value_and_grad at the NEXT extrapolation point for one possible outcome of the line search
// (c < SGW-1: "trial c accepted"; c = SGW-1: "step rejected", where the next point is x_k itself).  When the
// outcome is one of those, the next iteration starts with its gradient already known and the line search is
// off the critical path; otherwise every warp computes the gradient as usual.  Candidates are built with the
// same arithmetic as the real update, so the result is bit-identical to the sequential solve.
// ---------------------------------------------------------------------------------
// CL = 1: the two halves of the team are the two CTAs of a thread-block cluster (line-search warps on one SM,
// speculation warps on the neighbouring SM, so that neither slows the other down); slots and gradients are
// exchanged through distributed shared memory and the team barrier is the cluster barrier.
template <int NU, int W, int LSW, int SGW, int CL>
__device__ __forceinline__ void apg_solve_latency(const KParams& P, Warp<NU, W>& c, const Team<1>& tm, const float (&x0)[NX],
                                                  float s, sdempc_info& inf, float* trace) {
    constexpr int TW = LSW + SGW;
    const int lane = c.lane;
    const int n = P.H * NU;
    const int l = tm.ls_index;              // 0..LSW-1: line-search warp; LSW..TW-1: speculation warp
    const bool is_spec = (l >= LSW);
    const int cand = l - LSW;               // speculation candidate of this warp (if is_spec)
    float* const ls_slots = tm.scratch;            // [2][LSW][2]
    float* const sg_slots = tm.scratch + 4 * LSW;  // [SGW]
    auto team_bar = [&]() {
        if constexpr (CL) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        else asm volatile("bar.sync %0, %1;" ::"r"(tm.ls_bar_id), "r"(TW * 32) : "memory");
    };
    for (int i = lane; i < n; i += 32) c.yk[i] = c.xk[i];
    __syncwarp();
    float Jx = 0.f, Jp = 0.f, fy = 0.f, gsq = 0.f, sum_ls = 0.f, sum_s = 0.f, init_cost = 0.f;
    int k = 1, no_improve = 0, it = 0, ls_round = 0;
    bool boot = true;
    for (;;) {
        if (boot) {   // gradient at y_k computed by every replica (first iteration, or a speculation miss)
            ++it;
            fy = rollout_fwd<NU, W, 1>(P, c, c.yk, x0);
            rollout_bwd<NU, W>(P, c, c.yk);
            if (it == 1) { Jx = fy; init_cost = fy; }
        }
        {
            float part = 0.f;
            for (int i = lane; i < n; i += 32) { const float gi = c.g[i]; part = fma_(gi, gi, part); }
            gsq = warp_butterfly(part);
        }
        if (P.reset_option == 1) { s = s * P.inc_f; s = s > P.max_step ? P.max_step : s; }
        const float s0 = s;
        const float beta = apg_momentum_tab(P, k);
        // ---- speculation warps: value_and_grad at the candidate next point, into g2 ----
        bool spec_valid = false;
        if (SGW > 0 && is_spec) {
            const bool rej = (cand == SGW - 1);
            if (rej || cand <= P.maxls) {
                float s_c = s0;
                for (int q = 0; q < cand && !rej; ++q) s_c = s_c * P.dec_f;
                for (int i = lane; i < n; i += 32) {
                    const int ii = i % NU;
                    float v;
                    if (rej) v = c.xk[i];
                    else {
                        const float xv = clipf(fma_(-s_c, c.g[i], c.yk[i]), P.u_lo[ii], P.u_hi[ii]);
                        v = clipf(fma_(beta, xv - c.xk[i], xv), P.u_lo[ii], P.u_hi[ii]);
                    }
                    c.xp[i] = v;
                }
                __syncwarp();
                spec_valid = true;
            }
        }
        // ---- line search in rounds of LSW concurrent trials (see apg_solve) ----
        bool ok = false;
        int n_ls = 0, jsel = 0, base = 0;
        float s_base = s0;
        float* slot = nullptr;
        for (int round = 0;; ++round) {
            const int j = base + l;
            slot = ls_slots + 2 * LSW * (ls_round & 1);
            ++ls_round;
            float Jl = 0.f, dec = 0.f;
            bool do_trial = !is_spec && j <= P.maxls;
            if (do_trial) {
                float s_l = s_base;
                for (int q = 0; q < l; ++q) s_l = s_l * P.dec_f;
                float part = 0.f;
                for (int i = lane; i < n; i += 32) {
                    const int ii = i % NU;
                    const float gi = c.g[i], yi = c.yk[i];
                    const float xv = clipf(fma_(-s_l, gi, yi), P.u_lo[ii], P.u_hi[ii]);
                    c.xp[i] = xv;
                    part = fma_(gi, xv - yi, part);
                }
                dec = warp_butterfly(part);
                __syncwarp();
            }
            const bool do_grad = is_spec && spec_valid && round == 0;
            if (do_grad) {   // speculative value_and_grad at c.xp -> c.g2
                float* gsave = c.g;
                c.g = c.g2;
                const float fc = rollout_fwd<NU, W, 1>(P, c, c.xp, x0);
                rollout_bwd<NU, W>(P, c, c.xp);
                c.g = gsave;
                if (lane == 0) sg_slots[cand] = fc;
            }
            if (do_trial) {
                Jl = rollout_fwd<NU, W, 0>(P, c, c.xp, x0);
                if (lane == 0) { slot[2 * l] = Jl; slot[2 * l + 1] = (Jl <= fma_(P.coef, dec, fy)) ? 1.f : 0.f; }
            }
            team_bar();
            int q = 0;
            for (; q < LSW && base + q <= P.maxls; ++q)
                if (slot[2 * q + 1] != 0.f) { ok = true; break; }
            if (ok) { jsel = base + q; break; }
            if (base + LSW > P.maxls) { jsel = P.maxls; break; }   // every trial failed
            for (int r = 0; r < LSW; ++r) s_base = s_base * P.dec_f;
            base += LSW;
        }
        s = s0;
        for (int q = 0; q < jsel; ++q) s = s * P.dec_f;
        Jp = slot[2 * (jsel - base)];
        n_ls = jsel + 1;
        // adopt the selected trial point (same arithmetic as the warp that evaluated it)
        for (int i = lane; i < n; i += 32) {
            const int ii = i % NU;
            c.xp[i] = clipf(fma_(-s, c.g[i], c.yk[i]), P.u_lo[ii], P.u_hi[ii]);
        }
        __syncwarp();
        sum_ls = sum_ls + (float)n_ls;
        sum_s = sum_s + s;
        const bool accept = ok && (Jp <= Jx);
        bool converged = false;
        if (accept) {
            for (int i = lane; i < n; i += 32) {
                const int ii = i % NU;
                const float xv = c.xp[i];
                c.yk[i] = clipf(fma_(beta, xv - c.xk[i], xv), P.u_lo[ii], P.u_hi[ii]);
                c.xk[i] = xv;
            }
            const float Jprev = Jx;
            Jx = Jp; ++k; no_improve = 0;
            const float tol = P.atol + P.rtol * fabsf(Jprev);
            converged = (fabsf(Jprev - Jx) <= tol) || (Jx <= P.atol);
        } else {
            for (int i = lane; i < n; i += 32) c.yk[i] = c.xk[i];
            k = 1; ++no_improve;
        }
        __syncwarp();
        if (trace != nullptr && l == 0 && lane == 0) {
            float* tr = trace + (size_t)(it - 1) * SDEMPC_TRACE_W;
            tr[0] = fy; tr[1] = Jp; tr[2] = s; tr[3] = (float)n_ls; tr[4] = accept ? 1.f : 0.f; tr[5] = Jx; tr[6] = gsq; tr[7] = (float)k;
        }
        if (it >= P.max_iter || no_improve >= P.max_no_improve || converged || !(fy == fy)) break;
        // ---- did a speculation warp already compute the gradient at the new y_k? ----
        int hit = -1;
        if (SGW > 0) {
            if (!accept) hit = SGW - 1;
            else if (jsel < SGW - 1) hit = jsel;
        }
        if (hit >= 0) {
            ++it;
            fy = sg_slots[hit];
            const float* src = tm.spec0_base + (size_t)hit * tm.ws_stride + P.o_g2;
            for (int i = lane; i < n; i += 32) c.g[i] = src[i];
            __syncwarp();
            boot = false;
        } else {
            boot = true;
        }
        if (SGW > 0) team_bar();   // g2 / slots may be overwritten by the next pass
    }
    if (SGW > 0) team_bar();       // nobody still reads this solve's slots when the next one starts writing
    (void)rollout_fwd<NU, W, 2>(P, c, c.xk, x0);
    __syncwarp();
    inf.avg_linesearch = __fdiv_rn(sum_ls, (float)it);
    inf.stepsize = s;
    inf.num_steps = (float)it;
    inf.grad_sqr = gsq;
    inf.avg_stepsize = __fdiv_rn(sum_s, (float)it);
    inf.init_cost = init_cost;
    inf.opt_cost = (Jx == Jx) ? Jx : __int_as_float(0x7f800000);
    inf.solve_time_us = 0.f;
}

// Philox noise of this warp's particle for one solve: xi[t][0..5], t < H (lane t)
__device__ __forceinline__ void gen_noise(int lane, unsigned long long seed, unsigned long long tick,
                                          uint32_t particle, uint32_t sub0, int H, float* dst /*[H][8]*/) {
    const int t = lane;
    if (t < H) {
        const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        const uint32_t c2 = (uint32_t)tick, c3 = ((uint32_t)(tick >> 32)) & 0x3FFFFFFFu;
        uint32_t r[4];
        float n[6];
        philox4x32_10((uint32_t)t, particle, c2, c3 | (sub0 << 30), k0, k1, r);
        box_muller(r[0], r[1], n[0], n[1]);
        box_muller(r[2], r[3], n[2], n[3]);
        philox4x32_10((uint32_t)t, particle, c2, c3 | ((sub0 + 1) << 30), k0, k1, r);
        box_muller(r[0], r[1], n[4], n[5]);
#pragma unroll
        for (int i = 0; i < 6; ++i) dst[t * 8 + i] = n[i];
    }
}

// Reference window into c.xref[(H+1)][16] (internal frame); lane t builds row t.
__device__ __forceinline__ void build_window(const KParams& P, int lane, float* xref_dst, const float* xref_win_b,
                                             const float* curr_t_b, const float* xdes_b, float t_override, bool use_override) {
    const int t = lane;
    const bool enu = (P.flags & SDEMPC_F_FRAME_ENU) != 0;
    if (t <= P.H) {
        float row[NX];
        if (xref_win_b != nullptr && !use_override) {
            float tmp[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) tmp[i] = __ldg(xref_win_b + t * NX + i);
            if (enu) enu_ned(tmp, row);
            else {
#pragma unroll
                for (int i = 0; i < NX; ++i) row[i] = tmp[i];
            }
        } else if (curr_t_b != nullptr || use_override) {
            float tt = use_override ? t_override : __ldg(curr_t_b);
            for (int s = 0; s < t; ++s) tt = tt + P.dt[s];
            traj_interp(P.traj, P.T, tt, row);
        } else {
            float tmp[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) tmp[i] = __ldg(xdes_b + i);
            if (enu) enu_ned(tmp, row);
            else {
#pragma unroll
                for (int i = 0; i < NX; ++i) row[i] = tmp[i];
            }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) xref_dst[t * 16 + i] = row[i];
    }
}

}  // namespace sdempc
