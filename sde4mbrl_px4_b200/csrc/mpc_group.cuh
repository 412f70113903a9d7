// mpc_group.cuh — throughput variant of the solve: one warp owns GP independent problems.
//
// ncu on the per-warp kernel (profiles/r1a_*, tools/contention_probe.py) shows the FMA pipe of an SM
// sub-partition saturating with two resident warps, and ~45 % of a step's FMA-pipe instructions being the
// rigid-body / cost algebra that the per-warp kernel replicates in all 32 lanes.  Here that algebra runs
// with lane = problem (one evaluation serves GP problems), while the network phases keep lane = hidden
// unit and iterate over the warp's problems (weights stay in registers and are reused GP times).
// The APG control flow is per lane (problem): line-search trial counts, accept / reject and early stops
// diverge freely; the network loops only visit the problems that need the rollout.
// Every per-problem operation is the same SPEC-ARITH sequence as in the per-warp kernel (shared
// phys_* / mlp_* / bwd_* pieces), so results are bit-identical.  P = 1 only.
#pragma once
#include "mpc_kernels.cuh"

namespace sdempc {

template <int NU, int W, int GP>
struct Group {
    float* base;     // shared-memory region of this warp's problem 0
    int stride;      // floats per problem region
    float2* mtape;   // global activation tape of this warp: [GP][H][2][W]
    int H;
    int xstride;     // floats between the two sets of per-warp exchange buffers
    __device__ __forceinline__ float* reg(int b) const { return base + b * stride; }
    __device__ __forceinline__ float2* mt(int b, int t) const { return mtape + ((size_t)b * H + t) * 2 * W; }
};

__device__ __forceinline__ void store13(float* p, const float (&x)[NX]) {
    *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(x[4], x[5], x[6], x[7]);
    *reinterpret_cast<float4*>(p + 8) = make_float4(x[8], x[9], x[10], x[11]);
    p[12] = x[12];
}

// Forward rollouts of the problems in `mask` at the control sequences stored at offset `useq_off` of each
// problem region.  Lane b carries problem b (state, cost); lanes outside `mask` compute on stale data and
// never store.  mode 0: cost only; 1: record the adjoint tape; 2: record the state tape only.  The mode is a
// run-time argument and the network phase always processes two problems per pass (an odd one is paired with
// itself), so the kernel holds ONE copy of the forward code: the instruction footprint, not the issue rate,
// limited the first version of this kernel (58 % of stall samples were instruction-fetch).  Returns the
// cost in lane b.
template <int NU, int W, int GP>
__device__ __forceinline__ float g_rollout_fwd(const KParams& P, Warp<NU, W>& c, const Group<NU, W, GP>& gw, int useq_off,
                                            unsigned mask, int mode) {
    constexpr int NIN = 6 + NU;
    const int lane = c.lane;
    const bool mine = (lane < GP) && ((mask >> lane) & 1u);
    const bool rec = (mode == 1);
    float* reg = gw.reg(lane < GP ? lane : 0);
    const float* useq = reg + useq_off;
    float x[NX];
    load13(reg + P.o_xtape, x);   // row 0 of the state tape holds the problem's initial state for the whole solve
    float up[NU], u[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) up[i] = reg[P.o_uprev + i];
    float Jp = 0.f, disc = 1.f;
    for (int t = 0; t < P.H; ++t) {
        load_u<NU>(useq, t, u);
        {
            float z[NIN];
            phys_features<NU>(x, u, z);
            if (lane < GP) {
                float* zb = reg + P.o_zb;
#pragma unroll
                for (int i = 0; i < NIN; ++i) zb[i] = z[i];
            }
        }
        __syncwarp();
        const int ob_off = rec ? (P.o_stape + t * 20) : P.o_lz;
        for (unsigned m = mask; m;) {   // two problems per pass: independent chains hide each other's latency
            const int b0 = __ffs(m) - 1;
            m &= m - 1;
            const int b1 = m ? (__ffs(m) - 1) : b0;
            m &= m - 1;
            float* r0 = gw.reg(b0);
            float* r1 = gw.reg(b1);
            float zb[2][NIN];
#pragma unroll
            for (int i = 0; i < NIN; i += 2) {
                const float2 a = lds2(r0 + P.o_zb + i), d = lds2(r1 + P.o_zb + i);
                zb[0][i] = a.x; zb[0][i + 1] = a.y; zb[1][i] = d.x; zb[1][i + 1] = d.y;
            }
            float* const ob2[2] = {r0 + ob_off, r1 + ob_off};
            float2* const mt2[2] = {gw.mt(b0, t), gw.mt(b1, t)};
            mlp_forward_n<NU, W, -1, 2>(P, c, zb, mt2, ob2, gw.xstride, rec);
        }
        float* ob = reg + ob_off;
        float r[6], sig[6], xi[6], xr[NX], xn[NX], rn;
        load_r_sig(ob, r, sig);
        load6(reg + P.o_xi + t * 8, xi);
        load13(reg + P.o_xref + (t + 1) * 16, xr);
        const float l = phys_step<NU>(P, t, x, u, up, r, sig, xi, xr, xn, rn);
        if (rec && mine) { ob[18] = rn; ob[19] = disc; }
        if (mode != 0 && mine) store13(reg + P.o_xtape + (t + 1) * 16, xn);
        Jp = fma_(disc, l, Jp);
        disc = disc * P.discount;
#pragma unroll
        for (int i = 0; i < NU; ++i) up[i] = u[i];
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = xn[i];
    }
    __syncwarp();
    if (SDEMPC_RATE_ON(P)) {   // soft input-rate constraint of the problems in the pass (mpc_kernels.cuh)
        for (unsigned m = mask; m; m &= m - 1) {
            const int b = __ffs(m) - 1;
            const float* rb = gw.reg(b);
            const float v = rate_cost<NU>(P, lane, rb + useq_off, rb + P.o_uprev);
            if (lane == b) Jp = Jp + v;
        }
    }
    return Jp;
}

// Adjoint sweeps of the problems in `mask` (after g_rollout_fwd<MODE 1> at the same sequences): gradient of
// problem b -> its g buffer.
template <int NU, int W, int GP>
__device__ __forceinline__ void g_rollout_bwd(const KParams& P, Warp<NU, W>& c, const Group<NU, W, GP>& gw, int useq_off,
                                              unsigned mask) {
    constexpr int NIN = 6 + NU;
    const int lane = c.lane;
    const bool mine = (lane < GP) && ((mask >> lane) & 1u);
    float* reg = gw.reg(lane < GP ? lane : 0);
    const float* useq = reg + useq_off;
    float lam[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) lam[i] = 0.f;
    float gp[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) gp[i] = 0.f;
    for (int t = P.H - 1; t >= 0; --t) {
        float x[NX], u[NU], up[NU];
        BwdMid mid;
        float gu[NU];
        load13(reg + P.o_xtape + t * 16, x);
        load_u<NU>(useq, t, u);
        if (t == 0) {
#pragma unroll
            for (int i = 0; i < NU; ++i) up[i] = reg[P.o_uprev + i];
        } else load_u<NU>(useq, t - 1, up);
        {
            float xn[NX], xr[NX], r012[3], sig[6], dsg[6], rn, disc, xi[6];
            load13(reg + P.o_xtape + (t + 1) * 16, xn);
            load13(reg + P.o_xref + (t + 1) * 16, xr);
            load_step_tape(reg + P.o_stape + t * 20, r012, sig, dsg, rn, disc);
            load6(reg + P.o_xi + t * 8, xi);
            float2 lo[6];
            bwd_pre<NU>(P, t, x, xn, xr, u, r012, sig, dsg, rn, disc, xi, lam, mid, lo, gu);
            if (lane < GP) {
                float* lb = reg + P.o_lob;
#pragma unroll
                for (int i = 0; i < 6; ++i) *reinterpret_cast<float2*>(lb + 2 * i) = lo[i];
            }
        }
        __syncwarp();
        for (unsigned m = mask; m;) {
            const int b0 = __ffs(m) - 1;
            m &= m - 1;
            const int b1 = m ? (__ffs(m) - 1) : b0;
            m &= m - 1;
            float* r0 = gw.reg(b0);
            float* r1 = gw.reg(b1);
            float2 lob[2][6];
#pragma unroll
            for (int i = 0; i < 6; i += 2) {
                const float4 a = lds4(r0 + P.o_lob + 2 * i), d = lds4(r1 + P.o_lob + 2 * i);
                lob[0][i] = xy(a); lob[0][i + 1] = zw(a); lob[1][i] = xy(d); lob[1][i + 1] = zw(d);
            }
            const float2* const mt2[2] = {gw.mt(b0, t), gw.mt(b1, t)};
            float* const lz2[2] = {r0 + P.o_lz, r1 + P.o_lz};
            mlp_backward_n<NU, W, 2>(c, lob, mt2, lz2, gw.xstride);
        }
        float lz[NIN];
#pragma unroll
        for (int i = 0; i < NIN; i += 2) { const float2 a = lds2(reg + P.o_lz + i); lz[i] = a.x; lz[i + 1] = a.y; }
        bwd_post<NU>(P, x, u, up, mid, lz, gu, gp, lam);
        if (mine) {
#pragma unroll
            for (int i = 0; i < NU; ++i) reg[P.o_g + t * NU + i] = gu[i];
        }
    }
    __syncwarp();
    if (SDEMPC_RATE_ON(P)) {
        for (unsigned m = mask; m; m &= m - 1) {
            float* rb = gw.reg(__ffs(m) - 1);
            rate_grad_add<NU>(P, lane, rb + useq_off, rb + P.o_uprev, rb + P.o_g);
        }
        __syncwarp();
    }
}

// APG solves of the warp's problems (lane b = problem b).  On entry xk of every problem holds the shifted,
// clipped plan and row 0 of its state tape the initial state.  `active0`: lanes that hold a problem.
// Written as a three-state machine (gradient / line-search trial / final pass) around ONE inlined forward
// rollout and ONE inlined adjoint sweep, to keep the instruction footprint small.
template <int NU, int W, int GP>
__device__ __forceinline__ void g_apg_solve(const KParams& P, Warp<NU, W>& c, const Group<NU, W, GP>& gw,
                                            float s, bool active0, sdempc_info& inf, float* trace /* this lane's problem or null */) {
    const int lane = c.lane;
    const int n = P.H * NU;
    const unsigned FULL = 0xffffffffu;
    const unsigned all = __ballot_sync(FULL, active0);
    for (unsigned m = all; m; m &= m - 1) {
        float* rb = gw.reg(__ffs(m) - 1);
        for (int i = lane; i < n; i += 32) rb[P.o_yk + i] = rb[P.o_xk + i];
    }
    __syncwarp();
    float Jx = 0.f, Jp = 0.f, fy = 0.f, gsq = 0.f, sum_ls = 0.f, sum_s = 0.f, init_cost = 0.f, dec = 0.f;
    int k = 1, no_improve = 0, it = 0, j = 0, n_ls = 0;
    bool active = active0, need = false, ok = false;
    enum { S_GRAD = 0, S_LS = 1, S_FINAL = 2 };
    int state = S_GRAD;
    unsigned am = 0u;
    for (;;) {
        int off = P.o_xk, mode = 2;
        unsigned mask = all;
        if (state == S_GRAD) {
            am = __ballot_sync(FULL, active);
            if (am == 0u) state = S_FINAL;
            else {
                if (active) ++it;
                off = P.o_yk; mask = am; mode = 1;
            }
        }
        if (state == S_LS) {
            const unsigned nm = __ballot_sync(FULL, need);
            if (nm == 0u) {
                // ---- line search finished for every problem of this iteration: accept / reject ----
                if (active) { sum_ls = sum_ls + (float)n_ls; sum_s = sum_s + s; }
                const bool accept = active && ok && (Jp <= Jx);
                const float beta = apg_momentum(P, k);
                for (unsigned m = am; m; m &= m - 1) {
                    const int b = __ffs(m) - 1;
                    float* rb = gw.reg(b);
                    const bool ab = __shfl_sync(FULL, accept ? 1 : 0, b) != 0;
                    const float bb = __shfl_sync(FULL, beta, b);
                    if (ab) {
                        for (int i = lane; i < n; i += 32) {
                            const int ii = i % NU;
                            const float xv = rb[P.o_xp + i];
                            rb[P.o_yk + i] = clipf(fma_(bb, xv - rb[P.o_xk + i], xv), P.u_lo[ii], P.u_hi[ii]);
                            rb[P.o_xk + i] = xv;
                        }
                    } else {
                        for (int i = lane; i < n; i += 32) rb[P.o_yk + i] = rb[P.o_xk + i];
                    }
                }
                __syncwarp();
                if (active) {
                    bool converged = false;
                    if (accept) {
                        const float Jprev = Jx;
                        Jx = Jp; ++k; no_improve = 0;
                        const float tol = P.atol + P.rtol * fabsf(Jprev);
                        converged = (fabsf(Jprev - Jx) <= tol) || (Jx <= P.atol);
                    } else {
                        k = 1; ++no_improve;
                    }
                    if (trace != nullptr) {
                        float* tr = trace + (size_t)(it - 1) * SDEMPC_TRACE_W;
                        tr[0] = fy; tr[1] = Jp; tr[2] = s; tr[3] = (float)n_ls; tr[4] = accept ? 1.f : 0.f; tr[5] = Jx; tr[6] = gsq; tr[7] = (float)k;
                    }
                    if (it >= P.max_iter || no_improve >= P.max_no_improve || converged || !(fy == fy)) active = false;
                }
                state = S_GRAD;
                continue;
            }
            // ---- next trial point of every problem that still needs one ----
            for (unsigned m = nm; m; m &= m - 1) {
                const int b = __ffs(m) - 1;
                float* rb = gw.reg(b);
                const float sb = __shfl_sync(FULL, s, b);
                float part = 0.f;
                for (int i = lane; i < n; i += 32) {
                    const int ii = i % NU;
                    const float gi = rb[P.o_g + i], yi = rb[P.o_yk + i];
                    const float xv = clipf(fma_(-sb, gi, yi), P.u_lo[ii], P.u_hi[ii]);
                    rb[P.o_xp + i] = xv;
                    part = fma_(gi, xv - yi, part);
                }
                const float v = warp_butterfly(part);
                if (lane == b) dec = v;
            }
            __syncwarp();
            off = P.o_xp; mask = nm; mode = 0;
        }
        const float Jw = g_rollout_fwd<NU, W, GP>(P, c, gw, off, mask, mode);
        if (state == S_GRAD) {
            g_rollout_bwd<NU, W, GP>(P, c, gw, P.o_yk, am);
            if (active) {
                fy = Jw;   // P = 1: the particle mean is the particle (times 1/1)
                if (it == 1) { Jx = fy; init_cost = fy; }
            }
            for (unsigned m = am; m; m &= m - 1) {
                const int b = __ffs(m) - 1;
                const float* g = gw.reg(b) + P.o_g;
                float part = 0.f;
                for (int i = lane; i < n; i += 32) { const float gi = g[i]; part = fma_(gi, gi, part); }
                const float v = warp_butterfly(part);
                if (lane == b) gsq = v;
            }
            if (active && P.reset_option == 1) { s = s * P.inc_f; s = s > P.max_step ? P.max_step : s; }
            need = active; ok = false; j = 0; n_ls = 0;
            state = S_LS;
        } else if (state == S_LS) {
            if (need) {
                Jp = Jw;
                n_ls = j + 1;
                ok = (Jp <= fma_(P.coef, dec, fy));
                if (ok) need = false;
                else if (j < P.maxls) { s = s * P.dec_f; ++j; }
                else need = false;
            }
        } else {
            break;
        }
    }
    const float itf = (float)(it > 0 ? it : 1);
    inf.avg_linesearch = __fdiv_rn(sum_ls, itf);
    inf.stepsize = s;
    inf.num_steps = (float)it;
    inf.grad_sqr = gsq;
    inf.avg_stepsize = __fdiv_rn(sum_s, itf);
    inf.init_cost = init_cost;
    inf.opt_cost = (Jx == Jx) ? Jx : __int_as_float(0x7f800000);
    inf.solve_time_us = 0.f;
}

}  // namespace sdempc
