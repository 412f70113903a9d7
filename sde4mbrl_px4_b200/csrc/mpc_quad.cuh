// mpc_quad.cuh — throughput kernel, second mapping: LPP lanes per problem, 32 / LPP problems per warp.
//
// The block profile of the group kernel (profiles/README.md, r1e) shows two structural losses: the rigid-body /
// cost algebra runs once per 4 problems with 4 of 32 lanes useful (32 % of all instructions), and the network
// pass of a problem pair spends ~40 % of its instructions on exchange, addressing and the 12-lane output layer.
// Here a problem is owned by LPP (4 or 8) adjacent lanes for everything:
//   * rigid body, cost and their adjoints are replicated in the LPP lanes of a problem (one evaluation serves
//     32 / LPP problems, all of them useful);
//   * each lane owns W / LPP hidden units of BOTH networks (drift, diffusion packed in one FFMA2); weights are
//     read from a role-interleaved shared-memory image (the LPP lanes of a problem read LPP consecutive 16-byte
//     chunks, the problems of the warp read the same ones: one wavefront per LDS.128), activations are
//     exchanged through a 64-float buffer per problem (8 or 4 distinct rows per LDS.128: broadcast inside a
//     problem, conflict free between problems because region strides are 4 mod 32 floats);
//   * the output layer and the input-adjoint layer are split by rows over the LPP lanes.
// The warp's problems advance in lockstep through gradient / line-search-trial / accept phases; a problem that
// needs no further trial (or has stopped) keeps computing on stale data and stores nothing.
// Every per-problem operation is the SPEC-ARITH sequence of the other kernels (same dot-product partial sums,
// same reduction trees: the 32-lane butterfly is evaluated as in-lane adds followed by LPP-lane shuffles), so
// results are bit-identical.  P = 1 only.
#pragma once
#include "mpc_kernels.cuh"

namespace sdempc {

// Role-interleaved weight image: every row is [LPP roles][4 floats]; role r owns hidden units UPQ*r .. UPQ*r+UPQ-1,
// output rows oo*LPP + r and input rows ii*LPP + r.  A 4-float chunk holds (drift, diffusion) for two consecutive
// inputs of the contraction.
template <int NU, int W, int LPP>
struct QLayout {
    static constexpr int NIN = 6 + NU;
    static constexpr int UPQ = W / LPP;
    static constexpr int OO = (6 + LPP - 1) / LPP;
    static constexpr int II = (NIN + LPP - 1) / LPP;
    static constexpr int CH = 4 * LPP;
    static constexpr int W1 = 0;                                  // [UPQ][NIN/2] rows: W1[j][2kp], W1[j][2kp+1]
    static constexpr int B1 = W1 + UPQ * (NIN / 2) * CH;          // [UPQ/2] rows: b1[j], b1[j+1]
    static constexpr int W2 = B1 + (UPQ / 2) * CH;                // [UPQ][W/2] rows
    static constexpr int B2 = W2 + UPQ * (W / 2) * CH;            // [UPQ/2] rows
    static constexpr int W3 = B2 + (UPQ / 2) * CH;                // [OO][W/2] rows: W3[o][2kp], W3[o][2kp+1]
    static constexpr int B3 = W3 + OO * (W / 2) * CH;             // [OO] rows: (b3 drift, b3 diffusion, 0, 0)
    static constexpr int W3C = B3 + OO * CH;                      // [UPQ][3] rows: W3[2op][j], W3[2op+1][j]
    static constexpr int W2T = W3C + UPQ * 3 * CH;                // [UPQ][W/2] rows: W2[2jp][k], W2[2jp+1][k]
    static constexpr int W1T = W2T + UPQ * (W / 2) * CH;          // [II][W/2] rows: W1[2jp][i], W1[2jp+1][i]
    static constexpr int TOTAL = W1T + II * (W / 2) * CH;
    static_assert(W % LPP == 0 && UPQ % 2 == 0 && NIN % 2 == 0 && W % 2 == 0, "unsupported shape");
};

template <int NU, int W, int LPP>
struct Quad {
    using L = QLayout<NU, W, LPP>;
    static constexpr int QP = 32 / LPP;
    int lane, q, r;
    const float* wr;   // staged weight image + 4 * r
    float* reg;        // shared-memory region of this lane's problem
    float* mt;         // global activation tape of this lane's problem + 4 * r: [H][4 W] floats
    float* st;         // global step tape of this lane's problem: [H][20] floats
};

// 32-lane butterfly sum of SPEC-ARITH for a problem owned by LPP lanes: lane r holds the strided partials of
// virtual lanes r + LPP * m.  Levels 16 .. LPP pair partials of the same lane, the rest are shuffles.
template <int LPP>
__device__ __forceinline__ float quad_butterfly(float (&pm)[32 / LPP]) {
    constexpr int M = 32 / LPP;
#pragma unroll
    for (int mo = M / 2; mo >= 1; mo >>= 1)
#pragma unroll
        for (int m = 0; m < mo; ++m) pm[m] = pm[m] + pm[m + mo];
    float p = pm[0];
#pragma unroll
    for (int off = LPP / 2; off >= 1; off >>= 1) p = p + __shfl_xor_sync(0xffffffffu, p, off);
    return p;
}

__device__ __forceinline__ float4 pack4(float2 a, float2 b) { return make_float4(a.x, a.y, b.x, b.y); }

// Both networks of every problem of the warp: z (registers, replicated in the problem's lanes) -> r, sig (and
// dsg = d sigma / d s when rec).  rec: record the activations of (problem, step t) in the global tape.
template <int NU, int W, int LPP>
__device__ __forceinline__ void q_mlp_forward(const KParams& P, const Quad<NU, W, LPP>& c, const float (&z)[6 + NU], int t, bool rec,
                                              float (&r6)[6], float (&sig)[6], float (&dsg)[6]) {
    using L = QLayout<NU, W, LPP>;
    constexpr int NIN = L::NIN, UPQ = L::UPQ, OO = L::OO, CH = L::CH;
    const float2 zero = make_float2(0.f, 0.f);
    float* hA = c.reg + P.o_bufA;
    float* hB = c.reg + P.o_bufB;
    float* obs = c.reg + P.o_lz;
    float* mt = c.mt + (size_t)t * 4 * W;
    // ---- layer 1 ----
    {
        float2 h[UPQ];
#pragma unroll
        for (int jj = 0; jj < UPQ; ++jj) {
            float2 a[4];
            const float4 bv = lds4(c.wr + L::B1 + (jj / 2) * CH);
            a[0] = (jj & 1) ? zw(bv) : xy(bv);
            a[1] = a[2] = a[3] = zero;
#pragma unroll
            for (int kp = 0; kp < NIN / 2; ++kp) {
                const float4 wv = lds4(c.wr + L::W1 + (jj * (NIN / 2) + kp) * CH);
                a[(2 * kp) & 3] = fma2_(xy(wv), splat(z[2 * kp]), a[(2 * kp) & 3]);
                a[(2 * kp + 1) & 3] = fma2_(zw(wv), splat(z[2 * kp + 1]), a[(2 * kp + 1) & 3]);
            }
            h[jj] = det_tanh2(add2_(add2_(a[0], a[1]), add2_(a[2], a[3])));
        }
#pragma unroll
        for (int c2 = 0; c2 < UPQ / 2; ++c2) {
            const float4 v = pack4(h[2 * c2], h[2 * c2 + 1]);
            *reinterpret_cast<float4*>(hA + (UPQ * c.r + 2 * c2) * 2) = v;
            if (rec) *reinterpret_cast<float4*>(mt + c2 * CH) = v;
        }
    }
    __syncwarp();
    // ---- layer 2 ----
    {
        float2 a[UPQ][4];
#pragma unroll
        for (int jj = 0; jj < UPQ; ++jj) {
            const float4 bv = lds4(c.wr + L::B2 + (jj / 2) * CH);
            a[jj][0] = (jj & 1) ? zw(bv) : xy(bv);
            a[jj][1] = a[jj][2] = a[jj][3] = zero;
        }
#pragma unroll
        for (int kp = 0; kp < W / 2; ++kp) {
            const float4 hv = lds4(hA + 4 * kp);
#pragma unroll
            for (int jj = 0; jj < UPQ; ++jj) {
                const float4 wv = lds4(c.wr + L::W2 + (jj * (W / 2) + kp) * CH);
                a[jj][(2 * kp) & 3] = fma2_(xy(wv), xy(hv), a[jj][(2 * kp) & 3]);
                a[jj][(2 * kp + 1) & 3] = fma2_(zw(wv), zw(hv), a[jj][(2 * kp + 1) & 3]);
            }
        }
        float2 h[UPQ];
#pragma unroll
        for (int jj = 0; jj < UPQ; ++jj) h[jj] = det_tanh2(add2_(add2_(a[jj][0], a[jj][1]), add2_(a[jj][2], a[jj][3])));
#pragma unroll
        for (int c2 = 0; c2 < UPQ / 2; ++c2) {
            const float4 v = pack4(h[2 * c2], h[2 * c2 + 1]);
            *reinterpret_cast<float4*>(hB + (UPQ * c.r + 2 * c2) * 2) = v;
            if (rec) *reinterpret_cast<float4*>(mt + (UPQ / 2 + c2) * CH) = v;
        }
    }
    __syncwarp();
    // ---- output layer: lane r computes rows oo * LPP + r of both networks ----
    {
        float2 a[OO][4];
#pragma unroll
        for (int oo = 0; oo < OO; ++oo) {
            a[oo][0] = lds2(c.wr + L::B3 + oo * CH);
            a[oo][1] = a[oo][2] = a[oo][3] = zero;
        }
#pragma unroll
        for (int kp = 0; kp < W / 2; ++kp) {
            const float4 hv = lds4(hB + 4 * kp);
#pragma unroll
            for (int oo = 0; oo < OO; ++oo) {
                const float4 wv = lds4(c.wr + L::W3 + (oo * (W / 2) + kp) * CH);
                a[oo][(2 * kp) & 3] = fma2_(xy(wv), xy(hv), a[oo][(2 * kp) & 3]);
                a[oo][(2 * kp + 1) & 3] = fma2_(zw(wv), zw(hv), a[oo][(2 * kp + 1) & 3]);
            }
        }
        float2 out[OO];
#pragma unroll
        for (int oo = 0; oo < OO; ++oo) out[oo] = add2_(add2_(a[oo][0], a[oo][1]), add2_(a[oo][2], a[oo][3]));
        float2 sp, sg;
        det_softplus_sigmoid2(make_float2(out[0].y, out[OO - 1].y), sp, sg, rec);
#pragma unroll
        for (int oo = 0; oo < OO; ++oo) {
            const int o = oo * LPP + c.r;
            if (o < 6) {
                const float s0 = P.sig0[o];
                obs[o] = out[oo].x;
                obs[6 + o] = s0 * (oo == 0 ? sp.x : sp.y);
                if (rec) obs[12 + o] = s0 * (oo == 0 ? sg.x : sg.y);
            }
        }
    }
    __syncwarp();
    {
        const float4 a = lds4(obs), b = lds4(obs + 4), d = lds4(obs + 8);
        r6[0] = a.x; r6[1] = a.y; r6[2] = a.z; r6[3] = a.w; r6[4] = b.x; r6[5] = b.y;
        sig[0] = b.z; sig[1] = b.w; sig[2] = d.x; sig[3] = d.y; sig[4] = d.z; sig[5] = d.w;
        if (rec) {
            const float4 e = lds4(obs + 12);
            const float2 f = lds2(obs + 16);
            dsg[0] = e.x; dsg[1] = e.y; dsg[2] = e.z; dsg[3] = e.w; dsg[4] = f.x; dsg[5] = f.y;
        }
    }
}

// Forward rollouts of the warp's problems at the control sequences at offset `useq_off` of each region.
// mode 0: cost only; 1: record the adjoint tapes; 2: record the state tape only.  `mine`: this lane's problem
// takes part (the others compute on stale data and store nothing).  Returns the cost of this lane's problem.
template <int NU, int W, int LPP>
__device__ __forceinline__ float q_rollout_fwd(const KParams& P, const Quad<NU, W, LPP>& c, int useq_off, int mode, bool mine) {
    float* reg = c.reg;
    const float* useq = reg + useq_off;
    const bool rec = (mode == 1) && mine;
    float x[NX];
    load13(reg + P.o_xtape, x);   // row 0 of the state tape holds the problem's initial state for the whole solve
    float up[NU], u[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) up[i] = reg[P.o_uprev + i];
    float Jp = 0.f, disc = 1.f;
    for (int t = 0; t < P.H; ++t) {
        load_u<NU>(useq, t, u);
        float r6[6], sig[6], dsg[6];
        {
            float z[6 + NU];
            phys_features<NU>(x, u, z);
            q_mlp_forward<NU, W, LPP>(P, c, z, t, rec, r6, sig, dsg);
        }
        float xi[6], xr[NX], xn[NX], rn;
        load6(reg + P.o_xi + t * 8, xi);
        load13(reg + P.o_xref + (t + 1) * 16, xr);
        const float l = phys_step<NU>(P, t, x, u, up, r6, sig, xi, xr, xn, rn);
        if (rec) {   // step tape row: r[6] sig[6] dsg[6] rn disc, spread over the problem's lanes
            float* row = c.st + t * 20;
            if (c.r == 0) *reinterpret_cast<float4*>(row) = make_float4(r6[0], r6[1], r6[2], r6[3]);
            if (c.r == 1) *reinterpret_cast<float4*>(row + 4) = make_float4(r6[4], r6[5], sig[0], sig[1]);
            if (c.r == 2) *reinterpret_cast<float4*>(row + 8) = make_float4(sig[2], sig[3], sig[4], sig[5]);
            if (c.r == 3) *reinterpret_cast<float4*>(row + 12) = make_float4(dsg[0], dsg[1], dsg[2], dsg[3]);
            if (c.r == 0) *reinterpret_cast<float4*>(row + 16) = make_float4(dsg[4], dsg[5], rn, disc);
        }
        if (mode != 0 && mine) {
            float* row = reg + P.o_xtape + (t + 1) * 16;
            if (c.r == 0) *reinterpret_cast<float4*>(row) = make_float4(xn[0], xn[1], xn[2], xn[3]);
            if (c.r == 1) *reinterpret_cast<float4*>(row + 4) = make_float4(xn[4], xn[5], xn[6], xn[7]);
            if (c.r == 2) *reinterpret_cast<float4*>(row + 8) = make_float4(xn[8], xn[9], xn[10], xn[11]);
            if (c.r == 3) row[12] = xn[12];
        }
        Jp = fma_(disc, l, Jp);
        disc = disc * P.discount;
#pragma unroll
        for (int i = 0; i < NU; ++i) up[i] = u[i];
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = xn[i];
    }
    __syncwarp();
    return Jp;
}

// Both networks transposed: lo (registers, replicated) and this lane's recorded activations -> lz (replicated).
template <int NU, int W, int LPP>
__device__ __forceinline__ void q_mlp_backward(const KParams& P, const Quad<NU, W, LPP>& c, const float2 (&lo)[6],
                                               const float2 (&h1)[W / LPP], const float2 (&h2)[W / LPP], float (&lz)[6 + NU]) {
    using L = QLayout<NU, W, LPP>;
    constexpr int NIN = L::NIN, UPQ = L::UPQ, II = L::II, CH = L::CH;
    const float2 zero = make_float2(0.f, 0.f), one = make_float2(1.f, 1.f);
    float* dA = c.reg + P.o_bufA;
    float* dB = c.reg + P.o_bufB;
    float* lzb = c.reg + P.o_lz;
    {
        float2 d[UPQ];
#pragma unroll
        for (int jj = 0; jj < UPQ; ++jj) {
            float2 w3[6];
#pragma unroll
            for (int op = 0; op < 3; ++op) {
                const float4 wv = lds4(c.wr + L::W3C + (jj * 3 + op) * CH);
                w3[2 * op] = xy(wv); w3[2 * op + 1] = zw(wv);
            }
            float2 a0 = fma2_(w3[0], lo[0], zero);
            float2 a1 = fma2_(w3[1], lo[1], zero);
            const float2 a2 = fma2_(w3[2], lo[2], zero);
            const float2 a3 = fma2_(w3[3], lo[3], zero);
            a0 = fma2_(w3[4], lo[4], a0);
            a1 = fma2_(w3[5], lo[5], a1);
            const float2 h = h2[jj];
            d[jj] = mul2_(add2_(add2_(a0, a1), add2_(a2, a3)), fma2_(make_float2(-h.x, -h.y), h, one));
        }
#pragma unroll
        for (int c2 = 0; c2 < UPQ / 2; ++c2) *reinterpret_cast<float4*>(dA + (UPQ * c.r + 2 * c2) * 2) = pack4(d[2 * c2], d[2 * c2 + 1]);
    }
    __syncwarp();
    {
        float2 a[UPQ][4];
#pragma unroll
        for (int kk = 0; kk < UPQ; ++kk) a[kk][0] = a[kk][1] = a[kk][2] = a[kk][3] = zero;
#pragma unroll
        for (int jp = 0; jp < W / 2; ++jp) {
            const float4 dv = lds4(dA + 4 * jp);
#pragma unroll
            for (int kk = 0; kk < UPQ; ++kk) {
                const float4 wv = lds4(c.wr + L::W2T + (kk * (W / 2) + jp) * CH);
                a[kk][(2 * jp) & 3] = fma2_(xy(wv), xy(dv), a[kk][(2 * jp) & 3]);
                a[kk][(2 * jp + 1) & 3] = fma2_(zw(wv), zw(dv), a[kk][(2 * jp + 1) & 3]);
            }
        }
        float2 d[UPQ];
#pragma unroll
        for (int kk = 0; kk < UPQ; ++kk) {
            const float2 h = h1[kk];
            d[kk] = mul2_(add2_(add2_(a[kk][0], a[kk][1]), add2_(a[kk][2], a[kk][3])), fma2_(make_float2(-h.x, -h.y), h, one));
        }
#pragma unroll
        for (int c2 = 0; c2 < UPQ / 2; ++c2) *reinterpret_cast<float4*>(dB + (UPQ * c.r + 2 * c2) * 2) = pack4(d[2 * c2], d[2 * c2 + 1]);
    }
    __syncwarp();
    {
        float2 a[II][4];
#pragma unroll
        for (int ii = 0; ii < II; ++ii) a[ii][0] = a[ii][1] = a[ii][2] = a[ii][3] = zero;
#pragma unroll
        for (int jp = 0; jp < W / 2; ++jp) {
            const float4 dv = lds4(dB + 4 * jp);
#pragma unroll
            for (int ii = 0; ii < II; ++ii) {
                const float4 wv = lds4(c.wr + L::W1T + (ii * (W / 2) + jp) * CH);
                a[ii][(2 * jp) & 3] = fma2_(xy(wv), xy(dv), a[ii][(2 * jp) & 3]);
                a[ii][(2 * jp + 1) & 3] = fma2_(zw(wv), zw(dv), a[ii][(2 * jp + 1) & 3]);
            }
        }
#pragma unroll
        for (int ii = 0; ii < II; ++ii) {
            const float2 s = add2_(add2_(a[ii][0], a[ii][1]), add2_(a[ii][2], a[ii][3]));
            const int i = ii * LPP + c.r;
            if (i < NIN) lzb[i] = s.x + s.y;
        }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NIN; i += 2) { const float2 a = lds2(lzb + i); lz[i] = a.x; lz[i + 1] = a.y; }
}

// Adjoint sweeps (after q_rollout_fwd mode 1 at the same sequences): gradient of this lane's problem -> its g buffer.
template <int NU, int W, int LPP>
__device__ __forceinline__ void q_rollout_bwd(const KParams& P, const Quad<NU, W, LPP>& c, int useq_off, bool mine) {
    using L = QLayout<NU, W, LPP>;
    constexpr int NIN = 6 + NU, UPQ = L::UPQ, CH = L::CH;
    float* reg = c.reg;
    const float* useq = reg + useq_off;
    float lam[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) lam[i] = 0.f;
    float gp[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) gp[i] = 0.f;
    for (int t = P.H - 1; t >= 0; --t) {
        // tapes (global, L2 resident): issue the loads first
        float2 h1[UPQ], h2[UPQ];
        {
            const float* mt = c.mt + (size_t)t * 4 * W;
#pragma unroll
            for (int c2 = 0; c2 < UPQ / 2; ++c2) {
                const float4 a = *reinterpret_cast<const float4*>(mt + c2 * CH);
                const float4 b = *reinterpret_cast<const float4*>(mt + (UPQ / 2 + c2) * CH);
                h1[2 * c2] = xy(a); h1[2 * c2 + 1] = zw(a);
                h2[2 * c2] = xy(b); h2[2 * c2 + 1] = zw(b);
            }
        }
        float x[NX], u[NU], up[NU];
        BwdMid mid;
        float gu[NU];
        float2 lo[6];
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
            load_step_tape(c.st + t * 20, r012, sig, dsg, rn, disc);
            load6(reg + P.o_xi + t * 8, xi);
            bwd_pre<NU>(P, t, x, xn, xr, u, r012, sig, dsg, rn, disc, xi, lam, mid, lo, gu);
        }
        float lz[NIN];
        q_mlp_backward<NU, W, LPP>(P, c, lo, h1, h2, lz);
        bwd_post<NU>(P, x, u, up, mid, lz, gu, gp, lam);
        if (mine && c.r == 0) {
#pragma unroll
            for (int i = 0; i < NU; ++i) reg[P.o_g + t * NU + i] = gu[i];
        }
    }
    __syncwarp();
}

// APG solves of the warp's problems in lockstep.  On entry xk of every problem holds the shifted, clipped plan
// and row 0 of its state tape the initial state.  All state below is per problem, replicated in its LPP lanes.
template <int NU, int W, int LPP>
__device__ __forceinline__ void q_apg_solve(const KParams& P, const Quad<NU, W, LPP>& c, float s, bool active0, sdempc_info& inf,
                                            float* trace /* this lane's problem or null */) {
    constexpr int M = 32 / LPP;
    const unsigned FULL = 0xffffffffu;
    const int n = P.H * NU;
    float* reg = c.reg;
#pragma unroll
    for (int m = 0; m < M; ++m)
        for (int i = c.r + LPP * m; i < n; i += 32) reg[P.o_yk + i] = reg[P.o_xk + i];
    __syncwarp();
    float Jx = 0.f, Jp = 0.f, fy = 0.f, gsq = 0.f, sum_ls = 0.f, sum_s = 0.f, init_cost = 0.f;
    int k = 1, no_improve = 0, it = 0;
    bool active = active0;
    while (__ballot_sync(FULL, active) != 0u) {
        if (active) ++it;
        {   // gradient at y_k
            const float Jw = q_rollout_fwd<NU, W, LPP>(P, c, P.o_yk, 1, active);
            q_rollout_bwd<NU, W, LPP>(P, c, P.o_yk, active);
            if (active) {
                fy = Jw;   // P = 1: the particle mean is the particle (times 1/1)
                if (it == 1) { Jx = fy; init_cost = fy; }
            }
            float pm[M];
#pragma unroll
            for (int m = 0; m < M; ++m) {
                pm[m] = 0.f;
                for (int i = c.r + LPP * m; i < n; i += 32) { const float gi = reg[P.o_g + i]; pm[m] = fma_(gi, gi, pm[m]); }
            }
            const float v = quad_butterfly<LPP>(pm);
            if (active) gsq = v;
        }
        if (active && P.reset_option == 1) { s = s * P.inc_f; s = s > P.max_step ? P.max_step : s; }
        bool need = active, ok = false;
        int j = 0, n_ls = 0;
        while (__ballot_sync(FULL, need) != 0u) {   // one trial of every problem that still needs one
            float dec;
            {
                float pm[M];
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    pm[m] = 0.f;
                    for (int i = c.r + LPP * m; i < n; i += 32) {
                        const int ii = i % NU;
                        const float gi = reg[P.o_g + i], yi = reg[P.o_yk + i];
                        const float xv = clipf(fma_(-s, gi, yi), P.u_lo[ii], P.u_hi[ii]);
                        if (need) reg[P.o_xp + i] = xv;
                        pm[m] = fma_(gi, xv - yi, pm[m]);
                    }
                }
                dec = quad_butterfly<LPP>(pm);
            }
            __syncwarp();
            const float Jw = q_rollout_fwd<NU, W, LPP>(P, c, P.o_xp, 0, need);
            if (need) {
                Jp = Jw;
                n_ls = j + 1;
                ok = (Jp <= fma_(P.coef, dec, fy));
                if (ok) need = false;
                else if (j < P.maxls) { s = s * P.dec_f; ++j; }
                else need = false;
            }
        }
        // accept / reject
        const bool accept = active && ok && (Jp <= Jx);
        if (active) {
            sum_ls = sum_ls + (float)n_ls;
            sum_s = sum_s + s;
            const float beta = __fdiv_rn((float)k, (float)(k + 3));
#pragma unroll
            for (int m = 0; m < M; ++m)
                for (int i = c.r + LPP * m; i < n; i += 32) {
                    if (accept) {
                        const int ii = i % NU;
                        const float xv = reg[P.o_xp + i];
                        reg[P.o_yk + i] = clipf(fma_(beta, xv - reg[P.o_xk + i], xv), P.u_lo[ii], P.u_hi[ii]);
                        reg[P.o_xk + i] = xv;
                    } else {
                        reg[P.o_yk + i] = reg[P.o_xk + i];
                    }
                }
            bool converged = false;
            if (accept) {
                const float Jprev = Jx;
                Jx = Jp; ++k; no_improve = 0;
                const float tol = P.atol + P.rtol * fabsf(Jprev);
                converged = (fabsf(Jprev - Jx) <= tol) || (Jx <= P.atol);
            } else {
                k = 1; ++no_improve;
            }
            if (trace != nullptr && c.r == 0) {
                float* tr = trace + (size_t)(it - 1) * SDEMPC_TRACE_W;
                tr[0] = fy; tr[1] = Jp; tr[2] = s; tr[3] = (float)n_ls; tr[4] = accept ? 1.f : 0.f; tr[5] = Jx; tr[6] = gsq; tr[7] = (float)k;
            }
            if (it >= P.max_iter || no_improve >= P.max_no_improve || converged || !(fy == fy)) active = false;
        }
        __syncwarp();
    }
    (void)q_rollout_fwd<NU, W, LPP>(P, c, P.o_xk, 2, active0);
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
