// mpc_tcsolve.cuh — the batched APG solve on the tensor-core mapping (tcgen05 + TMEM), SDEMPC_F_TENSOR.
//
// Serves m_mpc (/root/reference sde4mbrl_px4/mpc_controller/sde_control.py:400-416) for BATCHES of problems, where
// rows = rollouts is a real dense contraction (BASELINE configs 3 and 4).  Same arithmetic as mpc_tc.cuh (TF32
// operands, fp32 accumulation, tanh.approx), therefore NOT SPEC-ARITH: compared with the oracle through its
// teacher-forced mode and at cost level (tests/test_gpu_parity.py), never bit for bit.
//
// Mapping.  One CTA of 128 threads = 128 TMEM lanes = 128 concurrent rollouts ("slots"), owning PPC <= 128 / P
// problems (P = particles; the particles of a rollout are adjacent lanes).  The three kinds of rollout an APG
// iteration needs are all "tasks" dealt onto the slots, and the CTA runs them in lockstep through ONE copy of the
// forward step (three dense contractions per step on the tensor cores, operands in tensor memory, weights staged
// once in shared memory by a TMA bulk copy):
//   GRAD   slot (q, p): value_and_grad at y_k of problem q, particle p (records the adjoint tape, then the sweep);
//   LS     COMPACTED CONCURRENT LINE SEARCH: the problems that still need a trial share the 128 / P slots; every
//          round evaluates m = floor(slots / needy) consecutive Armijo trials per needy problem at once (trial j uses
//          s * dec^j exactly as the sequential search would), and each problem then takes the first passing trial in
//          order — the same selection as the sequential search.  A CTA of 25 problems runs all 5 trials of
//          iris_traj.yaml in one pass; a CTA of 128 problems runs trial 0 for everyone, then trials 1..3 for the
//          ~30 % that failed, then trial 4 for the rest: 3 passes instead of the 5 a lockstep sequential search costs;
//   FINAL  slot (q, p): forward pass at u* for the predicted mean trajectory.
// APG state.  Plans (x_k, y_k), gradient, reference window, noise and initial state live in a per-CTA global
// workspace laid out [entry][problem] (every access of lane = problem coalesces; it stays in L2); the scalars of
// problem q (step size, costs, counters) live in shared memory and are updated by thread q between phases.
// Noise (Philox + Box-Muller) and the reference window are produced ONCE per solve, not once per rollout.
#pragma once
#include "mpc_tc.cuh"

namespace sdempc {

// per-CTA workspace, float offsets; problem-indexed arrays have the constant row stride TCS_RS = 128 (one slot per
// possible problem of a CTA), so that every access is base + compile-time offset (with a run-time stride a quarter of the
// kernel's instructions were address arithmetic); rows beyond the CTA's problems are never touched
constexpr int TCS_RS = 128;
struct TCSWs {
    int RS, n, o_xk, o_yk, o_g, o_uprev, o_x0, o_xref, o_xi, o_xh, o_tape, total;
};
__host__ __device__ inline TCSWs tcs_ws_layout(int H, int NU, int RS, int TG) {
    TCSWs w;
    w.RS = RS; w.n = H * NU;
    int o = 0;
    w.o_xk = o; o += w.n * RS;
    w.o_yk = o; o += w.n * RS;
    w.o_g = o; o += w.n * RS;
    w.o_uprev = o; o += SDEMPC_MAX_NU * RS;
    w.o_x0 = o; o += 16 * RS;
    w.o_xref = o; o += (H + 1) * NX * RS;
    w.o_xi = o; o += H * 6 * 128;                 // indexed by slot row (q * P + p) < 128
    w.o_xh = o; o += 16 * 128;                    // x_H of the rows that recorded a tape (speculative build), by slot row
    w.o_tape = o; o += H * TG * 128 * 4;          // adjoint tape of the GRAD rows: [step][granule][row] float4
    w.total = o;
    return w;
}

template <int NU, int W>
struct TCSLayout : TCLayout<NU, W> {
    using L = TCLayout<NU, W>;
    // adjoint tape granules (16 bytes each) per row and step: h1, h2 as fp16 (N12 / 8 granules each), then 8 granules
    // of floats: r[0..2], sigma(6), dsigma/ds(6), 1/|q|, discount^t, x_t(13)
    static constexpr int S_H1 = 0, S_H2 = L::N12 / 8, S_ST = 2 * L::N12 / 8, STG = S_ST + 8;
};

// owner state of the CTA's problems (thread q updates problem q between phases) + the task list of a phase
struct TCSShared {
    float s[128], Jx[128], fy[128], gsq[128], sum_ls[128], sum_s[128], init_cost[128], Jp[128];
    int k[128], no_imp[128], it[128], n_ls[128], jn[128], first[128], cnt[128];
    unsigned char active[128], need[128], ok[128], acc[128];
    int wsum[4];
    float t_J[128], t_dec[128];
    short t_q[128], t_j[128];
    int ntasks;
    // speculative build (SPECG): phase of every problem, kind of every task, the slot row that holds a problem's tape
    float fy_next[128];
    short trow[128], b_q[128], scnt[128], firstT[128];
    int wsumL[4];
    unsigned char ph[128], t_kind[128], smask[128];
};

namespace tc {
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// softplus and its derivative (sigmoid) on the special-function unit: ~1e-6 relative, same class as tanh.approx
__device__ __forceinline__ void softplus_sigmoid_fast(float s, float& sp, float& sg) {
    const float e = ex2_approx(-fabsf(s) * 1.4426950408889634f);        // exp(-|s|) in (0, 1]
    const float r = rcp_approx(1.0f + e);
    sp = fmaxf(s, 0.f) + lg2_approx(1.0f + e) * 0.6931471805599453f;
    sg = s >= 0.f ? r : e * r;
}
}  // namespace tc

// Body of the solve kernel.  sB: weight image (TCLayout, forward + adjoint parts).  bars[0]: weight staging, bars[1]: MMA.
// SPECG = true: the build for CTAs with few problems (PPC <= slots / 4), where most slots would idle.
//   (1) SPECULATIVE GRADIENT PASSES.  Next to the line-search trials the free slots run the forward pass of the NEXT gradient
//       evaluation for the likely outcomes, recording a tape each: "R" at x_k, where a rejected step restarts from (30 % of
//       the iterations of the benchmark problems end that way, and x_k is known before the search), and one per trial j at
//       the y_{k+1} that trial gives if accepted (the expressions of plan_update, hence the same bits).  When the outcome has
//       a tape, the iteration's gradient needs only the adjoint sweep: two passes per iteration instead of three.
//   (2) PER-PROBLEM PHASES.  A CTA-wide state would make all problems pay for the one whose outcome had no tape.  So each
//       problem carries its own phase (0 needs the gradient's forward pass, 1 line search, 2 tape ready, 3 done); every loop
//       turn runs one forward pass over whatever forward tasks exist (gradient passes, trials, speculations, mixed) and
//       then one adjoint sweep over the problems whose tape is ready, each in the slot that recorded it.  A miss costs that
//       problem one extra turn, nobody else.
//   Problems are independent, every task evaluates the same expressions on the same operands as in the other build, and
//   a tensor-memory lane's result does not depend on which lane it is: the two builds return identical bits
//   (tests/test_gpu_parity.py::test_tensor_core_solve_speculative_build_returns_the_same_bits).
//   Measured (iris, 200 iterations, one CTA per SM): 7 problems per CTA 35.8 -> 25.0 ms, 14 per CTA 34.1 -> 25.5 ms, 28 per CTA
//   (two trials + R + one speculation each) 34.6 -> 33.5 ms; width 64: 28 per CTA 59.4 -> 51.0 ms, 4 problems x 8 particles
//   59.6 -> 49.8 ms.  (Before the forward pass loaded its operands a step ahead into registers — PFR below — the build LOST at
//   28 per CTA, 36.8 against 35.9 ms: four warps of rows with distinct operands exposed an L1 / L2 round trip per step.)
// PFR = true (the builds with the register budget of two CTAs per SM): the operands of step t + 1 (controls, gradient, x_k,
// noise, reference row: 31-35 values per row) are loaded into registers at the top of step t — a full step ahead of their
// use — instead of being prefetched towards L1 and loaded when needed (which still exposes an L1 / L2 round trip per step).
// RATE = true: the build that evaluates the soft input-rate constraint (u_slew_constr; a run-time branch for it in every build
// costs the four-CTA build 6 % — 32 more bytes of spill — for a feature one reference configuration uses).
template <int NU, int W, bool SPECG, bool PFR, bool RATE>
__device__ __forceinline__ void tc_solve_body(const KParams& P, unsigned char* sB, uint32_t* tmem_base_slot, uint64_t* bars,
                                              TCSShared& sh) {
    using L = TCSLayout<NU, W>;
    constexpr int NIN = L::NIN, N12 = L::N12;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* wbar = bars;
    uint64_t* bar = bars + 1;
    // ---- one-time setup: weights by TMA bulk copy, barriers, tensor memory ----
    if (tid == 0) {
        tc::mbar_init(wbar, 1);
        tc::mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        constexpr uint32_t BYTES = L::BYTES_GRAD, CH = 32768;
        tc::mbar_expect_tx(wbar, BYTES);
        for (uint32_t off = 0; off < BYTES; off += CH)
            tc::bulk_g2s(sB + off, reinterpret_cast<const unsigned char*>(P.wimg) + off, (BYTES - off) < CH ? (BYTES - off) : CH, wbar);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_base_slot)), "r"((uint32_t)L::COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = *tmem_base_slot;
    const uint32_t lane_addr = tb + ((uint32_t)(warp * 32) << 16);
    const uint32_t sb = tc::smem_u32(sB);
    const uint32_t id12 = tc::idesc_tf32(128, N12), id3 = tc::idesc_tf32(128, L::N3), idW = tc::idesc_tf32(128, W), id16 = tc::idesc_tf32(128, 16);
    uint32_t phase = 0;

    const int PP = P.P, NT = 128 / PP, PPC = P.tcs_ppc;
    const int b0 = blockIdx.x * PPC;
    const int nq = (P.B - b0) < PPC ? (P.B - b0) : PPC;
    const TCSWs ws = tcs_ws_layout(P.H, NU, P.tcs_rs, L::STG);
    float* wsb = P.tcs_ws + (size_t)blockIdx.x * ws.total;
    float* XK = wsb + ws.o_xk;
    float* YK = wsb + ws.o_yk;
    float* G = wsb + ws.o_g;
    float* UPREV = wsb + ws.o_uprev;
    float* X0 = wsb + ws.o_x0;
    float* XREF = wsb + ws.o_xref;
    float* XI = wsb + ws.o_xi;
    [[maybe_unused]] float* XH = wsb + ws.o_xh + tid;
    float4* tape = reinterpret_cast<float4*>(wsb + ws.o_tape) + tid;
    auto tp = [&](int t, int g) -> float4* { return tape + ((size_t)t * L::STG + g) * 128; };
    // the tape is written once and read once per iteration: streaming stores / loads (evict first), so that it does not
    // push the plans, the reference window and the noise of the resident CTAs out of L2
    const bool lone_cta = gridDim.x <= (unsigned)P.tcs_sms;   // one CTA per SM: the next step's tape fits L1, prefetch that far
    // (the speculative build records several tapes per problem and reads one: streaming there too)
    const bool lone_tape = lone_cta && !SPECG;
    auto stt = [&](int t, int g, float4 v) { if (lone_tape) *tp(t, g) = v; else __stcs(tp(t, g), v); };
    constexpr int RS = TCS_RS;
    const int n = ws.n;
    const bool enu = (P.flags & SDEMPC_F_FRAME_ENU) != 0;
    const float invP = __fdiv_rn(1.0f, (float)PP);
    const int pbase = lane & ~(PP - 1);
    auto pmean = [&](float v) -> float {
        if (PP == 1) return v;
        float s = __shfl_sync(0xffffffffu, v, pbase);
        for (int p = 1; p < PP; ++p) s = s + __shfl_sync(0xffffffffu, v, pbase + p);
        return s * invP;
    };

    // ---- per-solve preparation: state, plans, reference window (internal frame), noise ----
    if (tid < nq) {
        const int q = tid, b = b0 + q;
        float tmp[NX], x0[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) tmp[i] = __ldg(P.x + (size_t)b * NX + i);
        if (enu) enu_ned(tmp, x0);
        else {
#pragma unroll
            for (int i = 0; i < NX; ++i) x0[i] = tmp[i];
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) X0[i * RS + q] = x0[i];
        const float* pin = P.u_plan + (size_t)b * n;
#pragma unroll
        for (int i = 0; i < NU; ++i) UPREV[i * RS + q] = __ldg(pin + i);
        for (int t = 0; t < P.H; ++t) {
            const int ts = (P.flags & SDEMPC_F_NO_SHIFT) ? t : (t + 1 < P.H ? t + 1 : P.H - 1);
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                const float v = clipf(__ldg(pin + ts * NU + i), P.u_lo[i], P.u_hi[i]);
                XK[(t * NU + i) * RS + q] = v;
                YK[(t * NU + i) * RS + q] = v;
            }
        }
        const float s0 = P.info[b].stepsize;
        sh.s[q] = s0 > 0.f ? s0 : P.init_step;
        sh.Jx[q] = 0.f; sh.fy[q] = 0.f; sh.gsq[q] = 0.f; sh.sum_ls[q] = 0.f; sh.sum_s[q] = 0.f; sh.init_cost[q] = 0.f; sh.Jp[q] = 0.f;
        sh.k[q] = 1; sh.no_imp[q] = 0; sh.it[q] = 0; sh.n_ls[q] = 0; sh.jn[q] = 0;
        sh.active[q] = 1; sh.need[q] = 0; sh.ok[q] = 0;
        sh.ph[q] = 0;
    } else {
        sh.active[tid] = 0; sh.need[tid] = 0;
        sh.ph[tid] = 3;
    }
    for (int idx = tid; idx < nq * (P.H + 1); idx += 128) {
        const int q = idx / (P.H + 1), t = idx - q * (P.H + 1);
        float row[NX];
        tc_ref_row(P, b0 + q, t, row);
#pragma unroll
        for (int i = 0; i < NX; ++i) XREF[(t * NX + i) * RS + q] = row[i];
    }
    for (int idx = tid; idx < nq * PP * P.H; idx += 128) {
        const int r = idx % (nq * PP), t = idx / (nq * PP);
        const int q = r / PP, p = r - q * PP, b = b0 + q;
        float xi[6];
        if (P.xi_override != nullptr) {
#pragma unroll
            for (int i = 0; i < 6; ++i) xi[i] = __ldg(P.xi_override + (((size_t)b * PP + p) * P.H + t) * 6 + i);
        } else {
            const unsigned long long seed = P.rng[2 * (size_t)b], tick = P.rng[2 * (size_t)b + 1];
            const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
            const uint32_t c2 = (uint32_t)tick, c3 = ((uint32_t)(tick >> 32)) & 0x3FFFFFFFu;
            uint32_t w4[4];
            philox4x32_10((uint32_t)t, (uint32_t)p, c2, c3, k0, k1, w4);
            box_muller(w4[0], w4[1], xi[0], xi[1]);
            box_muller(w4[2], w4[3], xi[2], xi[3]);
            philox4x32_10((uint32_t)t, (uint32_t)p, c2, c3 | (1u << 30), k0, k1, w4);
            box_muller(w4[0], w4[1], xi[4], xi[5]);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) XI[(t * 6 + i) * 128 + r] = xi[i];
    }
    tc::mbar_wait_plain(wbar, 0);       // weight image landed (async proxy -> visible to tcgen05.mma and to ld.shared after the wait)
    __syncthreads();
    const float* bias2 = reinterpret_cast<const float*>(sB + L::BIAS2);
    const float* bias3 = reinterpret_cast<const float*>(sB + L::BIAS3);

    // plan update after a resolved line search (acc: 1 accept, 0 reject = restart from x_k, 2 untouched), all threads
    auto plan_update = [&]() {
        // plans: thread -> (problem oq = tid mod 2^k, part): consecutive threads touch consecutive problems (coalesced),
        // a problem's entries are split over 128 / 2^k threads; four independent entries are in flight per thread
        int npad = 1;
        while (npad < nq) npad <<= 1;
        const int oq = tid & (npad - 1), part = tid / npad, nparts = 128 / npad;
        const int a = oq < nq ? (int)sh.acc[oq] : 2;
        if (a != 2) {
            const float sq = sh.s[oq];
            const int kq = sh.k[oq];
            const float beta = apg_momentum_tab(P, kq);
            for (int i0 = part; i0 < n; i0 += 4 * nparts) {
                float xkv[4], yv[4], gv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = i0 + j * nparts;
                    if (i < n) {
                        xkv[j] = XK[i * RS + oq];
                        if (a == 1) { yv[j] = YK[i * RS + oq]; gv[j] = G[i * RS + oq]; }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = i0 + j * nparts;
                    if (i < n) {
                        const int ii = i % NU;
                        if (a == 1) {
                            const float xv = clipf(fma_(-sq, gv[j], yv[j]), P.u_lo[ii], P.u_hi[ii]);
                            YK[i * RS + oq] = clipf(fma_(beta, xv - xkv[j], xv), P.u_lo[ii], P.u_hi[ii]);
                            XK[i * RS + oq] = xv;
                        } else {
                            YK[i * RS + oq] = xkv[j];
                        }
                    }
                }
            }
        }
    };
    // owner bookkeeping of problem oq after its line search: accept / reject, momentum counter, trace, stop tests
    auto bookkeeping = [&](int oq) {
        const float s = sh.s[oq], Jp = sh.Jp[oq], Jx0 = sh.Jx[oq];
        sh.sum_ls[oq] += (float)sh.n_ls[oq];
        sh.sum_s[oq] += s;
        const bool accept = sh.acc[oq] == 1;
        bool converged = false;
        if (accept) {
            sh.Jx[oq] = Jp; sh.k[oq] += 1; sh.no_imp[oq] = 0;
            const float tol = P.atol + P.rtol * fabsf(Jx0);
            converged = (fabsf(Jx0 - Jp) <= tol) || (Jp <= P.atol);
        } else {
            sh.k[oq] = 1; sh.no_imp[oq] += 1;
        }
        const int it = sh.it[oq];
        if (P.trace != nullptr) {
            float* tr = P.trace + ((size_t)(b0 + oq) * P.max_iter + (it - 1)) * SDEMPC_TRACE_W;
            tr[0] = sh.fy[oq]; tr[1] = Jp; tr[2] = s; tr[3] = (float)sh.n_ls[oq]; tr[4] = accept ? 1.f : 0.f; tr[5] = sh.Jx[oq];
            tr[6] = sh.gsq[oq]; tr[7] = (float)sh.k[oq];
        }
        const float fyv = sh.fy[oq];
        if (it >= P.max_iter || sh.no_imp[oq] >= P.max_no_improve || converged || !(fyv == fyv)) sh.active[oq] = 0;
    };

    enum { S_GRAD = 0, S_LS = 1, S_FINAL = 2 };
    int state = S_GRAD;
    for (;;) {
        // ================= choose this pass's task for my slot =================
        int q = 0, pidx = tid % PP, mode = 2;
        bool valid = false;
        float s_t = 0.f;
        [[maybe_unused]] bool spec = false, tape_on = false, at_xk = false;
        [[maybe_unused]] float beta_t = 0.f;
        if constexpr (SPECG) {
            // forward tasks of this turn: one gradient forward pass per problem in phase 0; trials (+ speculations) of the
            // problems in phase 1 on the remaining slots
            const int ph = tid < nq ? (int)sh.ph[tid] : 3;
            const int n_gf = __syncthreads_count(ph == 0 ? 1 : 0);
            const int cnt = __syncthreads_count(ph == 1 ? 1 : 0);
            if (n_gf + cnt == 0) state = S_FINAL;          // every problem is done
            else {
                // A problem in its line search gets m slots: trials, and tape tasks for what may follow them — "R", the gradient
                // pass at x_k that a REJECTED step restarts from (known before the search: 30 % of the iterations end that way),
                // and speculations for the likeliest accepted trials (from the start of a search: trial 1 in 61 % of the
                // iterations, trial 0 in 5 %, trial 2 in 4 %)
                int mq = 0, nt = ph == 0 ? 1 : 0;
                unsigned smask = 0;                      // bit i: trial j0 + i has a speculation
                if (ph == 1) {
                    const int m = (NT - n_gf) / cnt, j0 = sh.jn[tid], rem = P.maxls + 1 - j0;
                    if (m < 4) mq = m < rem ? m : rem;                                // crowded: trials only
                    else {
                        mq = (m - m / 2) < rem ? (m - m / 2) : rem;
                        nt = (m - mq) < (mq + 1) ? (m - mq) : (mq + 1);
                        const bool from_start = j0 == 0 && mq >= 2;
                        for (int r = 0; r < nt - 1; ++r) smask |= 1u << (from_start ? (r == 0 ? 1 : r == 1 ? 0 : r) : r);
                    }
                }
                // slots: first every task that records a tape (gradient forward passes, speculations), then the trials — the
                // tape rows in use stay contiguous (scattered rows touch every line of the tape region: 4x the L2 footprint)
                int incT = nt, incL = mq;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int vT = __shfl_up_sync(0xffffffffu, incT, off), vL = __shfl_up_sync(0xffffffffu, incL, off);
                    if (lane >= off) { incT += vT; incL += vL; }
                }
                if (lane == 31) { sh.wsum[warp] = incT; sh.wsumL[warp] = incL; }
                __syncthreads();
                int baseT = 0, baseL = 0, totT = 0;
                for (int w = 0; w < 4; ++w) {
                    if (w < warp) { baseT += sh.wsum[w]; baseL += sh.wsumL[w]; }
                    totT += sh.wsum[w];
                }
                const int firstT = baseT + incT - nt, firstL = totT + baseL + incL - mq;
                if (ph == 0) {
                    sh.firstT[tid] = (short)firstT;
                    sh.t_q[firstT] = (short)tid; sh.t_j[firstT] = 0; sh.t_kind[firstT] = 2;
                } else if (ph == 1) {
                    sh.first[tid] = firstL; sh.cnt[tid] = mq; sh.firstT[tid] = (short)firstT; sh.scnt[tid] = (short)nt; sh.smask[tid] = (unsigned char)smask;
                    const int j0 = sh.jn[tid];
                    for (int j = 0; j < mq; ++j) { sh.t_q[firstL + j] = (short)tid; sh.t_j[firstL + j] = (short)(j0 + j); sh.t_kind[firstL + j] = 0; }
                    if (nt > 0) { sh.t_q[firstT] = (short)tid; sh.t_j[firstT] = 0; sh.t_kind[firstT] = 3; }
                    for (int j = 0, r = 1; j < mq; ++j)
                        if ((smask >> j) & 1u) { sh.t_q[firstT + r] = (short)tid; sh.t_j[firstT + r] = (short)(j0 + j); sh.t_kind[firstT + r] = 1; ++r; }
                }
                if (tid == 127) sh.ntasks = firstL + mq;
                __syncthreads();
                const int kslot = tid / PP;
                valid = kslot < sh.ntasks;
                q = valid ? (int)sh.t_q[kslot] : 0;
                const int kind = valid ? (int)sh.t_kind[kslot] : 0;
                mode = kind >= 2 ? 1 : 0;                // 2: gradient pass at y_k, 3: at x_k (restart), 1: speculation, 0: trial
                spec = kind == 1;
                at_xk = kind == 3;
                tape_on = valid && kind != 0;
                s_t = sh.s[q];
                const int dj = (valid && kind < 2) ? (int)sh.t_j[kslot] - sh.jn[q] : 0;
                for (int i = 0; i < dj; ++i) s_t = s_t * P.dec_f;
                beta_t = apg_momentum_tab(P, sh.k[q]);
            }
        }
        if (!SPECG && state == S_GRAD) {
            const int myq = tid / PP;
            const bool act = myq < nq && sh.active[myq];
            if (__syncthreads_or(act ? 1 : 0) == 0) { state = S_FINAL; }
            else { q = myq < nq ? myq : 0; valid = act; mode = 1; }
        }
        if (!SPECG && state == S_LS) {
            // compacted concurrent line search: deal the trials of the needy problems onto the slots (thread q = problem q)
            const bool nd = tid < nq && sh.need[tid];
            const int cnt = __syncthreads_count(nd ? 1 : 0);
            if (cnt == 0) {
                // ---- line search finished for every problem: accept / reject, momentum, stop tests ----
                if (tid < nq && sh.active[tid]) {
                    const float Jp = sh.Jp[tid];
                    sh.acc[tid] = (sh.ok[tid] && (Jp <= sh.Jx[tid])) ? 1 : 0;
                } else if (tid < 128) sh.acc[tid] = 2;     // 2: not an active problem, plans untouched
                __syncthreads();
                plan_update();
                __syncthreads();   // sh.k is read above, updated below
                if (tid < nq && sh.active[tid]) bookkeeping(tid);
                __syncthreads();   // plans and flags visible to the rows of the next pass
                state = S_GRAD;
                continue;
            }
            {
                int mq = 0;
                if (nd) {
                    const int m = NT / cnt > 0 ? NT / cnt : 1;
                    const int rem = P.maxls + 1 - sh.jn[tid];
                    mq = m < rem ? m : rem;
                }
                // exclusive prefix sum of mq over the CTA: warp scan, then the four warp totals through shared memory
                int inc = mq;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += v; }
                if (lane == 31) sh.wsum[warp] = inc;
                __syncthreads();
                int base = 0;
                for (int w = 0; w < warp; ++w) base += sh.wsum[w];
                const int first = base + inc - mq;
                if (nd) {
                    sh.first[tid] = first; sh.cnt[tid] = mq;
                    const int j0 = sh.jn[tid];
                    for (int j = 0; j < mq; ++j) { sh.t_q[first + j] = (short)tid; sh.t_j[first + j] = (short)(j0 + j); }
                }
                if (tid == 127) sh.ntasks = first + mq;
                __syncthreads();
            }
            const int kslot = tid / PP;
            valid = kslot < sh.ntasks;
            q = valid ? (int)sh.t_q[kslot] : 0;
            mode = 0;
            s_t = sh.s[q];
            const int dj = valid ? (int)sh.t_j[kslot] - sh.jn[q] : 0;
            for (int i = 0; i < dj; ++i) s_t = s_t * P.dec_f;
        }
        if (state == S_FINAL) {
            const int myq = tid / PP;
            valid = myq < nq;
            q = valid ? myq : 0;
            mode = 2;
        }
        // ================= forward rollout of my task (all 128 threads in lockstep) =================
        const float* useq = (mode == 2 || (SPECG && at_xk)) ? XK : YK;
        int xi_row = q * PP + pidx;
        const bool tape_w = SPECG ? tape_on : (mode == 1 && valid);
        float x[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = X0[i * RS + q];
        float up[NU], u[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) up[i] = UPREV[i * RS + q];
        const int b = b0 + q;
        if (mode == 2 && valid && pidx == 0) {
#pragma unroll
            for (int i = 0; i < NX; ++i) P.x_evol[(size_t)b * (P.H + 1) * NX + i] = __ldg(P.x + (size_t)b * NX + i);
        }
        float Jp = 0.f, disc = 1.f, dec = 0.f;
        const float* yq = useq + q;
        const float* gq = G + q;
        [[maybe_unused]] const float* xkq = XK + q;
        const float* xiq = XI + xi_row;
        const float* xrq = XREF + q;
        // operands of one step of my row
        float o_y[NU], o_g[NU], o_xk[NU], o_xi[6], o_xr[NX];
        auto load_ops = [&](int t) {
            const float* yt = yq + t * (NU * RS);
            const float* gt = gq + t * (NU * RS);
            const float* xit = xiq + t * (6 * 128);
            const float* xrt = xrq + (t + 1) * (NX * RS);
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                o_y[i] = yt[i * RS];
                o_g[i] = (mode == 0) ? gt[i * RS] : 0.f;
                if constexpr (SPECG) o_xk[i] = spec ? xkq[(t * NU + i) * RS] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) o_xi[i] = xit[i * 128];
#pragma unroll
            for (int i = 0; i < NX; ++i) o_xr[i] = xrt[i * RS];
        };
        if constexpr (PFR) load_ops(0);
        for (int t = 0; t < P.H; ++t) {
            if constexpr (!PFR) load_ops(t);
            float xi[6], xr[NX];
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                const float yv = o_y[i];
                if (mode == 0) {
                    const float gv = o_g[i];
                    const float xv = clipf(fma_(-s_t, gv, yv), P.u_lo[i], P.u_hi[i]);
                    dec = fma_(gv, xv - yv, dec);
                    u[i] = xv;
                    if constexpr (SPECG) {   // the y_{k+1} this trial gives if accepted: the expression of plan_update
                        if (spec) u[i] = clipf(fma_(beta_t, xv - o_xk[i], xv), P.u_lo[i], P.u_hi[i]);
                    }
                } else {
                    u[i] = yv;
                }
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) xi[i] = o_xi[i];
#pragma unroll
            for (int i = 0; i < NX; ++i) xr[i] = o_xr[i];
            if constexpr (PFR) {
                if (t + 1 < P.H) load_ops(t + 1);   // in flight during this step's contractions
            } else if (lone_cta && t + 1 < P.H) {   // next step's operands towards L1 while this step's contractions run (CTA alone on its SM)
                const float* yt = yq + t * (NU * RS);
                const float* gt = gq + t * (NU * RS);
                const float* xit = xiq + t * (6 * 128);
                const float* xrt = xrq + (t + 1) * (NX * RS);
#pragma unroll
                for (int i = 0; i < NU; ++i) {
                    tc::prefetch_l1(yt + (NU + i) * RS);
                    if (mode == 0) tc::prefetch_l1(gt + (NU + i) * RS);
                    if constexpr (SPECG) { if (spec) tc::prefetch_l1(xkq + ((t + 1) * NU + i) * RS); }
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) tc::prefetch_l1(xit + (6 + i) * 128);
#pragma unroll
                for (int i = 0; i < NX; ++i) tc::prefetch_l1(xrt + (NX + i) * RS);
            }
            // ---- layer 1 operand: [z, 1, 0 ...] ----
            {
                float z[NIN];
                phys_features<NU>(x, u, z);
#pragma unroll
                for (int c0 = 0; c0 < L::K1; c0 += 8) {
                    float a[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[i] = (c0 + i < NIN) ? z[(c0 + i < NIN) ? c0 + i : 0] : (c0 + i == NIN ? 1.f : 0.f);
                    tc::st8(lane_addr + L::C_A + c0, a);
                }
            }
            tc::publish();
            if (tid == 0) {
#pragma unroll
                for (int k8 = 0; k8 < L::K1 / 8; ++k8)
                    tc::mma_ts(tb + L::C_D12, tb + L::C_A + 8 * k8, tc::desc(sb + L::B1 + k8 * 2 * L::LBO, L::SBO1), id12, k8 > 0);
                tc::commit(bar);
            }
            tc::wait(bar, phase); phase ^= 1;
#pragma unroll
            for (int c0 = 0; c0 < N12; c0 += 16) {
                float v[16];
                tc::ld16(lane_addr + L::C_D12 + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = tc::tanh_approx(v[i]);
                tc::st16(lane_addr + L::C_A + c0, v);
                if (tape_w) {
                    stt(t, L::S_H1 + c0 / 8, tc::pack8(v));
                    stt(t, L::S_H1 + c0 / 8 + 1, tc::pack8(v + 8));
                }
            }
            tc::publish();
            if (tid == 0) {
#pragma unroll
                for (int nn = 0; nn < 2; ++nn)
#pragma unroll
                    for (int k8 = 0; k8 < W / 8; ++k8)
                        tc::mma_ts(tb + L::C_D12 + nn * W, tb + L::C_A + nn * W + 8 * k8,
                                   tc::desc(sb + L::B2 + nn * L::NET2 + k8 * 2 * L::LBO, L::SBOW), idW, k8 > 0);
                tc::commit(bar);
            }
            tc::wait(bar, phase); phase ^= 1;
#pragma unroll
            for (int c0 = 0; c0 < N12; c0 += 16) {
                float v[16];
                tc::ld16(lane_addr + L::C_D12 + c0, v);
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 bv = lds4(bias2 + c0 + i);
                    v[i] = tc::tanh_approx(v[i] + bv.x); v[i + 1] = tc::tanh_approx(v[i + 1] + bv.y);
                    v[i + 2] = tc::tanh_approx(v[i + 2] + bv.z); v[i + 3] = tc::tanh_approx(v[i + 3] + bv.w);
                }
                tc::st16(lane_addr + L::C_A + c0, v);
                if (tape_w) {
                    stt(t, L::S_H2 + c0 / 8, tc::pack8(v));
                    stt(t, L::S_H2 + c0 / 8 + 1, tc::pack8(v + 8));
                }
            }
            tc::publish();
            if (tid == 0) {
#pragma unroll
                for (int k8 = 0; k8 < L::K2 / 8; ++k8)
                    tc::mma_ts(tb + L::C_D3, tb + L::C_A + 8 * k8, tc::desc(sb + L::B3 + k8 * 2 * L::LBO, L::SBO2), id3, k8 > 0);
                tc::commit(bar);
            }
            tc::wait(bar, phase); phase ^= 1;
            float r6[6], sig[6], dsg[6];
            {
                float o[16];
                tc::ld16(lane_addr + L::C_D3, o);
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    r6[i] = o[i] + bias3[i];
                    float sp, sg;
                    tc::softplus_sigmoid_fast(o[6 + i] + bias3[6 + i], sp, sg);
                    sig[i] = P.sig0[i] * sp;
                    dsg[i] = P.sig0[i] * sg;
                }
            }
            float xn[NX], rn;
            const float l = phys_step<NU>(P, t, x, u, up, r6, sig, xi, xr, xn, rn);
            if (tape_w) {
                stt(t, L::S_ST, make_float4(r6[0], r6[1], r6[2], sig[0]));
                stt(t, L::S_ST + 1, make_float4(sig[1], sig[2], sig[3], sig[4]));
                stt(t, L::S_ST + 2, make_float4(sig[5], dsg[0], dsg[1], dsg[2]));
                stt(t, L::S_ST + 3, make_float4(dsg[3], dsg[4], dsg[5], rn));
                stt(t, L::S_ST + 4, make_float4(disc, x[0], x[1], x[2]));          // x is still x_t here
                stt(t, L::S_ST + 5, make_float4(x[3], x[4], x[5], x[6]));
                stt(t, L::S_ST + 6, make_float4(x[7], x[8], x[9], x[10]));
                stt(t, L::S_ST + 7, make_float4(x[11], x[12], 0.f, 0.f));
            }
            Jp = fma_(disc, l, Jp);
            disc = disc * P.discount;
            if (RATE && SDEMPC_RATE_ON(P)) {   // soft input-rate constraint (rate_cost in mpc_kernels.cuh): w_t e^2, e the violation of [lo, hi] by u_t - u_{t-1}
                const float wt = P.rate_w[t];
#pragma unroll
                for (int i = 0; i < NU; ++i) {
                    const float ds = u[i] - up[i];
                    const float e = ds > P.slew_hi[i] ? ds - P.slew_hi[i] : (ds < P.slew_lo[i] ? ds - P.slew_lo[i] : 0.f);
                    Jp = fma_(wt * e, e, Jp);
                }
            }
#pragma unroll
            for (int i = 0; i < NU; ++i) up[i] = u[i];
#pragma unroll
            for (int i = 0; i < NX; ++i) x[i] = xn[i];
            if (mode == 2) {   // predicted mean trajectory: particle mean, quaternion renormalised, external frame
                float rowv[NX], o[NX];
#pragma unroll
                for (int i = 0; i < NX; ++i) rowv[i] = pmean(xn[i]);
                quat_renorm(rowv + 6);
                if (enu) enu_ned(rowv, o);
                else {
#pragma unroll
                    for (int i = 0; i < NX; ++i) o[i] = rowv[i];
                }
                if (valid && pidx == 0) {
#pragma unroll
                    for (int i = 0; i < NX; ++i) P.x_evol[((size_t)b * (P.H + 1) + t + 1) * NX + i] = o[i];
                }
            }
        }
        const float Jm = pmean(Jp);
        if (state == S_FINAL) break;
        if (!SPECG && state == S_LS) {
            const int kslot = tid / PP;
            if (valid && pidx == 0) { sh.t_J[kslot] = Jm; sh.t_dec[kslot] = dec; }
            __syncthreads();
            if (tid < nq && sh.need[tid]) {   // take the first passing trial in order, exactly as the sequential search
                const int oq = tid, f = sh.first[oq], c = sh.cnt[oq];
                float s = sh.s[oq];
                const float fy = sh.fy[oq];
                int jn = sh.jn[oq];
                bool ok = false;
                float Jsel = 0.f;
                for (int i = 0; i < c; ++i) {
                    Jsel = sh.t_J[f + i];
                    ok = (Jsel <= fma_(P.coef, sh.t_dec[f + i], fy));
                    sh.n_ls[oq] = jn + 1;
                    if (ok) break;
                    if (jn < P.maxls) s = s * P.dec_f;
                    ++jn;
                }
                sh.s[oq] = s; sh.jn[oq] = jn; sh.Jp[oq] = Jsel; sh.ok[oq] = ok ? 1 : 0;
                sh.need[oq] = (!ok && jn <= P.maxls) ? 1 : 0;
            }
            __syncthreads();
            continue;
        }
        if constexpr (SPECG) {
            // ---- results of the forward tasks; problems whose line search resolved: accept / reject, plans, stop tests ----
            const int kslot = tid / PP;
            if (valid && pidx == 0) { sh.t_J[kslot] = Jm; sh.t_dec[kslot] = dec; }
            if (tape_on) {
#pragma unroll
                for (int i = 0; i < NX; ++i) XH[i * 128] = x[i];
            }
            sh.b_q[tid] = -1;
            __syncthreads();
            if (tid < nq) {
                const int oq = tid, ph = sh.ph[oq], f = sh.first[oq], fT = sh.firstT[oq];
                unsigned char a = 2;
                if (ph == 0) {                      // gradient forward pass done: its tape is in slot fT
                    sh.fy[oq] = sh.t_J[fT]; sh.trow[oq] = (short)fT; sh.ph[oq] = 2;
                } else if (ph == 1) {               // take the first passing trial in order, exactly as the sequential search
                    const int c = sh.cnt[oq];
                    float s = sh.s[oq];
                    const float fy = sh.fy[oq];
                    int jn = sh.jn[oq], isel = 0;
                    bool ok = false;
                    float Jsel = 0.f;
                    for (int i = 0; i < c; ++i) {
                        Jsel = sh.t_J[f + i];
                        ok = (Jsel <= fma_(P.coef, sh.t_dec[f + i], fy));
                        sh.n_ls[oq] = jn + 1;
                        isel = i;
                        if (ok) break;
                        if (jn < P.maxls) s = s * P.dec_f;
                        ++jn;
                    }
                    sh.s[oq] = s; sh.jn[oq] = jn; sh.Jp[oq] = Jsel; sh.ok[oq] = ok ? 1 : 0;
                    if (ok || jn > P.maxls) {       // resolved
                        a = (ok && (Jsel <= sh.Jx[oq])) ? 1 : 0;
                        const unsigned sm = sh.smask[oq];
                        int row = -1;                // the tape task that ran what this outcome needs next, if any
                        if (a == 0) { if (sh.scnt[oq] > 0) row = fT; }
                        else if ((sm >> isel) & 1u) row = fT + 1 + __popc(sm & ((1u << isel) - 1u));
                        sh.trow[oq] = (short)row;
                        if (row >= 0) sh.fy_next[oq] = sh.t_J[row];
                    }
                }
                sh.acc[oq] = a;
            } else sh.acc[tid] = 2;
            __syncthreads();
            plan_update();
            __syncthreads();   // sh.k is read above, updated below
            if (tid < nq && sh.acc[tid] != 2) {
                const int oq = tid;
                bookkeeping(oq);
                if (!sh.active[oq]) sh.ph[oq] = 3;
                else if (sh.trow[oq] >= 0) { sh.ph[oq] = 2; sh.fy[oq] = sh.fy_next[oq]; }   // speculation hit: sweep only
                else sh.ph[oq] = 0;
            }
            __syncthreads();
            // ---- adjoint sweep over the problems whose tape is ready, each in the slot that recorded it ----
            const bool rdy = tid < nq && sh.ph[tid] == 2;
            if (rdy) sh.b_q[sh.trow[tid]] = (short)tid;
            if (__syncthreads_or(rdy ? 1 : 0) == 0) continue;
            const int bq = (int)sh.b_q[kslot];
            valid = bq >= 0;
            q = valid ? bq : 0;
            xi_row = q * PP + pidx;
#pragma unroll
            for (int i = 0; i < NX; ++i) x[i] = valid ? XH[i * 128] : 0.f;
        }
        // ================= GRAD: adjoint sweep on the tensor cores, gradient -> G (particle mean) =================
        {
            float lam[NX], gp[NU], xn[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) { lam[i] = 0.f; xn[i] = x[i]; }   // x holds x_H after the forward loop
#pragma unroll
            for (int i = 0; i < NU; ++i) gp[i] = 0.f;
            float gsq = 0.f;
            auto ldt = [&](int t, int g) -> float4 { return !valid ? make_float4(0.f, 0.f, 0.f, 0.f) : lone_tape ? *tp(t, g) : __ldcs(tp(t, g)); };
            for (int t = P.H - 1; t >= 0; --t) {
                if (t > 0 && valid) {   // the sweep consumes its tape as it loads it: pull the previous step's granules towards the SM now
#pragma unroll
                    for (int g = 0; g < L::STG; ++g) { if (lone_cta) tc::prefetch_l1(tp(t - 1, g)); else tc::prefetch_l2(tp(t - 1, g)); }
                }
                float xt[NX], xi[6], xr[NX];
                // (loading these a step ahead into registers, as the forward pass does with its operands, LOSES here: 34.6 -> 36.8 ms
                // at 4096 problems — 32 more live registers across three contraction waits)
                const float4 s0 = ldt(t, L::S_ST), s1 = ldt(t, L::S_ST + 1), s2 = ldt(t, L::S_ST + 2), s3 = ldt(t, L::S_ST + 3),
                             s4 = ldt(t, L::S_ST + 4);
                {
                    const float4 c5 = ldt(t, L::S_ST + 5), c6 = ldt(t, L::S_ST + 6), c7 = ldt(t, L::S_ST + 7);
                    xt[0] = s4.y; xt[1] = s4.z; xt[2] = s4.w; xt[3] = c5.x; xt[4] = c5.y; xt[5] = c5.z; xt[6] = c5.w;
                    xt[7] = c6.x; xt[8] = c6.y; xt[9] = c6.z; xt[10] = c6.w; xt[11] = c7.x; xt[12] = c7.y;
                }
                {
                    const float* xit = XI + xi_row + t * (6 * 128);
                    const float* xrt = XREF + q + (t + 1) * (NX * RS);
                    const float* yt = YK + q + t * (NU * RS);
                    const float* ypt = (t == 0) ? UPREV + q : yt - NU * RS;
#pragma unroll
                    for (int i = 0; i < 6; ++i) xi[i] = xit[i * 128];
#pragma unroll
                    for (int i = 0; i < NX; ++i) xr[i] = xrt[i * RS];
#pragma unroll
                    for (int i = 0; i < NU; ++i) u[i] = yt[i * RS];
#pragma unroll
                    for (int i = 0; i < NU; ++i) up[i] = ypt[i * RS];
                }
                BwdMid mid;
                float gu[NU];
                float2 lo[6];
                {
                    const float r012[3] = {s0.x, s0.y, s0.z};
                    const float sg6[6] = {s0.w, s1.x, s1.y, s1.z, s1.w, s2.x};
                    const float dsg[6] = {s2.y, s2.z, s2.w, s3.x, s3.y, s3.z};
                    bwd_pre<NU>(P, t, xt, xn, xr, u, r012, sg6, dsg, s3.w, s4.x, xi, lam, mid, lo, gu);
                }
                {
                    const float a[16] = {lo[0].x, lo[1].x, lo[2].x, lo[3].x, lo[4].x, lo[5].x, lo[0].y, lo[1].y,
                                         lo[2].y, lo[3].y, lo[4].y, lo[5].y, 0.f, 0.f, 0.f, 0.f};
                    tc::st16(lane_addr + L::C_A, a);
                }
                tc::publish();
                if (tid == 0) {
#pragma unroll
                    for (int k8 = 0; k8 < 2; ++k8)
                        tc::mma_ts(tb + L::C_D12, tb + L::C_A + 8 * k8, tc::desc(sb + L::B3T + k8 * 2 * L::LBO, L::SBO16), id12, k8 > 0);
                    tc::commit(bar);
                }
                float4 g0 = ldt(t, L::S_H2), g1 = ldt(t, L::S_H2 + 1);
                tc::wait(bar, phase); phase ^= 1;
#pragma unroll
                for (int c0 = 0; c0 < N12; c0 += 16) {
                    float4 n0 = g0, n1 = g1;
                    if (c0 + 16 < N12) { n0 = ldt(t, L::S_H2 + (c0 + 16) / 8); n1 = ldt(t, L::S_H2 + (c0 + 16) / 8 + 1); }
                    float v[16];
                    tc::ld16(lane_addr + L::C_D12 + c0, v);
                    {
                        float h[8];
                        tc::unpack8(g0, h);
#pragma unroll
                        for (int m = 0; m < 8; ++m) v[m] = v[m] * fma_(-h[m], h[m], 1.f);
                        tc::unpack8(g1, h);
#pragma unroll
                        for (int m = 0; m < 8; ++m) v[8 + m] = v[8 + m] * fma_(-h[m], h[m], 1.f);
                    }
                    tc::st16(lane_addr + L::C_A + c0, v);
                    g0 = n0; g1 = n1;
                }
                tc::publish();
                if (tid == 0) {
#pragma unroll
                    for (int nn = 0; nn < 2; ++nn)
#pragma unroll
                        for (int k8 = 0; k8 < W / 8; ++k8)
                            tc::mma_ts(tb + L::C_D12 + nn * W, tb + L::C_A + nn * W + 8 * k8,
                                       tc::desc(sb + L::B2T + nn * L::NET2 + k8 * 2 * L::LBO, L::SBOW), idW, k8 > 0);
                    tc::commit(bar);
                }
                g0 = ldt(t, L::S_H1); g1 = ldt(t, L::S_H1 + 1);
                tc::wait(bar, phase); phase ^= 1;
#pragma unroll
                for (int c0 = 0; c0 < N12; c0 += 16) {
                    float4 n0 = g0, n1 = g1;
                    if (c0 + 16 < N12) { n0 = ldt(t, L::S_H1 + (c0 + 16) / 8); n1 = ldt(t, L::S_H1 + (c0 + 16) / 8 + 1); }
                    float v[16];
                    tc::ld16(lane_addr + L::C_D12 + c0, v);
                    {
                        float h[8];
                        tc::unpack8(g0, h);
#pragma unroll
                        for (int m = 0; m < 8; ++m) v[m] = v[m] * fma_(-h[m], h[m], 1.f);
                        tc::unpack8(g1, h);
#pragma unroll
                        for (int m = 0; m < 8; ++m) v[8 + m] = v[8 + m] * fma_(-h[m], h[m], 1.f);
                    }
                    tc::st16(lane_addr + L::C_A + c0, v);
                    g0 = n0; g1 = n1;
                }
                tc::publish();
                if (tid == 0) {
#pragma unroll
                    for (int k8 = 0; k8 < L::K2 / 8; ++k8)
                        tc::mma_ts(tb + L::C_D3, tb + L::C_A + 8 * k8, tc::desc(sb + L::B1T + k8 * 2 * L::LBO, L::SBO2), id16, k8 > 0);
                    tc::commit(bar);
                }
                tc::wait(bar, phase); phase ^= 1;
                float lz[NIN];
                {
                    float o[16];
                    tc::ld16(lane_addr + L::C_D3, o);
#pragma unroll
                    for (int i = 0; i < NIN; ++i) lz[i] = o[i];
                }
                bwd_post<NU>(P, xt, u, up, mid, lz, gu, gp, lam);
                if (RATE && SDEMPC_RATE_ON(P)) {   // its gradient (rate_grad_add): 2 w_t e_t - 2 w_{t+1} e_{t+1}
                    const float* yt = YK + q + t * (NU * RS);
                    const float wt = 2.f * P.rate_w[t], wn = (t + 1 < P.H) ? 2.f * P.rate_w[t + 1] : 0.f;
#pragma unroll
                    for (int i = 0; i < NU; ++i) {
                        const float hi = P.slew_hi[i], lo = P.slew_lo[i];
                        const float ds = u[i] - up[i];
                        float a = wt * (ds > hi ? ds - hi : (ds < lo ? ds - lo : 0.f));
                        if (t + 1 < P.H) {
                            const float dn = yt[(NU + i) * RS] - u[i];
                            a = a - wn * (dn > hi ? dn - hi : (dn < lo ? dn - lo : 0.f));
                        }
                        gu[i] = gu[i] + a;
                    }
                }
#pragma unroll
                for (int i = 0; i < NU; ++i) {
                    const float gm = pmean(gu[i]);
                    gsq = fma_(gm, gm, gsq);
                    if (valid && pidx == 0) G[q + t * (NU * RS) + i * RS] = gm;
                }
#pragma unroll
                for (int i = 0; i < NX; ++i) xn[i] = xt[i];
            }
            if (valid && pidx == 0) {
                if constexpr (!SPECG) sh.fy[q] = Jm;       // (the speculative build set it when the forward task resolved)
                sh.gsq[q] = gsq;
            }
        }
        __syncthreads();
        if (tid < nq && sh.active[tid] && (!SPECG || sh.ph[tid] == 2)) {   // thread q = problem q: open this iteration's line search
            const int oq = tid;
            if constexpr (SPECG) sh.ph[oq] = 1;
            const int it = sh.it[oq] + 1;
            sh.it[oq] = it;
            if (it == 1) { sh.Jx[oq] = sh.fy[oq]; sh.init_cost[oq] = sh.fy[oq]; }
            if (P.reset_option == 1) { float s = sh.s[oq] * P.inc_f; sh.s[oq] = s > P.max_step ? P.max_step : s; }
            sh.need[oq] = 1; sh.ok[oq] = 0; sh.jn[oq] = 0; sh.n_ls[oq] = 0;
        }
        __syncthreads();
        state = S_LS;
    }
    // ---- outputs: plan, telemetry ----
    __syncthreads();
    if (tid < nq) {
        const int oq = tid, b = b0 + oq;
        for (int i = 0; i < n; ++i) P.u_plan_out[(size_t)b * n + i] = XK[i * RS + oq];
        sdempc_info inf;
        const float itf = (float)(sh.it[oq] > 0 ? sh.it[oq] : 1);
        inf.avg_linesearch = __fdiv_rn(sh.sum_ls[oq], itf);
        inf.stepsize = sh.s[oq];
        inf.num_steps = (float)sh.it[oq];
        inf.grad_sqr = sh.gsq[oq];
        inf.avg_stepsize = __fdiv_rn(sh.sum_s[oq], itf);
        inf.init_cost = sh.init_cost[oq];
        const float Jx = sh.Jx[oq];
        inf.opt_cost = (Jx == Jx) ? Jx : __int_as_float(0x7f800000);
        inf.solve_time_us = 0.f;
        P.info_out[b] = inf;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"((uint32_t)L::COLS));
}

}  // namespace sdempc
