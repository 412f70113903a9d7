// mpc_entry.cuh — the __global__ entry points of the FP32 (SPEC-ARITH) kernels and the table entry that names every kernel
// compiled for one (nu, width, particles) combination.  Included by the instantiation units kern_*.cu (one or two
// combinations each, compiled in parallel) and, for the KernelChoice type only, by sdempc_api.cu.
// Compile the instantiation units with -fmad=false: every fused multiply-add of SPEC-ARITH is written explicitly.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "mpc_group.cuh"
#include "mpc_kernels.cuh"
#include "mpc_pcluster.cuh"

namespace sdempc {

constexpr int GROUP_GW = 8;   // warps per CTA of the throughput kernel
// problems per warp of the throughput kernel: what fits 227 KB of shared memory next to the staged weights
constexpr int group_gp(int nu, int w) { return (w == 32 && nu <= 4) ? 4 : (w == 32) ? 3 : 2; }
constexpr int SPEC_LSW = 4;   // line-search warps of the latency kernel: one per SM sub-partition
constexpr int SPEC_SGW = 4;   // speculative-gradient warps (candidates: trial 0/1/2 accepted, step rejected)
// Cluster latency kernel (mpc_pcluster.cuh): line-search and speculative-gradient REPLICAS of P warps each, at most 32 warps
// per problem.  Two shapes: "compact" = 4 warps per CTA (one per SM sub-partition; P = 8: 8 CTAs, the portable cluster
// size) and "wide" = the same warps at 2 per CTA on twice the SMs — the warps of an SM share its shared-memory pipe, which
// four width-64 warps saturate (weights read from shared memory every step) and four width-32 warps load enough to cost 6 %.
constexpr int pc_lsw(int P) { return P >= 8 ? 2 : 4; }
constexpr int pc_sgw(int P) { return P >= 8 ? 2 : 4; }
constexpr int PC_WPC = 4;
constexpr int pcw_wpc(int) { return 2; }   // warps per CTA of the wide shape (P = 1 on 8 CTAs of ONE warp: 6.40 ms against 6.03)
constexpr int pc_cluster_ctas(int P, int wpc) { return P * (pc_lsw(P) + pc_sgw(P)) / wpc; }

struct KernelChoice {
    void (*solve)(KParams);
    void (*rollout)(KParams);
    void (*closed)(KParams);
    void (*solve_spec)(KParams);   // latency mode: one problem per CTA, SPEC_LSW warps (P == 1 only)
    void (*closed_spec)(KParams);
    void (*solve_cl)(KParams);     // latency mode on a 2-CTA cluster (line search on one SM, speculation on its neighbour)
    void (*closed_cl)(KParams);
    void (*solve_pc)(KParams);     // cluster latency kernel, compact shape: pc_cluster_ctas(P, PC_WPC) CTAs per problem
    void (*solve_pcw)(KParams);    // ... wide shape: pc_cluster_ctas(P, pcw_wpc(P)) CTAs per problem (16 for P >= 4: non-portable size)
    void (*solve_group)(KParams);  // throughput mode: GROUP_GW warps x gp problems per CTA (P == 1)
    void (*rollout_tc)(KParams);   // tensor-core forward rollout (any power-of-two particle count), SDEMPC_F_TENSOR
    void (*rollout_tc_grad)(KParams);   // ... with the adjoint sweep
    void (*solve_tc)(KParams);     // tensor-core batched APG solve, SDEMPC_F_TENSOR
    void (*solve_tc_spec)(KParams);// ... with the speculative gradient pass (few problems per CTA)
    void (*solve_tc_rate)(KParams);// ... with the soft input-rate constraint
    void (*solve_tc_lat)(KParams); // ... built for two CTAs per SM (more registers): small and medium batches
    int tc_bytes, tc_bytes_grad, tc_tape_granules, tc_bytes_solve, tc_solve_tape_granules, tc_cols;
    int gp;
    int nu, W, P, G;
    int wimg_floats, wsmem_floats;
    bool wreg;
};

enum { MODE_SOLVE = 0, MODE_ROLLOUT = 1, MODE_CLOSED_LOOP = 2 };

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}

// Stage the weight image once per CTA: TMA bulk copy completed on an mbarrier (every thread then waits on it).
template <uint32_t BYTES>
__device__ __forceinline__ void stage_weights(float* ws, const float* wimg, uint64_t* bar) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, BYTES);
        constexpr uint32_t CH = 32768;   // keep each bulk request modest
        for (uint32_t off = 0; off < BYTES; off += CH)
            tma_bulk_g2s(reinterpret_cast<char*>(ws) + off, reinterpret_cast<const char*>(wimg) + off,
                         (BYTES - off) < CH ? (BYTES - off) : CH, bar);
    }
}

// Team-level epilogue shared by solve and rollout: mean trajectory -> external frame -> global
template <int NU, int W, int PP>
__device__ __forceinline__ void write_x_evol(const KParams& P, const Team<PP>& tm, Warp<NU, W>& c, float* dst) {
    tm.sync();   // all particles' state tapes complete
    if (tm.warp_in_team == 0 && tm.ls_index == 0 && dst != nullptr) {
        const int t = c.lane;
        if (t <= P.H) {
            const float invP = __fdiv_rn(1.0f, (float)PP);
            float row[NX], o[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) row[i] = tm.warp0_base[P.o_xtape + t * 16 + i];
#pragma unroll
            for (int p = 1; p < PP; ++p)
#pragma unroll
                for (int i = 0; i < NX; ++i) row[i] = row[i] + tm.warp0_base[p * tm.ws_stride + P.o_xtape + t * 16 + i];
#pragma unroll
            for (int i = 0; i < NX; ++i) row[i] = row[i] * invP;
            quat_renorm(row + 6);
            if (P.flags & SDEMPC_F_FRAME_ENU) enu_ned(row, o);
            else {
#pragma unroll
                for (int i = 0; i < NX; ++i) o[i] = row[i];
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) dst[t * NX + i] = o[i];
        }
    }
    tm.sync();   // tapes may be overwritten by the next problem
}

// CL = 1 (latency kernel on a 2-CTA cluster): CTA rank 0 holds the LSW line-search warps, rank 1 the SGW
// speculation warps of ONE problem; launched with cluster dimension 2.
template <int NU, int W, int PP, int G, int MODE, int LSW = 1, int SGW = 0, int CL = 0>
__global__ void __launch_bounds__(CL ? LSW * 32 : G * PP * (LSW + SGW) * 32, 1) mpc_kernel(const __grid_constant__ KParams P) {
    static_assert(!CL || (G == 1 && PP == 1 && LSW == SGW), "cluster mode: one problem, equal halves");
    using L = Layout<NU, W>;
    extern __shared__ __align__(128) float smem[];
    float* ws = smem;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::SMEM_FLOATS);
    float* team_base = smem + L::SMEM_FLOATS + 4;
    float* warp_base = team_base + G * P.team_stride;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WPT = CL ? LSW : PP * (LSW + SGW);   // warps per team (per CTA in cluster mode)
    unsigned crank = 0;
    if constexpr (CL) crank = cooperative_groups::this_cluster().block_rank();
    const int team = warp / WPT, wit = (warp % WPT) % PP, ls = (warp % WPT) / PP + (CL ? (int)crank * LSW : 0);

    stage_weights<L::SMEM_FLOATS * 4>(ws, P.wimg, bar);

    Warp<NU, W> c;
    c.lane = lane;
    c.ws = ws;
    float* wb = warp_base + (size_t)warp * P.ws_stride;
    c.xk = wb + P.o_xk; c.yk = wb + P.o_yk; c.g = wb + P.o_g; c.g2 = wb + P.o_g2; c.xp = wb + P.o_xp; c.uprev = wb + P.o_uprev;
    c.xref = wb + P.o_xref; c.xi = wb + P.o_xi; c.xtape = wb + P.o_xtape; c.stape = wb + P.o_stape;
    c.bufA = wb + P.o_bufA; c.bufB = wb + P.o_bufB; c.act3 = wb + P.o_act3; c.lz = wb + P.o_lz; c.red = wb + P.o_red;
    if (P.mtape_g != nullptr)
        c.mtape = P.mtape_g + ((size_t)blockIdx.x * (G * WPT) + warp) * (size_t)P.H * 2 * W;
    else
        c.mtape = reinterpret_cast<float2*>(wb + P.o_mtape);
    c.load_regs(P.wimg);

    Team<PP> tm;
    tm.warp_in_team = wit;
    tm.bar_id = 1 + team;
    tm.scratch = team_base + team * P.team_stride;
    tm.warp0_base = warp_base + (size_t)(team * WPT + ls * PP) * P.ws_stride;
    tm.team0_base = warp_base + (size_t)(team * WPT) * P.ws_stride;
    tm.spec0_base = tm.team0_base + (size_t)LSW * P.ws_stride;
    if constexpr (CL) {   // slots live in CTA 0, speculation gradients in CTA 1: distributed shared memory
        auto cluster = cooperative_groups::this_cluster();
        tm.scratch = cluster.map_shared_rank(team_base, 0);
        tm.spec0_base = cluster.map_shared_rank(warp_base, 1);
    }
    tm.ws_stride = P.ws_stride;
    tm.ls_index = ls;
    tm.ls_bar_id = 1 + team;

    mbar_wait(bar, 0);
    if constexpr (CL) cooperative_groups::this_cluster().sync();   // both CTAs resident before any DSMEM access

    const int n = P.H * NU;
    const bool enu = (P.flags & SDEMPC_F_FRAME_ENU) != 0;

    // one warp (team) per problem: problems are dealt across CTAs first, so a small batch spreads over all SMs
    const int b_first = CL ? (int)(blockIdx.x / 2) : (int)(team * gridDim.x + blockIdx.x);
    const int b_step = CL ? (int)(gridDim.x / 2) : (int)(gridDim.x * G);
    for (int b = b_first; b < P.B; b += b_step) {
        // ---- state ----
        float x0[NX];
        {
            const float* xs = P.x + (size_t)b * NX;
            float tmp[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) tmp[i] = __ldg(xs + i);
            if (enu) enu_ned(tmp, x0);
            else {
#pragma unroll
                for (int i = 0; i < NX; ++i) x0[i] = tmp[i];
            }
        }
        if constexpr (MODE != MODE_CLOSED_LOOP) {
            build_window(P, lane, c.xref, P.xref_win ? P.xref_win + (size_t)b * (P.H + 1) * NX : nullptr,
                                P.curr_t ? P.curr_t + b : nullptr, P.xdes ? P.xdes + (size_t)b * NX : nullptr, 0.f, false);
            if (P.xi_override != nullptr) {
                if (lane < P.H) {
                    const float* src = P.xi_override + (((size_t)b * PP + wit) * P.H + lane) * 6;
#pragma unroll
                    for (int i = 0; i < 6; ++i) c.xi[lane * 8 + i] = __ldg(src + i);
                }
            } else {
                gen_noise(lane, P.rng[2 * (size_t)b], P.rng[2 * (size_t)b + 1], (uint32_t)wit, 0u, P.H, c.xi);
            }
        }

        if constexpr (MODE == MODE_SOLVE) {
            const float* pin = P.u_plan + (size_t)b * n;   // staged input plan (never written by the kernel)
            if (lane < NU) c.uprev[lane] = __ldg(pin + lane);
            for (int i = lane; i < n; i += 32) {
                const int t = i / NU, ii = i % NU;
                const int ts = (P.flags & SDEMPC_F_NO_SHIFT) ? t : (t + 1 < P.H ? t + 1 : P.H - 1);
                c.xk[i] = clipf(__ldg(pin + ts * NU + ii), P.u_lo[ii], P.u_hi[ii]);
            }
            __syncwarp();
            float s = P.info[b].stepsize;
            s = s > 0.f ? s : P.init_step;
            sdempc_info inf;
            float* trp = P.trace ? P.trace + (size_t)b * P.max_iter * SDEMPC_TRACE_W : nullptr;
            if constexpr (LSW > 1) apg_solve_latency<NU, W, LSW, SGW, CL>(P, c, tm, x0, s, inf, trp);
            else apg_solve<NU, W, PP, 1>(P, c, tm, x0, s, inf, trp);
            float* pout = P.u_plan_out + (size_t)b * n;
            if (wit == 0 && ls == 0) {
                for (int i = lane; i < n; i += 32) pout[i] = c.xk[i];
                if (lane == 0) P.info_out[b] = inf;
            }
            write_x_evol<NU, W, PP>(P, tm, c, P.x_evol + (size_t)b * (P.H + 1) * NX);
        } else if constexpr (MODE == MODE_ROLLOUT) {
            const float* uin = P.u_in + (size_t)b * n;
            if (lane < NU) c.uprev[lane] = __ldg(P.uprev_in + (size_t)b * NU + lane);
            for (int i = lane; i < n; i += 32) c.xk[i] = __ldg(uin + i);
            __syncwarp();
            const float invP = __fdiv_rn(1.0f, (float)PP);
            const float Jw = rollout_fwd<NU, W, 1>(P, c, c.xk, x0);
            if (P.grad_out != nullptr) rollout_bwd<NU, W>(P, c, c.xk);
            const float J = team_mean_cost<PP>(tm, Jw, lane, invP);
            if (P.grad_out != nullptr) {
                team_mean_grad<NU, W, PP>(P, tm, c, n, invP);
                if (wit == 0)
                    for (int i = lane; i < n; i += 32) P.grad_out[(size_t)b * n + i] = c.g[i];
            }
            if (wit == 0 && lane == 0) P.cost_out[b] = J;
            write_x_evol<NU, W, PP>(P, tm, c, P.x_evol ? P.x_evol + (size_t)b * (P.H + 1) * NX : nullptr);
        } else {
            // ---- Monte-Carlo closed loop: `ticks` x (solve -> plant step) on device ----
            const unsigned long long seed = P.rng[2 * (size_t)b];
            unsigned long long tick = P.rng[2 * (size_t)b + 1];
            const float t0 = __ldg(P.t0 + b);
            for (int i = lane; i < n; i += 32) { const int ii = i % NU; c.xk[i] = clipf(P.uref[ii], P.u_lo[ii], P.u_hi[ii]); }
            __syncwarp();
            float s = P.init_step;
            float se = 0.f, me = 0.f, sc = 0.f, sn = 0.f;
            const float dt0 = P.dt[0];
            for (int k = 0; k < P.ticks; ++k, ++tick) {
                if (P.x_hist != nullptr && wit == 0 && ls == 0 && lane == 0) {
                    float o[NX];
                    if (enu) enu_ned(x0, o);
                    else {
#pragma unroll
                        for (int i = 0; i < NX; ++i) o[i] = x0[i];
                    }
#pragma unroll
                    for (int i = 0; i < NX; ++i) P.x_hist[((size_t)b * (P.ticks + 1) + k) * NX + i] = o[i];
                }
                build_window(P, lane, c.xref, nullptr, nullptr, nullptr, fma_((float)k, dt0, t0), true);
                gen_noise(lane, seed, tick, (uint32_t)wit, 0u, P.H, c.xi);
                // warm start: uprev = plan[0], shift
                float keep[(SDEMPC_MAX_H * SDEMPC_MAX_NU + 31) / 32];
                {
                    int m = 0;
                    for (int i = lane; i < n; i += 32, ++m) {
                        const int t = i / NU, ii = i % NU;
                        const int ts = (P.flags & SDEMPC_F_NO_SHIFT) ? t : (t + 1 < P.H ? t + 1 : P.H - 1);
                        keep[m] = clipf(c.xk[ts * NU + ii], P.u_lo[ii], P.u_hi[ii]);
                    }
                    if (lane < NU) c.uprev[lane] = c.xk[lane];
                    __syncwarp();
                    m = 0;
                    for (int i = lane; i < n; i += 32, ++m) c.xk[i] = keep[m];
                    __syncwarp();
                }
                sdempc_info inf;
                if constexpr (LSW > 1) apg_solve_latency<NU, W, LSW, SGW, CL>(P, c, tm, x0, s, inf, nullptr);
                else apg_solve<NU, W, PP, 1>(P, c, tm, x0, s, inf, nullptr);
                s = inf.stepsize;
                sc = sc + inf.opt_cost;
                sn = sn + inf.num_steps;
                float u0[NU];
                load_u<NU>(c.xk, 0, u0);
                if (P.u_hist != nullptr && wit == 0 && ls == 0 && lane == 0) {
#pragma unroll
                    for (int i = 0; i < NU; ++i) P.u_hist[((size_t)b * P.ticks + k) * NU + i] = u0[i];
                }
                // plant step: same SDE, one particle, Philox sub-stream 2 (every warp of the team
                // integrates the same plant state redundantly)
                tm.sync();
                gen_noise(lane, seed, tick, 0u, 2u, 1, c.xi);
                __syncwarp();
                (void)fwd_step<NU, W, 0>(P, c, 0, 1.f, x0, u0, u0);
                __syncwarp();
                float e2 = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) { const float d = x0[i] - c.xref[16 + i]; e2 = fma_(d, d, e2); }
                se = se + e2;
                me = e2 > me ? e2 : me;
                __syncwarp();   // the window is rebuilt by the next tick
            }
            if (wit == 0 && ls == 0 && lane == 0) {
                if (P.x_hist != nullptr) {
                    float o[NX];
                    if (enu) enu_ned(x0, o);
                    else {
#pragma unroll
                        for (int i = 0; i < NX; ++i) o[i] = x0[i];
                    }
#pragma unroll
                    for (int i = 0; i < NX; ++i) P.x_hist[((size_t)b * (P.ticks + 1) + P.ticks) * NX + i] = o[i];
                }
                const float tk = (float)P.ticks;
                P.stats[(size_t)b * 4 + 0] = __fsqrt_rn(__fdiv_rn(se, tk));
                P.stats[(size_t)b * 4 + 1] = __fsqrt_rn(me);
                P.stats[(size_t)b * 4 + 2] = __fdiv_rn(sc, tk);
                P.stats[(size_t)b * 4 + 3] = __fdiv_rn(sn, tk);
            }
            tm.sync();
        }
    }
    if constexpr (CL) cooperative_groups::this_cluster().sync();   // keep shared memory alive for the sibling CTA
}

// Latency kernel for P > 1 (mpc_pcluster.cuh): one problem per cluster of PP*(LSW+SGW)/WPC CTAs, WPC warps per CTA.
template <int NU, int W, int PP, int LSW, int SGW, int WPC = 4>
__global__ void __launch_bounds__(WPC * 32, 1) mpc_pcluster_kernel(const __grid_constant__ KParams P) {
    using L = Layout<NU, W>;
    using PC = PCluster<PP, LSW, SGW, WPC>;
    extern __shared__ __align__(128) float smem[];
    float* ws = smem;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::SMEM_FLOATS);
    float* team_base = smem + L::SMEM_FLOATS + 4;
    float* warp_base = team_base + P.team_stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto cluster = cooperative_groups::this_cluster();
    const int crank = (int)cluster.block_rank();

    stage_weights<L::SMEM_FLOATS * 4>(ws, P.wimg, bar);

    Warp<NU, W> c;
    c.lane = lane;
    c.ws = ws;
    float* wb = warp_base + (size_t)warp * P.ws_stride;
    c.xk = wb + P.o_xk; c.yk = wb + P.o_yk; c.g = wb + P.o_g; c.g2 = wb + P.o_g2; c.xp = wb + P.o_xp; c.uprev = wb + P.o_uprev;
    c.xref = wb + P.o_xref; c.xi = wb + P.o_xi; c.xtape = wb + P.o_xtape; c.stape = wb + P.o_stape;
    c.bufA = wb + P.o_bufA; c.bufB = wb + P.o_bufB; c.act3 = wb + P.o_act3; c.lz = wb + P.o_lz; c.red = wb + P.o_red;
    if (P.mtape_g != nullptr) c.mtape = P.mtape_g + ((size_t)blockIdx.x * WPC + warp) * (size_t)P.H * 2 * W;
    else c.mtape = reinterpret_cast<float2*>(wb + P.o_mtape);
    c.load_regs(P.wimg);

    PC pc;
    pc.gwi = crank * WPC + warp;
    pc.l = pc.gwi / PP;
    pc.p = pc.gwi % PP;
    pc.xc_local = team_base;
    pc.warp_base_local = warp_base;
    pc.ws_stride = P.ws_stride;

    mbar_wait(bar, 0);
    cluster.sync();   // every CTA of the cluster resident before any DSMEM access

    const int n = P.H * NU;
    const bool enu = (P.flags & SDEMPC_F_FRAME_ENU) != 0;
    for (int b = (int)(blockIdx.x / PC::CS); b < P.B; b += (int)(gridDim.x / PC::CS)) {
        float x0[NX];
        {
            float tmp[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) tmp[i] = __ldg(P.x + (size_t)b * NX + i);
            if (enu) enu_ned(tmp, x0);
            else {
#pragma unroll
                for (int i = 0; i < NX; ++i) x0[i] = tmp[i];
            }
        }
        build_window(P, lane, c.xref, P.xref_win ? P.xref_win + (size_t)b * (P.H + 1) * NX : nullptr,
                     P.curr_t ? P.curr_t + b : nullptr, P.xdes ? P.xdes + (size_t)b * NX : nullptr, 0.f, false);
        if (P.xi_override != nullptr) {
            if (lane < P.H) {
                const float* src = P.xi_override + (((size_t)b * PP + pc.p) * P.H + lane) * 6;
#pragma unroll
                for (int i = 0; i < 6; ++i) c.xi[lane * 8 + i] = __ldg(src + i);
            }
        } else {
            gen_noise(lane, P.rng[2 * (size_t)b], P.rng[2 * (size_t)b + 1], (uint32_t)pc.p, 0u, P.H, c.xi);
        }
        const float* pin = P.u_plan + (size_t)b * n;
        if (lane < NU) c.uprev[lane] = __ldg(pin + lane);
        for (int i = lane; i < n; i += 32) {
            const int t = i / NU, ii = i % NU;
            const int ts = (P.flags & SDEMPC_F_NO_SHIFT) ? t : (t + 1 < P.H ? t + 1 : P.H - 1);
            c.xk[i] = clipf(__ldg(pin + ts * NU + ii), P.u_lo[ii], P.u_hi[ii]);
        }
        __syncwarp();
        float s = P.info[b].stepsize;
        s = s > 0.f ? s : P.init_step;
        sdempc_info inf;
        apg_solve_pcluster<NU, W, PP, LSW, SGW, WPC>(P, c, pc, warp, x0, s, inf,
                                           P.trace ? P.trace + (size_t)b * P.max_iter * SDEMPC_TRACE_W : nullptr);
        if (pc.gwi == 0) {
            for (int i = lane; i < n; i += 32) P.u_plan_out[(size_t)b * n + i] = c.xk[i];
            if (lane == 0) P.info_out[b] = inf;
        }
        pc.barrier();   // every particle's state tape complete
        if (pc.gwi == 0 && lane <= P.H) {
            const float invP = __fdiv_rn(1.0f, (float)PP);
            float row[NX], o[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) row[i] = pc.region(0)[P.o_xtape + lane * 16 + i];
#pragma unroll
            for (int q = 1; q < PP; ++q) {
                const float* rq = pc.region(q) + P.o_xtape + lane * 16;
#pragma unroll
                for (int i = 0; i < NX; ++i) row[i] = row[i] + rq[i];
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) row[i] = row[i] * invP;
            quat_renorm(row + 6);
            if (enu) enu_ned(row, o);
            else {
#pragma unroll
                for (int i = 0; i < NX; ++i) o[i] = row[i];
            }
            float* dst = P.x_evol + (size_t)b * (P.H + 1) * NX + lane * NX;
#pragma unroll
            for (int i = 0; i < NX; ++i) dst[i] = o[i];
        }
        pc.barrier();   // tapes may be overwritten by the next problem
    }
    cluster.sync();
}

// Throughput kernel: GW warps per CTA, each warp owns GP problems (mpc_group.cuh).  P = 1.
template <int NU, int W, int GP, int GW>
__global__ void __launch_bounds__(GW * 32, 1) mpc_group_kernel(const __grid_constant__ KParams P) {
    using L = Layout<NU, W>;
    extern __shared__ __align__(128) float smem[];
    float* ws = smem;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::SMEM_FLOATS);
    float* xchg = smem + L::SMEM_FLOATS + 4;
    float* regions = xchg + GW * P.gx_stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    stage_weights<L::SMEM_FLOATS * 4>(ws, P.wimg, bar);

    Warp<NU, W> c;
    c.lane = lane;
    c.ws = ws;
    c.bufA = xchg + warp * P.gx_stride;
    c.bufB = c.bufA + 2 * W;
    c.act3 = c.bufB + 2 * W;
    c.xk = c.yk = c.g = c.xp = c.uprev = c.xref = c.xi = c.xtape = c.stape = c.lz = c.red = nullptr;
    c.mtape = nullptr;
    c.load_regs(P.wimg);

    Group<NU, W, GP> gw;
    gw.base = regions + (size_t)warp * GP * P.ws_stride;
    gw.stride = P.ws_stride;
    gw.H = P.H;
    gw.xstride = P.gx_stride / 2;
    gw.mtape = P.mtape_g + ((size_t)blockIdx.x * GW + warp) * GP * (size_t)P.H * 2 * W;

    mbar_wait(bar, 0);

    const int n = P.H * NU;
    const bool enu = (P.flags & SDEMPC_F_FRAME_ENU) != 0;
    // balanced assignment: every round spreads up to (#warps x GP) problems evenly over all warps of the grid
    const int nwarps = gridDim.x * GW, wid = blockIdx.x * GW + warp;
    for (int r0 = 0; r0 < P.B; r0 += nwarps * GP) {
        const int Br = (P.B - r0) < nwarps * GP ? (P.B - r0) : nwarps * GP;
        const int lo = (int)(((long long)wid * Br) / nwarps), hi = (int)(((long long)(wid + 1) * Br) / nwarps);
        const int b0 = r0 + lo;
        const int nprob = hi - lo;
        if (nprob <= 0) continue;
        for (int q = 0; q < nprob; ++q) {
            const int b = b0 + q;
            float* rb = gw.reg(q);
            build_window(P, lane, rb + P.o_xref, P.xref_win ? P.xref_win + (size_t)b * (P.H + 1) * NX : nullptr,
                         P.curr_t ? P.curr_t + b : nullptr, P.xdes ? P.xdes + (size_t)b * NX : nullptr, 0.f, false);
            if (P.xi_override != nullptr) {
                if (lane < P.H) {
                    const float* src = P.xi_override + ((size_t)b * P.H + lane) * 6;
#pragma unroll
                    for (int i = 0; i < 6; ++i) rb[P.o_xi + lane * 8 + i] = __ldg(src + i);
                }
            } else {
                gen_noise(lane, P.rng[2 * (size_t)b], P.rng[2 * (size_t)b + 1], 0u, 0u, P.H, rb + P.o_xi);
            }
            const float* pin = P.u_plan + (size_t)b * n;
            if (lane < NU) rb[P.o_uprev + lane] = __ldg(pin + lane);
            for (int i = lane; i < n; i += 32) {
                const int t = i / NU, ii = i % NU;
                const int ts = (P.flags & SDEMPC_F_NO_SHIFT) ? t : (t + 1 < P.H ? t + 1 : P.H - 1);
                rb[P.o_xk + i] = clipf(__ldg(pin + ts * NU + ii), P.u_lo[ii], P.u_hi[ii]);
            }
        }
        // lane q carries problem b0 + q
        const bool mine = lane < nprob;
        const int bq = b0 + (mine ? lane : 0);
        float x0[NX];
        {
            float tmp[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) tmp[i] = __ldg(P.x + (size_t)bq * NX + i);
            if (enu) enu_ned(tmp, x0);
            else {
#pragma unroll
                for (int i = 0; i < NX; ++i) x0[i] = tmp[i];
            }
        }
        if (mine) store13(gw.reg(lane) + P.o_xtape, x0);
        float s = P.info[bq].stepsize;
        s = s > 0.f ? s : P.init_step;
        __syncwarp();
        sdempc_info inf;
        g_apg_solve<NU, W, GP>(P, c, gw, s, mine, inf,
                               (P.trace != nullptr && mine) ? P.trace + (size_t)bq * P.max_iter * SDEMPC_TRACE_W : nullptr);
        if (mine) P.info_out[bq] = inf;
        for (int q = 0; q < nprob; ++q) {
            const int b = b0 + q;
            const float* rb = gw.reg(q);
            for (int i = lane; i < n; i += 32) P.u_plan_out[(size_t)b * n + i] = rb[P.o_xk + i];
            if (lane <= P.H) {
                float row[NX], o[NX];
#pragma unroll
                for (int i = 0; i < NX; ++i) row[i] = rb[P.o_xtape + lane * 16 + i] * 1.0f;   // mean over P = 1 particle
                quat_renorm(row + 6);
                if (enu) enu_ned(row, o);
                else {
#pragma unroll
                    for (int i = 0; i < NX; ++i) o[i] = row[i];
                }
                float* dst = P.x_evol + (size_t)b * (P.H + 1) * NX + lane * NX;
#pragma unroll
                for (int i = 0; i < NX; ++i) dst[i] = o[i];
            }
        }
        __syncwarp();
    }
}


template <int NU, int W, int PP, int G>
KernelChoice make_choice() {
    using L = Layout<NU, W>;
    KernelChoice k;
    k.solve = mpc_kernel<NU, W, PP, G, MODE_SOLVE>;
    k.rollout = mpc_kernel<NU, W, PP, G, MODE_ROLLOUT>;
    k.closed = mpc_kernel<NU, W, PP, G, MODE_CLOSED_LOOP>;
    k.solve_spec = nullptr;
    k.closed_spec = nullptr;
    k.solve_cl = k.closed_cl = nullptr;
    k.solve_group = nullptr;
    k.solve_pc = k.solve_pcw = nullptr;
    k.rollout_tc = k.rollout_tc_grad = k.solve_tc = k.solve_tc_lat = k.solve_tc_spec = k.solve_tc_rate = nullptr;
    k.tc_bytes = k.tc_bytes_grad = k.tc_tape_granules = k.tc_bytes_solve = k.tc_solve_tape_granules = k.tc_cols = 0;
    if constexpr (PP == 2 || PP == 4 || PP == 8) k.solve_pc = mpc_pcluster_kernel<NU, W, PP, pc_lsw(PP), pc_sgw(PP), PC_WPC>;
    if constexpr (PP == 1 || PP == 2 || PP == 4 || PP == 8) k.solve_pcw = mpc_pcluster_kernel<NU, W, PP, pc_lsw(PP), pc_sgw(PP), pcw_wpc(PP)>;
    k.gp = group_gp(NU, W);
    if constexpr (PP == 1) k.solve_group = mpc_group_kernel<NU, W, group_gp(NU, W), GROUP_GW>;
    if constexpr (PP == 1) {
        k.solve_spec = mpc_kernel<NU, W, 1, 1, MODE_SOLVE, SPEC_LSW, SPEC_SGW>;
        k.closed_spec = mpc_kernel<NU, W, 1, 1, MODE_CLOSED_LOOP, SPEC_LSW, SPEC_SGW>;
        k.solve_cl = mpc_kernel<NU, W, 1, 1, MODE_SOLVE, SPEC_LSW, SPEC_SGW, 1>;
        k.closed_cl = mpc_kernel<NU, W, 1, 1, MODE_CLOSED_LOOP, SPEC_LSW, SPEC_SGW, 1>;
    }
    k.nu = NU; k.W = W; k.P = PP; k.G = G;
    k.wimg_floats = L::TOTAL; k.wsmem_floats = L::SMEM_FLOATS; k.wreg = L::WREG;
    return k;
}


}  // namespace sdempc
