// mpc_pcluster.cuh — latency kernel for P > 1 particles: one problem per thread-block cluster.
//
// The P particle rollouts of a problem are independent given the control sequence, and so are the line-search
// trials of an iteration.  A cluster of P*(LSW+SGW)/WPC CTAs (8 = the portable size, or 16 where the device takes it; WPC = 4 warps
// each, one per SM sub-partition, so that no two warps share an issue port — or WPC = 2 for width 64, whose weights are
// read from shared memory every step: four warps keep an SM's shared-memory pipe 69 % busy and wait on it) holds LSW + SGW replicas x P particles of ONE problem: warp (l, p) integrates particle p.
// Replicas l < LSW evaluate line-search trial base + l.  Replicas l >= LSW are SPECULATIVE-GRADIENT replicas (as in
// apg_solve_latency for P = 1): while the trials run, replica LSW + c computes value_and_grad at the next extrapolation
// point for one possible outcome of the search — an accepted trial (SGW = 4: trials 0, 1, 2; SGW = 2: trial 1, the
// outcome of 61 % of the iterations of the benchmark problems) or, the last candidate, the REJECTED step, whose restart
// point x_k is known before the search (24-30 % of the iterations).  On a hit the next iteration starts with its
// gradient known and the critical path of an iteration is one value_and_grad (40 step evaluations) instead of
// value_and_grad + trial (60); on a miss every replica computes the gradient as before.
// Particle means (cost: sequential sum in particle order times 1/P; gradient likewise) and the per-trial
// (J, slope) pairs are exchanged through distributed shared memory, one cluster barrier per exchange.
// Candidates are built with the expressions of the real update and every replica takes the same decisions from the same
// exchanged values: same SPEC-ARITH sequences as every other kernel, bit-identical results.
#pragma once
#include <cooperative_groups.h>

#include "mpc_kernels.cuh"

namespace sdempc {

template <int PP, int LSW, int SGW, int WPC = 4>
struct PCluster {
    static constexpr int TW = PP * (LSW + SGW);   // warps of the team (<= 64: two exchange values per lane)
    static_assert(TW <= 64 && TW % WPC == 0 && 32 % PP == 0, "team of at most 64 warps, whole CTAs, replicas inside a half");
    static constexpr int CS = TW / WPC;   // CTAs of the cluster
    int l, p, gwi;                        // replica, particle, warp index in the team
    float* xc_local;                      // this CTA's exchange area: [2 parities][WPC warps][2 floats]
    float* warp_base_local;               // this CTA's per-warp regions
    int ws_stride;
    __device__ __forceinline__ void barrier() const {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    // exchange slot / region of team warp `w` (possibly in another CTA of the cluster)
    __device__ __forceinline__ const float* slot(int w, int parity) const {
        const float* loc = xc_local + (parity * WPC + (w % WPC)) * 2;
        return cooperative_groups::this_cluster().map_shared_rank(loc, w / WPC);
    }
    __device__ __forceinline__ const float* region(int w) const {
        const float* loc = warp_base_local + (size_t)(w % WPC) * ws_stride;
        return cooperative_groups::this_cluster().map_shared_rank(loc, w / WPC);
    }
};

// publish (a, b) of this warp, barrier, return every team warp's pair: lane w holds warp w's pair in (.x, .y) and warp
// (w + 32)'s in (.z, .w)
template <int PP, int LSW, int SGW, int WPC>
__device__ __forceinline__ float4 pc_exchange(const PCluster<PP, LSW, SGW, WPC>& pc, int lane, int warp_in_cta, int parity, float a, float b) {
    constexpr int TW = PCluster<PP, LSW, SGW, WPC>::TW;
    if (lane == 0) {
        float* s = pc.xc_local + (parity * WPC + warp_in_cta) * 2;
        s[0] = a; s[1] = b;
    }
    pc.barrier();
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < TW) { const float* s = pc.slot(lane, parity); v.x = s[0]; v.y = s[1]; }
    if (TW > 32 && lane + 32 < TW) { const float* s = pc.slot(lane + 32, parity); v.z = s[0]; v.w = s[1]; }
    return v;
}

// mean over the particles of replica r of the first components (sequential sum in particle order, times 1/P); the PP
// warps of a replica never straddle the two halves of the exchange (PP divides 32)
template <int PP>
__device__ __forceinline__ float pc_replica_mean(const float4& v, int r, float invP) {
    const int w0 = r * PP, l0 = w0 & 31;
    const float src = w0 < 32 ? v.x : v.z;
    float acc = __shfl_sync(0xffffffffu, src, l0);
#pragma unroll
    for (int q = 1; q < PP; ++q) acc = acc + __shfl_sync(0xffffffffu, src, l0 + q);
    return acc * invP;
}
// second component published by the first warp of replica r
template <int PP>
__device__ __forceinline__ float pc_replica_second(const float4& v, int r) {
    const int w0 = r * PP;
    return __shfl_sync(0xffffffffu, w0 < 32 ? v.y : v.w, w0 & 31);
}

template <int NU, int W, int PP, int LSW, int SGW, int WPC>
__device__ __forceinline__ void apg_solve_pcluster(const KParams& P, Warp<NU, W>& c, const PCluster<PP, LSW, SGW, WPC>& pc, int warp_in_cta,
                                                   const float (&x0)[NX], float s, sdempc_info& inf, float* trace) {
    const int lane = c.lane;
    const int n = P.H * NU;
    const float invP = __fdiv_rn(1.0f, (float)PP);
    const int l = pc.l;
    const bool is_spec = l >= LSW;
    const int cand = l - LSW;                                  // speculation candidate of this replica (if is_spec)
    auto cand_trial = [](int cc) { return SGW == 2 ? 1 : cc; };   // the accepted trial candidate cc < SGW - 1 stands for
    for (int i = lane; i < n; i += 32) c.yk[i] = c.xk[i];
    __syncwarp();
    float Jx = 0.f, Jp = 0.f, fy = 0.f, gsq = 0.f, sum_ls = 0.f, sum_s = 0.f, init_cost = 0.f;
    float fc[SGW > 0 ? SGW : 1];                               // particle-mean cost at each candidate point
    int k = 1, no_improve = 0, it = 0, xpar = 0;
    bool boot = true;
    // particle mean of the per-warp gradients (in g2) of replica r -> this warp's c.g
    auto mean_grad = [&](int r) {
        for (int i = lane; i < n; i += 32) {
            float a = pc.region(r * PP)[P.o_g2 + i];
#pragma unroll
            for (int q = 1; q < PP; ++q) a = a + pc.region(r * PP + q)[P.o_g2 + i];
            c.g[i] = a * invP;
        }
        __syncwarp();
    };
    for (;;) {
        if (boot) {   // gradient at y_k by every replica (first iteration, or a speculation miss): this warp's particle, then the mean
            ++it;
            float* gsave = c.g;
            c.g = c.g2;
            const float Jw = rollout_fwd<NU, W, 1>(P, c, c.yk, x0);
            rollout_bwd<NU, W>(P, c, c.yk);
            c.g = gsave;
            const float4 v = pc_exchange<PP, LSW, SGW, WPC>(pc, lane, warp_in_cta, xpar, Jw, 0.f);
            xpar ^= 1;
            fy = pc_replica_mean<PP>(v, l, invP);
            mean_grad(l);
            if (it == 1) { Jx = fy; init_cost = fy; }
            if (SGW > 0) pc.barrier();   // a speculation warp overwrites its g2 next; its replica's other warps may still be reading it
        }
        {
            float part = 0.f;
            for (int i = lane; i < n; i += 32) { const float gi = c.g[i]; part = fma_(gi, gi, part); }
            gsq = warp_butterfly(part);
        }
        if (P.reset_option == 1) { s = s * P.inc_f; s = s > P.max_step ? P.max_step : s; }
        const float s0 = s;
        const float beta = apg_momentum_tab(P, k);
        // ---- speculation replicas: the candidate next point (expressions of the real update below) ----
        bool spec_valid = false;
        if (SGW > 0 && is_spec) {
            const bool rej = (cand == SGW - 1);
            const int jc = cand_trial(cand);
            if (rej || jc <= P.maxls) {
                float s_c = s0;
                for (int q = 0; q < jc && !rej; ++q) s_c = s_c * P.dec_f;
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
        bool ok = false;
        int jsel = 0, base = 0;
        float s_base = s0;
        for (int round = 0;; ++round) {   // rounds of LSW concurrent trials; the speculations run next to the first
            const int j = base + l;
            float Jw = 0.f, dec = 0.f;
            if (!is_spec && j <= P.maxls) {
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
                Jw = rollout_fwd<NU, W, 0>(P, c, c.xp, x0);
            }
            if (SGW > 0 && is_spec && spec_valid && round == 0) {   // speculative value_and_grad at c.xp -> c.g2
                float* gsave = c.g;
                c.g = c.g2;
                Jw = rollout_fwd<NU, W, 1>(P, c, c.xp, x0);
                rollout_bwd<NU, W>(P, c, c.xp);
                c.g = gsave;
            }
            const float4 v = pc_exchange<PP, LSW, SGW, WPC>(pc, lane, warp_in_cta, xpar, Jw, dec);
            xpar ^= 1;
            if (round == 0) {
#pragma unroll
                for (int cc = 0; cc < SGW; ++cc) fc[cc] = pc_replica_mean<PP>(v, LSW + cc, invP);
            }
            int q = 0;
            float Jq = 0.f;
            for (; q < LSW && base + q <= P.maxls; ++q) {
                Jq = pc_replica_mean<PP>(v, q, invP);
                const float dq = pc_replica_second<PP>(v, q);
                if (Jq <= fma_(P.coef, dq, fy)) { ok = true; break; }
            }
            if (ok) { jsel = base + q; Jp = Jq; break; }
            if (base + LSW > P.maxls) { jsel = P.maxls; Jp = Jq; break; }   // every trial failed: Jq is the last trial's cost
            for (int r = 0; r < LSW; ++r) s_base = s_base * P.dec_f;
            base += LSW;
        }
        s = s0;
        for (int q = 0; q < jsel; ++q) s = s * P.dec_f;
        const int n_ls = jsel + 1;
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
        if (trace != nullptr && pc.gwi == 0 && lane == 0) {
            float* tr = trace + (size_t)(it - 1) * SDEMPC_TRACE_W;
            tr[0] = fy; tr[1] = Jp; tr[2] = s; tr[3] = (float)n_ls; tr[4] = accept ? 1.f : 0.f; tr[5] = Jx; tr[6] = gsq; tr[7] = (float)k;
        }
        if (it >= P.max_iter || no_improve >= P.max_no_improve || converged || !(fy == fy)) break;
        // ---- did a speculation replica already compute the gradient at the new y_k? ----
        int hit = -1;
        if (SGW > 0) {
            if (!accept) hit = SGW - 1;
            else {
#pragma unroll
                for (int cc = 0; cc < SGW - 1; ++cc)
                    if (jsel == cand_trial(cc)) hit = cc;
            }
        }
        if (hit >= 0) {
            ++it;
            fy = fc[hit];
            mean_grad(LSW + hit);
            boot = false;
        } else {
            boot = true;
        }
        if (SGW > 0) pc.barrier();   // the speculation replicas' g2 may be overwritten by the next pass
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

}  // namespace sdempc
