// mpc_tc.cuh — tensor-core variant of the batched forward rollout (cost evaluation), tcgen05 + TMEM.
//
// Mapping: one CTA of 128 threads owns 128 rollouts (rows = problems x particles; the particles of a problem are
// adjacent lanes, their means are warp shuffles); thread i is rollout (row) i and TMEM lane i.  The rigid
// body, the cost and the noise run per thread with the state in registers (every lane useful), and each network
// layer is ONE dense contraction over the CTA's rows on the 5th-generation tensor cores:
//     D[128 x N] (TMEM, fp32) = A[128 x K] (TMEM, written by the rows' own threads with tcgen05.st) * B[N x K]^T
// with B = the layer's weights for BOTH networks staged once in shared memory (K-major, no swizzle; layer 2 is
// block diagonal over [drift | diffusion] columns and is contracted as two W x W blocks, the output layer as one
// 16 x N12 operand with zero blocks).  The layer-1 bias rides on a constant-one
// input (its K is padded to 16 anyway); the other biases are added in the epilogue so that a CTA needs only
// 2 * N12 tensor-memory columns (128 for width 32: four CTAs per SM hide each other's MMA round trips).
// kind::tf32: operands are read as TF32 (10-bit mantissa), accumulation is fp32, so this path is NOT SPEC-ARITH: it is compared with the oracle at a stated tolerance (tests/test_gpu_parity.py,
// DESIGN.md section 5), never bit for bit.  tanh is tanh.approx.f32 (same error class as the TF32 products).
// Per step: st(z) -> MMA -> ld/tanh/st -> MMA -> ld/tanh/st -> MMA -> ld -> softplus, rigid body, cost.
// tools/tc_probe.cu is the isolated check of the descriptor / TMEM conventions used here.
#pragma once
#include <cuda_fp16.h>
#include "mpc_kernels.cuh"

namespace sdempc {

template <int NU, int W>
struct TCLayout {
    static constexpr int NIN = 6 + NU;
    static constexpr int N12 = 2 * W;                      // hidden columns: [drift units | diffusion units]
    static constexpr int K1 = ((NIN + 1 + 7) / 8) * 8;     // inputs + the constant one, padded to the MMA K (8)
    static constexpr int K2 = N12;                         // hidden units of both networks
    static constexpr int N3 = 16;                          // 6 + 6 outputs, padded
    // operand images: [rows / 8][K / 4 chunks][8 rows][16 bytes]
    static constexpr int SBO1 = (K1 / 4) * 128, SBO2 = (K2 / 4) * 128, SBOW = (W / 4) * 128, LBO = 128;
    // Layer 2 is block diagonal over [drift | diffusion]: it is stored and contracted as two W x W blocks (half the
    // shared memory and half the tensor work of the N12 x N12 form; for width 64 this is what lets two CTAs of the
    // adjoint variant share an SM).  NET2 = bytes of one block.
    static constexpr int NET2 = (W / 8) * SBOW;
    static constexpr int B1 = 0;
    static constexpr int B2 = B1 + (N12 / 8) * SBO1;
    static constexpr int B3 = B2 + 2 * NET2;
    static constexpr int BIAS2 = B3 + (N3 / 8) * SBO2;     // float b2[N12], then float b3[16]
    static constexpr int BIAS3 = BIAS2 + N12 * 4;
    static constexpr int BYTES = BIAS3 + 16 * 4;           // forward image
    // adjoint image (appended): transposed operands, no bias.  B3T [N12 rows x 16], B2T 2 x [W x W], B1T [16 x N12]
    static constexpr int SBO16 = (16 / 4) * 128;
    static constexpr int B3T = BYTES;
    static constexpr int B2T = B3T + (N12 / 8) * SBO16;
    static constexpr int B1T = B2T + 2 * NET2;
    static constexpr int BYTES_GRAD = B1T + (16 / 8) * SBO2;
    // activation / step tape of the adjoint in global memory, 16-byte granules laid out [step][granule][row]:
    // h1 (N12 halves), h2 (N12 halves), then 9 granules of floats: what the sweep reads of the step tape (r[0..2],
    // sig6, dsg6, 1/|q|, discount^t), the noise (6) and the state x_t (13).
    // The hidden activations are kept as fp16: 10 mantissa bits, what the next layer's TF32 operand read keeps of
    // them anyway (|h| <= 1, so the range is no issue); this is 37 % less tape traffic for the HBM-bound sweep.
    static constexpr int T_H1 = 0, T_H2 = N12 / 8, T_ST = 2 * N12 / 8, TG = T_ST + 9;
    // tensor-memory columns (fp32 each): D of layers 1-2 | A; the output layer's D reuses the first 16 columns
    static constexpr int C_D12 = 0, C_D3 = 0, C_A = N12;
    static constexpr int COLS = 2 * N12;                   // 128 or 256: a power of two >= 32
    static_assert(K1 <= K2 && COLS <= 512, "tensor memory columns");
    __host__ __device__ static constexpr int off(int sbo, int row, int k) { return (row / 8) * sbo + (k / 4) * 128 + (row % 8) * 16 + (k % 4) * 4; }
};

namespace tc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, int sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((128 >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
// D[tmem] (+)= A[tmem] * B[smem]^T, one K = 8 slice
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(db),
                 "r"(idesc), "r"(acc), "r"(0u));
}
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nTCW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra TCD;\nbra TCW;\nTCD:\n}\n" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, float (&o)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                 "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])));
}
__device__ __forceinline__ void st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])));
}
// all rows have written their operand: make it visible to the tensor core, then one thread issues
__device__ __forceinline__ void publish() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// eight activations <-> one 16-byte tape granule of fp16 pairs
__device__ __forceinline__ float4 pack8(const float* v) {
    const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]), c = __floats2half2_rn(v[4], v[5]),
                  d = __floats2half2_rn(v[6], v[7]);
    float4 o;
    o.x = __uint_as_float(*reinterpret_cast<const uint32_t*>(&a)); o.y = __uint_as_float(*reinterpret_cast<const uint32_t*>(&b));
    o.z = __uint_as_float(*reinterpret_cast<const uint32_t*>(&c)); o.w = __uint_as_float(*reinterpret_cast<const uint32_t*>(&d));
    return o;
}
__device__ __forceinline__ void unpack8(const float4 g, float (&h)[8]) {
    const uint32_t w[4] = {__float_as_uint(g.x), __float_as_uint(g.y), __float_as_uint(g.z), __float_as_uint(g.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        h[2 * i] = f.x; h[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_plain(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nTMW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra TMD;\nbra TMW;\nTMD:\n}\n" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
}  // namespace tc

// reference row t (internal frame) of problem b: explicit window, trajectory time or fixed set-point
__device__ __forceinline__ void tc_ref_row(const KParams& P, int b, int t, float (&row)[NX]) {
    const bool enu = (P.flags & SDEMPC_F_FRAME_ENU) != 0;
    if (P.xref_win != nullptr || P.xdes != nullptr) {
        const float* src = P.xref_win ? P.xref_win + ((size_t)b * (P.H + 1) + t) * NX : P.xdes + (size_t)b * NX;
        float tmp[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) tmp[i] = __ldg(src + i);
        if (enu) enu_ned(tmp, row);
        else {
#pragma unroll
            for (int i = 0; i < NX; ++i) row[i] = tmp[i];
        }
    } else {
        float tt = __ldg(P.curr_t + b);
        for (int s = 0; s < t; ++s) tt = tt + P.dt[s];
        traj_interp(P.traj, P.T, tt, row);
    }
}

// Body of the kernel (entry point in sdempc_tc.cu).  wimg: TCLayout image in global memory.  bar: two mbarriers.
template <int NU, int W, bool GRAD>
__device__ __forceinline__ void tc_rollout_body(const KParams& P, unsigned char* sB, uint32_t* tmem_base_slot, uint64_t* bar) {
    using L = TCLayout<NU, W>;
    constexpr int NIN = L::NIN, N12 = L::N12;
    constexpr int IMG = GRAD ? L::BYTES_GRAD : L::BYTES;
    const int tid = threadIdx.x, warp = tid >> 5;
    // ---- one-time setup: weights to shared memory by a TMA bulk copy (SASS UBLKCP), barriers, tensor memory ----
    uint64_t* wbar = bar + 1;            // bar[0]: contraction complete, bar[1]: weight image landed
    if (tid == 0) {
        tc::mbar_init(bar, 1);
        tc::mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        constexpr uint32_t CH = 32768;
        tc::mbar_expect_tx(wbar, IMG);
        for (uint32_t off = 0; off < (uint32_t)IMG; off += CH)
            tc::bulk_g2s(sB + off, reinterpret_cast<const unsigned char*>(P.wimg) + off, ((uint32_t)IMG - off) < CH ? ((uint32_t)IMG - off) : CH, wbar);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_base_slot)), "r"((uint32_t)L::COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tc::mbar_wait_plain(wbar, 0);        // the async proxy wrote what tcgen05.mma (async proxy) and the bias loads read
    const uint32_t tb = *tmem_base_slot;
    const uint32_t lane_addr = tb + ((uint32_t)(warp * 32) << 16);
    const uint32_t sb = tc::smem_u32(sB);
    const uint32_t id12 = tc::idesc_tf32(128, N12), id3 = tc::idesc_tf32(128, L::N3), idW = tc::idesc_tf32(128, W);
    uint32_t phase = 0;

    // row = problem * particles + particle: the particles of a problem are adjacent lanes of one warp (the
    // particle count is a power of two <= 32), so the particle means are warp shuffles in particle order
    const int pp = P.P, lane = tid & 31, pbase = lane & ~(pp - 1);
    const int row = blockIdx.x * 128 + tid, nrows = P.B * pp;
    const bool valid = row < nrows;
    const int rr = valid ? row : nrows - 1;   // rows past the batch shadow the last row and store nothing
    const int b = rr / pp, part = rr - b * pp;
    const bool writer = valid && part == 0;
    const float invP = __fdiv_rn(1.0f, (float)pp);
    auto pmean = [&](float v) -> float {      // SPEC: sum over the particles in index order, times 1/P
        if (pp == 1) return v;
        float s = __shfl_sync(0xffffffffu, v, pbase);
        for (int p = 1; p < pp; ++p) s = s + __shfl_sync(0xffffffffu, v, pbase + p);
        return s * invP;
    };
    const bool enu = (P.flags & SDEMPC_F_FRAME_ENU) != 0;
    float x[NX];
    {
        float tmp[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) tmp[i] = __ldg(P.x + (size_t)b * NX + i);
        if (enu) enu_ned(tmp, x);
        else {
#pragma unroll
            for (int i = 0; i < NX; ++i) x[i] = tmp[i];
        }
    }
    if (writer && P.x_evol != nullptr) {
#pragma unroll
        for (int i = 0; i < NX; ++i) P.x_evol[(size_t)b * (P.H + 1) * NX + i] = __ldg(P.x + (size_t)b * NX + i);
    }
    float up[NU], u[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) up[i] = __ldg(P.uprev_in + (size_t)b * NU + i);
    const float* bias2 = reinterpret_cast<const float*>(sB + L::BIAS2);
    const float* bias3 = reinterpret_cast<const float*>(sB + L::BIAS3);
    // this thread's float4 slot of tape granule g at step t: coalesced over the CTA's rows
    float4* tape = reinterpret_cast<float4*>(P.mtape_g) + (size_t)blockIdx.x * P.H * L::TG * 128 + tid;
    auto tp = [&](int t, int g) -> float4* { return tape + ((size_t)t * L::TG + g) * 128; };
    const unsigned long long seed = P.rng ? P.rng[2 * (size_t)b] : 0ull, tick = P.rng ? P.rng[2 * (size_t)b + 1] : 0ull;
    float Jp = 0.f, disc = 1.f;
    for (int t = 0; t < P.H; ++t) {
#pragma unroll
        for (int i = 0; i < NU; ++i) u[i] = __ldg(P.u_in + ((size_t)b * P.H + t) * NU + i);
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
        // ---- tanh -> layer 2 operand ----
#pragma unroll
        for (int c0 = 0; c0 < N12; c0 += 16) {
            float v[16];
            tc::ld16(lane_addr + L::C_D12 + c0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = tc::tanh_approx(v[i]);
            tc::st16(lane_addr + L::C_A + c0, v);
            if constexpr (GRAD) {
                *tp(t, L::T_H1 + c0 / 8) = tc::pack8(v);
                *tp(t, L::T_H1 + c0 / 8 + 1) = tc::pack8(v + 8);
            }
        }
        tc::publish();
        if (tid == 0) {
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int k8 = 0; k8 < W / 8; ++k8)
                    tc::mma_ts(tb + L::C_D12 + n * W, tb + L::C_A + n * W + 8 * k8,
                               tc::desc(sb + L::B2 + n * L::NET2 + k8 * 2 * L::LBO, L::SBOW), idW, k8 > 0);
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
            if constexpr (GRAD) {
                *tp(t, L::T_H2 + c0 / 8) = tc::pack8(v);
                *tp(t, L::T_H2 + c0 / 8 + 1) = tc::pack8(v + 8);
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
                det_softplus_sigmoid_opt(o[6 + i] + bias3[6 + i], sp, sg, GRAD);
                sig[i] = P.sig0[i] * sp;
                dsg[i] = P.sig0[i] * sg;
            }
        }
        // ---- noise, reference row, rigid body + Euler-Maruyama + cost ----
        float xi[6], xr[NX], xn[NX], rn;
        if (P.xi_override != nullptr) {
#pragma unroll
            for (int i = 0; i < 6; ++i) xi[i] = __ldg(P.xi_override + ((size_t)rr * P.H + t) * 6 + i);
        } else {
            const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
            const uint32_t c2 = (uint32_t)tick, c3 = ((uint32_t)(tick >> 32)) & 0x3FFFFFFFu;
            uint32_t r[4];
            philox4x32_10((uint32_t)t, (uint32_t)part, c2, c3, k0, k1, r);
            box_muller(r[0], r[1], xi[0], xi[1]);
            box_muller(r[2], r[3], xi[2], xi[3]);
            philox4x32_10((uint32_t)t, (uint32_t)part, c2, c3 | (1u << 30), k0, k1, r);
            box_muller(r[0], r[1], xi[4], xi[5]);
        }
        tc_ref_row(P, b, t + 1, xr);
        const float l = phys_step<NU>(P, t, x, u, up, r6, sig, xi, xr, xn, rn);
        if constexpr (GRAD) {
            *tp(t, L::T_ST) = make_float4(r6[0], r6[1], r6[2], sig[0]);
            *tp(t, L::T_ST + 1) = make_float4(sig[1], sig[2], sig[3], sig[4]);
            *tp(t, L::T_ST + 2) = make_float4(sig[5], dsg[0], dsg[1], dsg[2]);
            *tp(t, L::T_ST + 3) = make_float4(dsg[3], dsg[4], dsg[5], rn);
            *tp(t, L::T_ST + 4) = make_float4(disc, xi[0], xi[1], xi[2]);
            *tp(t, L::T_ST + 5) = make_float4(xi[3], xi[4], xi[5], x[0]);      // x is still x_t here
            *tp(t, L::T_ST + 6) = make_float4(x[1], x[2], x[3], x[4]);
            *tp(t, L::T_ST + 7) = make_float4(x[5], x[6], x[7], x[8]);
            *tp(t, L::T_ST + 8) = make_float4(x[9], x[10], x[11], x[12]);
        }
        Jp = fma_(disc, l, Jp);
        disc = disc * P.discount;
#pragma unroll
        for (int i = 0; i < NU; ++i) up[i] = u[i];
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = xn[i];
        if (P.x_evol != nullptr) {            // predicted mean trajectory: particle mean, quaternion renormalised
            float rowv[NX], o[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) rowv[i] = pmean(xn[i]);
            quat_renorm(rowv + 6);
            if (enu) enu_ned(rowv, o);
            else {
#pragma unroll
                for (int i = 0; i < NX; ++i) o[i] = rowv[i];
            }
            if (writer) {
#pragma unroll
                for (int i = 0; i < NX; ++i) P.x_evol[((size_t)b * (P.H + 1) + t + 1) * NX + i] = o[i];
            }
        }
    }
    {
        const float Jm = pmean(Jp);
        if (writer) P.cost_out[b] = Jm;
    }
    if constexpr (GRAD) {
        // ---- adjoint sweep: three transposed contractions per step (W3^T, W2^T, W1^T), tapes from global memory ----
        const uint32_t id16 = tc::idesc_tf32(128, 16);
        float lam[NX], gp[NU], xn[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) { lam[i] = 0.f; xn[i] = x[i]; }   // x holds x_H after the forward loop
#pragma unroll
        for (int i = 0; i < NU; ++i) gp[i] = 0.f;
        for (int t = P.H - 1; t >= 0; --t) {
            // the sweep consumes its tape as it loads it (57 % of the stall samples were these loads): pull the previous
            // step's granules into L2 while this step's contractions run (0.600 -> 0.576 ms; two or three steps ahead: no gain)
            if (t > 0) {
#pragma unroll
                for (int g = 0; g < L::TG; ++g) tc::prefetch_l2(tp(t - 1, g));
            }
            float xt[NX], xi[6];
            const float4 s0 = *tp(t, L::T_ST), s1 = *tp(t, L::T_ST + 1), s2 = *tp(t, L::T_ST + 2), s3 = *tp(t, L::T_ST + 3),
                         s4 = *tp(t, L::T_ST + 4);
            {
                const float4 a = *tp(t, L::T_ST + 5), c = *tp(t, L::T_ST + 6), d = *tp(t, L::T_ST + 7), e = *tp(t, L::T_ST + 8);
                xi[0] = s4.y; xi[1] = s4.z; xi[2] = s4.w; xi[3] = a.x; xi[4] = a.y; xi[5] = a.z;
                xt[0] = a.w; xt[1] = c.x; xt[2] = c.y; xt[3] = c.z; xt[4] = c.w; xt[5] = d.x; xt[6] = d.y; xt[7] = d.z; xt[8] = d.w;
                xt[9] = e.x; xt[10] = e.y; xt[11] = e.z; xt[12] = e.w;
            }
#pragma unroll
            for (int i = 0; i < NU; ++i) u[i] = __ldg(P.u_in + ((size_t)b * P.H + t) * NU + i);
#pragma unroll
            for (int i = 0; i < NU; ++i)
                up[i] = (t == 0) ? __ldg(P.uprev_in + (size_t)b * NU + i) : __ldg(P.u_in + ((size_t)b * P.H + t - 1) * NU + i);
            BwdMid mid;
            float gu[NU];
            float2 lo[6];
            {
                float xr[NX];
                tc_ref_row(P, b, t + 1, xr);
                const float r012[3] = {s0.x, s0.y, s0.z};
                const float sg6[6] = {s0.w, s1.x, s1.y, s1.z, s1.w, s2.x};
                const float dsg[6] = {s2.y, s2.z, s2.w, s3.x, s3.y, s3.z};
                const float rn = s3.w, dsc = s4.x;
                bwd_pre<NU>(P, t, xt, xn, xr, u, r012, sg6, dsg, rn, dsc, xi, lam, mid, lo, gu);
            }
            {   // output adjoints -> A[.., 16]: drift rows 0..5, diffusion rows 6..11
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
            // the hidden-activation granules are loaded one 16-column chunk ahead of their use, the first chunk while
            // the contraction is in flight (0.486 -> 0.472 ms per 65 536-row value_and_grad)
            float4 g0_T_H2 = *tp(t, L::T_H2), g1_T_H2 = *tp(t, L::T_H2 + 1);
            tc::wait(bar, phase); phase ^= 1;
#pragma unroll
            for (int c0 = 0; c0 < N12; c0 += 16) {
                float4 n0 = g0_T_H2, n1 = g1_T_H2;
                if (c0 + 16 < N12) { n0 = *tp(t, L::T_H2 + (c0 + 16) / 8); n1 = *tp(t, L::T_H2 + (c0 + 16) / 8 + 1); }
                float v[16];
                tc::ld16(lane_addr + L::C_D12 + c0, v);
                {
                    float h[8];
                    tc::unpack8(g0_T_H2, h);
#pragma unroll
                    for (int m = 0; m < 8; ++m) v[m] = v[m] * fma_(-h[m], h[m], 1.f);
                    tc::unpack8(g1_T_H2, h);
#pragma unroll
                    for (int m = 0; m < 8; ++m) v[8 + m] = v[8 + m] * fma_(-h[m], h[m], 1.f);
                }
                tc::st16(lane_addr + L::C_A + c0, v);
                g0_T_H2 = n0; g1_T_H2 = n1;
            }
            tc::publish();
            if (tid == 0) {
#pragma unroll
                for (int n = 0; n < 2; ++n)
#pragma unroll
                    for (int k8 = 0; k8 < W / 8; ++k8)
                        tc::mma_ts(tb + L::C_D12 + n * W, tb + L::C_A + n * W + 8 * k8,
                                   tc::desc(sb + L::B2T + n * L::NET2 + k8 * 2 * L::LBO, L::SBOW), idW, k8 > 0);
                tc::commit(bar);
            }
            float4 g0_T_H1 = *tp(t, L::T_H1), g1_T_H1 = *tp(t, L::T_H1 + 1);
            tc::wait(bar, phase); phase ^= 1;
#pragma unroll
            for (int c0 = 0; c0 < N12; c0 += 16) {
                float4 n0 = g0_T_H1, n1 = g1_T_H1;
                if (c0 + 16 < N12) { n0 = *tp(t, L::T_H1 + (c0 + 16) / 8); n1 = *tp(t, L::T_H1 + (c0 + 16) / 8 + 1); }
                float v[16];
                tc::ld16(lane_addr + L::C_D12 + c0, v);
                {
                    float h[8];
                    tc::unpack8(g0_T_H1, h);
#pragma unroll
                    for (int m = 0; m < 8; ++m) v[m] = v[m] * fma_(-h[m], h[m], 1.f);
                    tc::unpack8(g1_T_H1, h);
#pragma unroll
                    for (int m = 0; m < 8; ++m) v[8 + m] = v[8 + m] * fma_(-h[m], h[m], 1.f);
                }
                tc::st16(lane_addr + L::C_A + c0, v);
                g0_T_H1 = n0; g1_T_H1 = n1;
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
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                const float gm = pmean(gu[i]);
                if (writer) P.grad_out[((size_t)b * P.H + t) * NU + i] = gm;
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) xn[i] = xt[i];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"((uint32_t)L::COLS));
}

}  // namespace sdempc
