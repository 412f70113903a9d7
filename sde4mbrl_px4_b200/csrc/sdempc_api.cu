// sdempc_api.cu — __global__ entry points and the C ABI of include/sdempc.h.
//
// Build (see __graft_entry__.build()):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
//        -shared -Xcompiler -fPIC -cudart static sdempc_api.cu -o ../libsdempc.so
// -fmad=false is REQUIRED: every fused multiply-add of SPEC-ARITH is written
// explicitly (__fmaf_rn / __ffma2_rn); nothing else may be contracted.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mpc_entry.cuh"
#include "tc_api.h"

using namespace sdempc;

// FP32-pipe probe: 8 independent FFMA chains per thread, 16 warps per SM (4 per sub-partition).  The roofline
// denominator of bench.py: what the FMA pipe of THIS device sustains at the clocks it runs at, measured in the same run.
__global__ void __launch_bounds__(512, 1) fp32_probe_kernel(float* out, int iters) {
    float a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 1.0f + threadIdx.x * 1e-6f + i; b[i] = 0.999f + i * 1e-7f; c[i] = 0.5f * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep)
#pragma unroll
            for (int i = 0; i < 8; ++i) c[i] = __fmaf_rn(a[i], b[i], c[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// =====================================================================================
// host side
// =====================================================================================
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) return fail(SDEMPC_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

// Compiled (nu, width, particles, teams-per-CTA) combinations: one table entry per instantiation (kern_*.cu), the
// tensor-core kernels of the (nu, width) added from tc_api.h.
namespace sdempc {
KernelChoice choice_4_32_1(); KernelChoice choice_4_32_2(); KernelChoice choice_4_32_4(); KernelChoice choice_4_32_8();
KernelChoice choice_6_32_1(); KernelChoice choice_6_32_8(); KernelChoice choice_4_64_1(); KernelChoice choice_6_64_1();
KernelChoice choice_6_64_8();
}
static const std::vector<KernelChoice>& choices() {
    static const std::vector<KernelChoice> v = [] {
        std::vector<KernelChoice> c = {
#ifdef SDEMPC_DEV_IRIS_ONLY   // quick experiment builds (link kern_a.o only): the bench shape only
            choice_4_32_1(),
#else
            choice_4_32_1(), choice_4_32_2(), choice_4_32_4(), choice_4_32_8(), choice_6_32_1(), choice_6_32_8(),
            choice_4_64_1(), choice_6_64_1(), choice_6_64_8(),
#endif
        };
        for (auto& k : c)
            if (const TCKernels* t = tc_kernels(k.nu, k.W)) {   // one set of tensor-core kernels per (nu, width): the particle count is a run-time row mapping
                k.rollout_tc = t->rollout; k.rollout_tc_grad = t->rollout_grad; k.solve_tc = t->solve; k.solve_tc_lat = t->solve_lat; k.solve_tc_spec = t->solve_spec; k.solve_tc_rate = t->solve_rate;
                k.tc_bytes = t->bytes; k.tc_bytes_grad = t->bytes_grad; k.tc_bytes_solve = t->bytes_solve;
                k.tc_tape_granules = t->tape_granules; k.tc_solve_tape_granules = t->solve_tape_granules; k.tc_cols = t->cols;
            }
        return c;
    }();
    return v;
}

struct sdempc_handle {
    sdempc_config cfg;
    sdempc_model_header mh;
    std::vector<float> weights;       // raw blob payload
    std::vector<float> wimg;          // packed image
    std::vector<float> traj_ext, traj_int;
    int T = 0;
    int device = 0;
    KernelChoice kc;
    KParams kp;                       // template (config + model + layout)
    KParams kp_group;                 // same with the per-problem layout of the group kernel
    std::vector<float> wimg_tc;       // tensor-core operand image (TCLayout), forward part then adjoint part
    float* d_wimg_tc = nullptr;
    float* d_tape_tc = nullptr; size_t tape_tc_bytes = 0;
    float* d_tcs_ws = nullptr; size_t tcs_ws_bytes = 0;   // per-CTA workspaces of the tensor-core solve
    int tcs_ppc_override = 0;                             // experiments: SDEMPC_TC_PPC
    int tcs_spec = 1;                                     // SDEMPC_TC_SPEC=0: never the speculative build (tests compare the two); 2: only up to slots / 8
    size_t smem_bytes_group = 0;
    size_t smem_bytes = 0, smem_bytes_spec = 0, smem_bytes_cl = 0;
    // lazily created device state
    bool dev_ready = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float* d_wimg = nullptr;
    float* d_beta = nullptr;          // momentum table
    float* d_traj = nullptr;
    float2* d_mtape = nullptr;
    size_t mtape_warps = 0;
    float2* d_mtape_group = nullptr;
    size_t mtape_group_n = 0;
    char* d_in = nullptr; char* h_in = nullptr; size_t in_cap = 0;
    char* d_out = nullptr; char* h_out = nullptr; size_t out_cap = 0;
    float* d_trace = nullptr; size_t trace_cap = 0;
    char* d_flush = nullptr;
    int sm_count = 0;
    int64_t launches = 0;
    int last_grid = 0;
    int regs = 0, regs_tc = 0, regs_tc_lat = 0, regs_tc_spec = 0, regs_tc_rate = 0;
    int pcw_clusters = 0;                                 // clusters of the wide shape of the cluster latency kernel the device holds at once
    // staged launch
    KParams staged;
    int staged_B = 0;
    bool staged_ok = false, staged_spec = false, staged_group = false, staged_cl = false, staged_pc = false, staged_pcw = false, staged_direct_in = false, staged_tc = false, staged_tc_lat = false, staged_tc_spec = false, staged_tc_rate = false;
    float last_ms = 0.f;
};

static void pack_weights(sdempc_handle* h) {
    const int NU = h->mh.nu, W = h->mh.width, NIN = 6 + NU;
    const int PAIR = 2 * W + 4;
    const int W2T = 0, W1T = W2T + W * PAIR, W3RS = W + 4, W3R = W1T + NIN * PAIR, B3 = W3R + 12 * W3RS, PART_A = B3 + 16;
    const int W1PS = (2 * NIN) % 8 == 4 ? 2 * NIN : 2 * NIN + 4;
    const int W1P = PART_A, W2P = W1P + W * W1PS, W3C = W2P + W * PAIR, B1 = W3C + W * 12, B2 = B1 + 2 * W, TOTAL = B2 + 2 * W;
    h->wimg.assign(TOTAL, 0.f);
    const float* p = h->weights.data();
    const float *W1[2], *b1[2], *W2[2], *b2[2], *W3[2], *b3[2];
    for (int n = 0; n < 2; ++n) {
        W1[n] = p; p += (size_t)W * NIN;
        b1[n] = p; p += W;
        W2[n] = p; p += (size_t)W * W;
        b2[n] = p; p += W;
        W3[n] = p; p += 6 * (size_t)W;
        b3[n] = p; p += 6;
    }
    float* I = h->wimg.data();
    for (int n = 0; n < 2; ++n) {
        for (int k = 0; k < W; ++k)
            for (int j = 0; j < W; ++j) I[W2T + k * PAIR + 2 * j + n] = W2[n][j * W + k];
        for (int i = 0; i < NIN; ++i)
            for (int j = 0; j < W; ++j) I[W1T + i * PAIR + 2 * j + n] = W1[n][j * NIN + i];
        for (int o = 0; o < 6; ++o) {
            for (int k = 0; k < W; ++k) I[W3R + (6 * n + o) * W3RS + k] = W3[n][o * W + k];
            I[B3 + 6 * n + o] = b3[n][o];
        }
        for (int j = 0; j < W; ++j) {
            for (int k = 0; k < NIN; ++k) I[W1P + j * W1PS + 2 * k + n] = W1[n][j * NIN + k];
            for (int k = 0; k < W; ++k) I[W2P + j * PAIR + 2 * k + n] = W2[n][j * W + k];
            for (int o = 0; o < 6; ++o) I[W3C + j * 12 + 2 * o + n] = W3[n][o * W + j];
            I[B1 + 2 * j + n] = b1[n][j];
            I[B2 + 2 * j + n] = b2[n][j];
        }
    }
}

// Operand image of the tensor-core path (TCLayout in mpc_tc.cuh): K-major rows of [W1 | b1] (b1 multiplies a
// constant-one input), W2 as two W x W blocks (drift, diffusion), W3 block diagonal over the [drift | diffusion]
// hidden columns, then b2 and b3.
static void pack_weights_tc(sdempc_handle* h) {
    const int NU = h->mh.nu, W = h->mh.width, NIN = 6 + NU, N12 = 2 * W;
    const int K1 = ((NIN + 1 + 7) / 8) * 8, K2 = N12, N3 = 16;
    const int SBO1 = (K1 / 4) * 128, SBO2 = (K2 / 4) * 128, SBOW = (W / 4) * 128, NET2 = (W / 8) * SBOW;
    const int B1 = 0, B2 = B1 + (N12 / 8) * SBO1, B3 = B2 + 2 * NET2, BIAS2 = B3 + (N3 / 8) * SBO2, BIAS3 = BIAS2 + N12 * 4,
              BYTES = BIAS3 + 16 * 4;
    auto off = [](int sbo, int row, int k) { return ((row / 8) * sbo + (k / 4) * 128 + (row % 8) * 16 + (k % 4) * 4) / 4; };
    const int SBO16 = (16 / 4) * 128, B3T = BYTES, B2T = B3T + (N12 / 8) * SBO16, B1T = B2T + 2 * NET2,
              BYTES_GRAD = B1T + (16 / 8) * SBO2;
    h->wimg_tc.assign(BYTES_GRAD / 4, 0.f);
    float* I = h->wimg_tc.data();
    const float* p = h->weights.data();
    for (int n = 0; n < 2; ++n) {
        const float* W1 = p; p += (size_t)W * NIN;
        const float* b1 = p; p += W;
        const float* W2 = p; p += (size_t)W * W;
        const float* b2 = p; p += W;
        const float* W3 = p; p += 6 * (size_t)W;
        const float* b3 = p; p += 6;
        for (int j = 0; j < W; ++j) {
            const int row = n * W + j;
            for (int k = 0; k < NIN; ++k) I[B1 / 4 + off(SBO1, row, k)] = W1[j * NIN + k];
            I[B1 / 4 + off(SBO1, row, NIN)] = b1[j];   // multiplies the constant-one input
            for (int k = 0; k < W; ++k) I[(B2 + n * NET2) / 4 + off(SBOW, j, k)] = W2[j * W + k];   // block n of layer 2
            I[BIAS2 / 4 + row] = b2[j];
            // adjoint operands: row = the hidden column that receives the adjoint
            for (int o = 0; o < 6; ++o) I[B3T / 4 + off(SBO16, row, n * 6 + o)] = W3[o * W + j];          // (W3^T)[j][o]
            for (int k = 0; k < W; ++k) I[(B2T + n * NET2) / 4 + off(SBOW, j, k)] = W2[k * W + j];         // (W2^T)[j][k], block n
            for (int i = 0; i < NIN; ++i) I[B1T / 4 + off(SBO2, i, row)] = W1[j * NIN + i];               // (W1^T)[i][j]
        }
        for (int o = 0; o < 6; ++o) {
            const int row = n * 6 + o;
            for (int k = 0; k < W; ++k) I[B3 / 4 + off(SBO2, row, n * W + k)] = W3[o * W + k];
            I[BIAS3 / 4 + row] = b3[o];
        }
    }
}

static int align4(int v) { return (v + 3) & ~3; }

// host mirror of tcs_ws_layout (mpc_tcsolve.cuh): floats of one CTA's workspace of the tensor-core solve
struct TCSWsHost { size_t total; };
static TCSWsHost tcs_ws_floats(int H, int NU, int RS, int TG) {
    const size_t n = (size_t)H * NU;
    TCSWsHost w;
    w.total = 3 * n * RS + (size_t)SDEMPC_MAX_NU * RS + 16 * (size_t)RS + (size_t)(H + 1) * NX * RS + (size_t)H * 6 * 128 + 16 * 128 +
              (size_t)H * TG * 128 * 4;
    return w;
}

static void build_kparams(sdempc_handle* h) {
    KParams& k = h->kp;
    memset(&k, 0, sizeof k);
    const sdempc_config& c = h->cfg;
    const sdempc_model_header& m = h->mh;
    k.H = c.horizon; k.P = c.num_particles; k.max_iter = c.max_iter; k.max_no_improve = c.max_no_improvement_iter;
    k.maxls = c.maxls; k.reset_option = c.reset_option; k.flags = c.flags;
    for (int t = 0; t < SDEMPC_MAX_H; ++t) { k.dt[t] = c.dt[t]; k.sdt[t] = sqrtf(c.dt[t]); }
    k.discount = c.discount;
    for (int i = 0; i < SDEMPC_MAX_NU; ++i) { k.u_lo[i] = c.u_lo[i]; k.u_hi[i] = c.u_hi[i]; k.uref[i] = c.uref[i]; }
    k.uerr = c.uerr;
    for (int i = 0; i < 3; ++i) { k.perr[i] = c.perr[i]; k.verr[i] = c.verr[i]; k.qerr[i] = c.qerr[i]; k.werr[i] = c.werr[i]; }
    k.res_mult = c.res_mult; k.slew = c.u_slew_coeff;
    k.slewc = c.u_slew_constr_coeff;
    {   // stage weights of the rate constraint: the running float product discount^t times the coefficient
        float disc = 1.0f;
        for (int t = 0; t < SDEMPC_MAX_H; ++t) { k.rate_w[t] = disc * c.u_slew_constr_coeff; disc = disc * c.discount; }
    }
    for (int i = 0; i < SDEMPC_MAX_NU; ++i) { k.slew_lo[i] = c.u_slew_lo[i]; k.slew_hi[i] = c.u_slew_hi[i]; }
    k.init_step = c.init_stepsize; k.max_step = c.max_stepsize; k.coef = c.coef; k.dec_f = c.decrease_factor;
    k.inc_f = c.increase_factor; k.atol = c.atol; k.rtol = c.rtol;
    // derived model constants: one IEEE float operation each (the oracle derives them identically)
    k.inv_m = 1.0f / m.mass; k.grav = m.gravity; k.kT = m.k_thrust; k.kT2 = 2.0f * m.k_thrust;
    for (int i = 0; i < 3; ++i) { k.J[i] = m.inertia[i]; k.Jinv[i] = 1.0f / m.inertia[i]; }
    k.Jd[0] = m.inertia[2] - m.inertia[1]; k.Jd[1] = m.inertia[0] - m.inertia[2]; k.Jd[2] = m.inertia[1] - m.inertia[0];
    for (int r = 0; r < 3; ++r)
        for (int i = 0; i < SDEMPC_MAX_NU; ++i) k.mixer[r][i] = m.mixer[r * SDEMPC_MAX_NU + i];
    for (int i = 0; i < 6; ++i) k.sig0[i] = m.sigma_prior[i];
    // per-warp shared layout
    const int H = c.horizon, NU = c.nu, W = m.width, n = align4(H * NU);
    int o = 0;
    k.o_xk = o; o += n; k.o_yk = o; o += n; k.o_g = o; o += n; k.o_xp = o; o += n; k.o_g2 = o; o += n;
    k.o_uprev = o; o += 8;
    k.o_xref = o; o += (H + 1) * 16;
    k.o_xi = o; o += H * 8;
    k.o_xtape = o; o += (H + 1) * 16;
    k.o_stape = o; o += H * 20;
    k.o_bufA = o; o += 2 * W; k.o_bufB = o; o += 2 * W;
    k.o_act3 = o; o += 2 * W + 8;
    k.o_lz = o; o += 24;
    k.o_red = o; o += 8;
    const bool tape_smem = (W == 32);
    k.o_mtape = o;
    if (tape_smem) o += H * 4 * W;
    k.ws_stride = align4(o) + 4;   // +4 floats: skew consecutive warp regions across banks
    k.team_stride = 32;
    const KernelChoice& kc = h->kc;
    const size_t floats = (size_t)kc.wsmem_floats + 4 + (size_t)kc.G * k.team_stride + (size_t)kc.G * kc.P * k.ws_stride;
    h->smem_bytes = floats * 4;
    h->smem_bytes_spec = ((size_t)kc.wsmem_floats + 4 + (size_t)k.team_stride + (size_t)(SPEC_LSW + SPEC_SGW) * k.ws_stride) * 4;
    h->smem_bytes_cl = ((size_t)kc.wsmem_floats + 4 + (size_t)k.team_stride + (size_t)SPEC_LSW * k.ws_stride) * 4;
    // group kernel: per-problem regions hold only problem data; exchange buffers are per warp; the
    // activation tape lives in global memory (L2 resident)
    KParams& g = h->kp_group;
    g = k;
    int q = 0;
    g.o_xk = q; q += n; g.o_yk = q; q += n; g.o_g = q; q += n; g.o_xp = q; q += n;
    g.o_uprev = q; q += 8;
    g.o_xref = q; q += (H + 1) * 16;
    g.o_xi = q; q += H * 8;
    g.o_xtape = q; q += (H + 1) * 16;
    g.o_stape = q; q += H * 20;
    g.o_lz = q; q += 20;
    g.o_lob = q; q += 12;
    g.o_zb = q; q += 12;
    g.o_bufA = g.o_bufB = g.o_act3 = g.o_red = g.o_mtape = 0;
    g.ws_stride = align4(q) + 4;
    g.gx_stride = 2 * (6 * W + 8);   // two sets: the network phases process two problems per pass
    h->smem_bytes_group = ((size_t)kc.wsmem_floats + 4 + (size_t)GROUP_GW * g.gx_stride + (size_t)GROUP_GW * kc.gp * g.ws_stride) * 4;
}

static int ensure_device(sdempc_handle* h) {
    if (h->dev_ready) {   // several handles on different GPUs may share a thread: always select this handle's device
        CUDA_TRY(cudaSetDevice(h->device));
        return 0;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SDEMPC_ECUDA, "no usable CUDA device (%s); the MPC solve has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (h->device < 0 || h->device >= ndev) return fail(SDEMPC_EINVAL, "device %d out of range (have %d)", h->device, ndev);
    CUDA_TRY(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, h->device));
    if (prop.major != 10)
        return fail(SDEMPC_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", h->device, prop.major, prop.minor);
    h->sm_count = prop.multiProcessorCount;
    if (h->smem_bytes > (size_t)prop.sharedMemPerBlockOptin)
        return fail(SDEMPC_EINVAL, "configuration needs %zu bytes of shared memory per CTA (limit %zu)", h->smem_bytes,
                    (size_t)prop.sharedMemPerBlockOptin);
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&h->ev0));
    CUDA_TRY(cudaEventCreate(&h->ev1));
    CUDA_TRY(cudaMalloc(&h->d_wimg, h->wimg.size() * 4));
    CUDA_TRY(cudaMemcpy(h->d_wimg, h->wimg.data(), h->wimg.size() * 4, cudaMemcpyHostToDevice));
    {   // momentum table with the oracle's float operations, one entry per value of the counter k
        std::vector<float> tab((size_t)h->cfg.max_iter + 2, 0.f);
        for (int k = 1; k < (int)tab.size(); ++k) {
            if (h->cfg.moment_scale == 0.f) tab[k] = (float)k / (float)(k + 3);
            else {
                float b = h->cfg.beta_init;
                for (int i = 1; i < k && b < 1.0f; ++i) { b = b / h->cfg.moment_scale; b = b > 1.0f ? 1.0f : b; }
                tab[k] = b;
            }
        }
        CUDA_TRY(cudaMalloc(&h->d_beta, tab.size() * 4));
        CUDA_TRY(cudaMemcpy(h->d_beta, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
        h->kp.beta_tab = h->d_beta; h->kp_group.beta_tab = h->d_beta;
        h->kp.beta_adaptive = h->kp_group.beta_adaptive = (h->cfg.moment_scale != 0.f) ? 1 : 0;
    }
    if (!h->traj_int.empty()) {
        CUDA_TRY(cudaMalloc(&h->d_traj, h->traj_int.size() * 4));
        CUDA_TRY(cudaMemcpy(h->d_traj, h->traj_int.data(), h->traj_int.size() * 4, cudaMemcpyHostToDevice));
    }
    for (auto fn : {h->kc.solve, h->kc.rollout, h->kc.closed}) {
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    }
    if (h->kc.solve_spec) {
        CUDA_TRY(cudaFuncSetAttribute(h->kc.solve_spec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes_spec));
        CUDA_TRY(cudaFuncSetAttribute(h->kc.closed_spec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes_spec));
    }
    if (h->kc.solve_cl) {
        CUDA_TRY(cudaFuncSetAttribute(h->kc.solve_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes_cl));
        CUDA_TRY(cudaFuncSetAttribute(h->kc.closed_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes_cl));
    }
    if (h->kc.solve_pc)
        CUDA_TRY(cudaFuncSetAttribute(h->kc.solve_pc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes_cl));
    if (h->kc.solve_pcw) {   // wide cluster shape; 16 CTAs are a non-portable size: ask the device how many it holds at once (0: none)
        const int cs = pc_cluster_ctas(h->kc.P, pcw_wpc(h->kc.P));
        h->pcw_clusters = 0;
        bool ok = cudaFuncSetAttribute(h->kc.solve_pcw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes_cl) == cudaSuccess;
        if (ok && cs > 8) ok = cudaFuncSetAttribute(h->kc.solve_pcw, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        if (ok) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs); cfg.blockDim = dim3(pcw_wpc(h->kc.P) * 32); cfg.dynamicSmemBytes = h->smem_bytes_cl;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            int nc = 0;
            if (cudaOccupancyMaxActiveClusters(&nc, h->kc.solve_pcw, &cfg) == cudaSuccess) h->pcw_clusters = nc;
        }
        (void)cudaGetLastError();
        if (const char* e = getenv("SDEMPC_PCW")) { if (atoi(e) == 0) h->pcw_clusters = 0; }   // experiments: never the wide shape
    }
    if (h->kc.solve_group) {
        if (h->smem_bytes_group > (size_t)prop.sharedMemPerBlockOptin) h->kc.solve_group = nullptr;   // does not fit: use one warp per problem
        else CUDA_TRY(cudaFuncSetAttribute(h->kc.solve_group, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes_group));
    }
    if (h->kc.rollout_tc) {
        CUDA_TRY(cudaFuncSetAttribute(h->kc.rollout_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, h->kc.tc_bytes));
        CUDA_TRY(cudaFuncSetAttribute(h->kc.rollout_tc_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, h->kc.tc_bytes_grad));
        CUDA_TRY(cudaFuncSetAttribute(h->kc.solve_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, h->kc.tc_bytes_solve));
        CUDA_TRY(cudaFuncSetAttribute(h->kc.solve_tc_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, h->kc.tc_bytes_solve));
        CUDA_TRY(cudaFuncSetAttribute(h->kc.solve_tc_spec, cudaFuncAttributeMaxDynamicSharedMemorySize, h->kc.tc_bytes_solve));
        CUDA_TRY(cudaFuncSetAttribute(h->kc.solve_tc_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, h->kc.tc_bytes_solve));
        cudaFuncAttributes ft;
        CUDA_TRY(cudaFuncGetAttributes(&ft, h->kc.solve_tc)); h->regs_tc = ft.numRegs;
        CUDA_TRY(cudaFuncGetAttributes(&ft, h->kc.solve_tc_lat)); h->regs_tc_lat = ft.numRegs;
        CUDA_TRY(cudaFuncGetAttributes(&ft, h->kc.solve_tc_spec)); h->regs_tc_spec = ft.numRegs;
        CUDA_TRY(cudaFuncGetAttributes(&ft, h->kc.solve_tc_rate)); h->regs_tc_rate = ft.numRegs;
        CUDA_TRY(cudaMalloc(&h->d_wimg_tc, h->wimg_tc.size() * 4));
        CUDA_TRY(cudaMemcpy(h->d_wimg_tc, h->wimg_tc.data(), h->wimg_tc.size() * 4, cudaMemcpyHostToDevice));
    }
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, h->kc.solve));
    h->regs = fa.numRegs;
    h->dev_ready = true;
    return 0;
}

static int grid_for(const sdempc_handle* h, int B) {
    return std::max(1, std::min(B, h->sm_count));   // one persistent CTA per SM at most; problems dealt across CTAs first
}

static bool use_spec(const sdempc_handle* h, int B) {
    // latency regime: at most one problem per SM, or forced by the flag; bit-identical to the batched kernel
    if (h->kc.solve_spec == nullptr || h->cfg.maxls < 1 || (h->cfg.flags & SDEMPC_F_SEQUENTIAL_LS)) return false;
    if (h->cfg.flags & SDEMPC_F_SPECULATIVE_LS) return true;
    return !(h->cfg.flags & SDEMPC_F_GROUP) && B <= h->sm_count;
}

// cluster variant of the latency kernel: two SMs per problem
static bool use_cluster(const sdempc_handle* h, int B) {
    return use_spec(h, B) && h->kc.solve_cl != nullptr && 2 * B <= h->sm_count && !(h->cfg.flags & SDEMPC_F_NO_CLUSTER);
}

// cluster latency kernel (mpc_pcluster.cuh), compact shape: P > 1, the whole batch resident at once
static bool use_pcluster(const sdempc_handle* h, int B) {
    if (h->kc.solve_pc == nullptr || h->cfg.maxls < 1) return false;
    if (h->cfg.flags & (SDEMPC_F_SEQUENTIAL_LS | SDEMPC_F_NO_CLUSTER)) return false;
    return B * pc_cluster_ctas(h->kc.P, PC_WPC) <= h->sm_count;
}
// ... wide shape (two warps per SM), while the device holds all of the batch's clusters at once
static bool use_pcluster_wide(const sdempc_handle* h, int B) {
    if (h->kc.solve_pcw == nullptr || h->cfg.maxls < 1 || B > h->pcw_clusters) return false;
    if (h->cfg.flags & (SDEMPC_F_SEQUENTIAL_LS | SDEMPC_F_NO_CLUSTER)) return false;
    if (h->kc.P == 1 && !use_spec(h, B)) return false;
    return B * pc_cluster_ctas(h->kc.P, pcw_wpc(h->kc.P)) <= h->sm_count;
}

static bool use_group(const sdempc_handle* h, int B) {
    // measured crossover against one warp per problem (tools/batch_sweep.py): ~13 problems per SM with 4 problems
    // per warp (iris), ~10 with 2 (width 64); with 3 (nu = 6, width 32) it never wins at these sizes: flag only
    if (h->kc.solve_group == nullptr || (h->cfg.flags & SDEMPC_F_SEQUENTIAL_LS) || use_spec(h, B)) return false;
    if (h->cfg.flags & SDEMPC_F_GROUP) return true;
    const int per_sm = h->kc.gp == 4 ? 13 : h->kc.gp == 2 ? 10 : (1 << 20);
    return B > per_sm * h->sm_count;
}

// Tensor-core solve (SDEMPC_F_TENSOR): problems per CTA and which build of the kernel.  A CTA has NT = 128 / P rollout slots.
// Measured (iris, 200 iterations): a CTA alone on its SM is latency bound (2.5 us per step evaluation), two CTAs on an SM slow
// each other 1.4x, four 1.8x; NT / 4 problems per CTA need two line-search passes per iteration at most (one in 97 % of the
// iterations), NT problems about three plus a full-width gradient pass.  The build with the register budget of two CTAs per SM
// (164 registers, no spill) is 11 % faster per CTA than the four-CTA build (128 registers): 4096 problems 35.8 against 40.1 ms,
// 8192: 40.4 / 45.8, 16 384: 53.8 / 59.0, 32 768: 74.9 / 79.5, but 65 536 (two waves at two CTAs per SM): 139 / 111.
// So: one CTA per SM with up to NT / 4 problems; then two CTAs per SM, growing the CTAs up to NT problems (two-CTA build);
// beyond 2 x SMs x NT problems the four-CTA build with the SMs filled at NT / 4 problems per CTA first.
static int tcs_problems_per_cta(const sdempc_handle* h, int B, bool* lat_build) {
    const int P = h->cfg.num_particles, NT = 128 / P;
    const int max_res = 512 / std::max(1, h->kc.tc_cols);           // resident CTAs per SM tensor memory allows (4 or 2)
    const int base = std::max(1, NT / 4);
    *lat_build = true;
    if (h->tcs_ppc_override > 0) {
        const int ppc = std::min(h->tcs_ppc_override, NT);
        *lat_build = (B + ppc - 1) / ppc <= 2 * h->sm_count;
        return ppc;
    }
    int ppc = (B + h->sm_count - 1) / h->sm_count;                  // one CTA per SM
    if (ppc <= base) return std::max(1, ppc);
    ppc = std::max(base, (B + 2 * h->sm_count - 1) / (2 * h->sm_count));   // two CTAs per SM, growing
    if (ppc <= NT) return ppc;
    *lat_build = false;                                              // large batch: all the CTAs tensor memory allows
    const int resident = h->sm_count * max_res;
    return std::max(1, std::min(std::max(base, (B + resident - 1) / resident), NT));
}

static int ensure_tcs_ws(sdempc_handle* h, int grid, int rs) {
    const TCSWsHost w = tcs_ws_floats(h->cfg.horizon, h->cfg.nu, rs, h->kc.tc_solve_tape_granules);
    const size_t need = (size_t)grid * w.total * 4;
    if (need <= h->tcs_ws_bytes) return 0;
    if (h->d_tcs_ws) cudaFree(h->d_tcs_ws);
    h->d_tcs_ws = nullptr; h->tcs_ws_bytes = 0;
    CUDA_TRY(cudaMalloc(&h->d_tcs_ws, need));
    h->tcs_ws_bytes = need;
    return 0;
}

static int ensure_mtape_group(sdempc_handle* h, int grid) {
    const size_t tapes = (size_t)grid * GROUP_GW * h->kc.gp;
    if (tapes <= h->mtape_group_n) return 0;
    if (h->d_mtape_group) cudaFree(h->d_mtape_group);
    h->d_mtape_group = nullptr;
    CUDA_TRY(cudaMalloc(&h->d_mtape_group, tapes * (size_t)h->cfg.horizon * 2 * h->mh.width * sizeof(float2)));
    h->mtape_group_n = tapes;
    return 0;
}

static int ensure_mtape(sdempc_handle* h, int grid) {
    if (h->mh.width == 32) return 0;
    const size_t warps = (size_t)grid * std::max(h->kc.G * h->kc.P, SPEC_LSW + SPEC_SGW);
    if (warps <= h->mtape_warps) return 0;
    if (h->d_mtape) cudaFree(h->d_mtape);
    h->d_mtape = nullptr;
    CUDA_TRY(cudaMalloc(&h->d_mtape, warps * (size_t)h->cfg.horizon * 2 * h->mh.width * sizeof(float2)));
    h->mtape_warps = warps;
    return 0;
}

static int ensure_io(sdempc_handle* h, size_t in_bytes, size_t out_bytes) {
    if (in_bytes > h->in_cap) {
        if (h->d_in) cudaFree(h->d_in);
        if (h->h_in) cudaFreeHost(h->h_in);
        h->d_in = nullptr; h->h_in = nullptr; h->in_cap = 0;
        const size_t cap = in_bytes + in_bytes / 4 + 4096;
        CUDA_TRY(cudaMalloc(&h->d_in, cap));
        CUDA_TRY(cudaMallocHost(&h->h_in, cap));
        h->in_cap = cap;
    }
    if (out_bytes > h->out_cap) {
        if (h->d_out) cudaFree(h->d_out);
        if (h->h_out) cudaFreeHost(h->h_out);
        h->d_out = nullptr; h->h_out = nullptr; h->out_cap = 0;
        const size_t cap = out_bytes + out_bytes / 4 + 4096;
        CUDA_TRY(cudaMalloc(&h->d_out, cap));
        CUDA_TRY(cudaMallocHost(&h->h_out, cap));
        h->out_cap = cap;
    }
    return 0;
}

// is p page-locked host memory (cudaHostAlloc / cudaHostRegister, e.g. through sdempc_host_register)?
static bool is_pinned(const void* p) {
    if (p == nullptr) return true;
    cudaPointerAttributes at;
    const bool ok = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost;
    (void)cudaGetLastError();
    return ok;
}

// Sequential packer of 16-byte aligned sub-buffers into the device IN block.
// Inputs of a call are packed into one pinned staging block and sent with one copy — or, when the caller passes page-locked
// arrays (direct != nullptr), every array goes with its own asynchronous copy: straight from the caller's memory if it is
// page-locked, from its slot of the staging block otherwise.
struct Packer {
    char* hbase; char* dbase; size_t off = 0;
    cudaStream_t direct = nullptr;
    bool failed = false, all_direct = true;
    template <typename T>
    const T* put(const T* src, size_t count) {
        if (src == nullptr) return nullptr;
        const size_t bytes = count * sizeof(T);
        if (direct) {
            const void* from = src;
            if (!is_pinned(src)) { memcpy(hbase + off, src, bytes); from = hbase + off; all_direct = false; }
            failed |= cudaMemcpyAsync(dbase + off, from, bytes, cudaMemcpyHostToDevice, direct) != cudaSuccess;
        } else if (hbase) memcpy(hbase + off, src, bytes);
        const T* d = reinterpret_cast<const T*>(dbase + off);
        off += (bytes + 15) & ~(size_t)15;
        return d;
    }
    template <typename T>
    T* reserve(size_t count) {
        T* d = reinterpret_cast<T*>(dbase + off);
        off += (count * sizeof(T) + 15) & ~(size_t)15;
        return d;
    }
};

static size_t a16(size_t b) { return (b + 15) & ~(size_t)15; }

static int check_solve_args(const sdempc_handle* h, const sdempc_solve_args* a) {
    if (!a || a->B < 1) return fail(SDEMPC_EINVAL, "B must be >= 1");
    if (!a->x || !a->u_plan || !a->x_evol || !a->info) return fail(SDEMPC_EINVAL, "x, u_plan, x_evol and info are required");
    if (!a->xref_win && !a->curr_t && !a->xdes) return fail(SDEMPC_EINVAL, "one of xref_win, curr_t, xdes is required");
    if (!a->xref_win && a->curr_t && h->traj_int.empty())
        return fail(SDEMPC_ESTATE, "trajectory mode (curr_t) needs sdempc_set_trajectory() first");
    if (!a->rng && !a->xi_override) return fail(SDEMPC_EINVAL, "rng is required unless xi_override is given");
    return 0;
}

// Stage inputs of a solve into pinned memory + async H2D; fills h->staged.
static int stage_solve(sdempc_handle* h, const sdempc_solve_args* a) {
    int rc = check_solve_args(h, a);
    if (rc) return rc;
    if ((rc = ensure_device(h))) return rc;
    const int B = a->B, H = h->cfg.horizon, NU = h->cfg.nu, P = h->cfg.num_particles, n = H * NU;
    const bool use_win = a->xref_win != nullptr, use_t = !use_win && a->curr_t != nullptr;
    size_t in_bytes = a16((size_t)B * NX * 4) + a16((size_t)B * n * 4) + a16((size_t)B * sizeof(sdempc_info)) + 64;
    if (use_win) in_bytes += a16((size_t)B * (H + 1) * NX * 4);
    else if (use_t) in_bytes += a16((size_t)B * 4);
    else in_bytes += a16((size_t)B * NX * 4);
    if (a->xi_override) in_bytes += a16((size_t)B * P * H * 6 * 4);
    else in_bytes += a16((size_t)B * 16);
    const size_t out_bytes = a16((size_t)B * (H + 1) * NX * 4) + a16((size_t)B * n * 4) + a16((size_t)B * sizeof(sdempc_info)) + 64;
    if ((rc = ensure_io(h, in_bytes, out_bytes))) return rc;
    // SDEMPC_F_TENSOR: the batched solve on the tensor-core mapping (explicit opt-in: TF32 products, not SPEC-ARITH)
    const bool tcs = (h->cfg.flags & SDEMPC_F_TENSOR) != 0;
    if (tcs && (!h->kc.solve_tc || P > 32 || (P & (P - 1)) != 0))
        return fail(SDEMPC_EINVAL, "SDEMPC_F_TENSOR: the tensor-core solve supports 1, 2, 4, ... 32 particles");
    const bool pclw = !tcs && use_pcluster_wide(h, B), pcl = !tcs && (pclw || use_pcluster(h, B));
    const bool spec = !tcs && !pcl && use_spec(h, B), group = !tcs && use_group(h, B), cl = !tcs && !pcl && use_cluster(h, B);
    // throughput kernel: enough CTAs that a warp holds ~2+ problems when the batch is small, all SMs otherwise
    bool tcs_lat = false;
    const int ppc = tcs ? tcs_problems_per_cta(h, B, &tcs_lat) : 0, rs = 128;   // TCS_RS (mpc_tcsolve.cuh)
    const int grid = tcs ? (B + ppc - 1) / ppc : pclw ? B * pc_cluster_ctas(h->kc.P, pcw_wpc(h->kc.P)) : pcl ? B * pc_cluster_ctas(h->kc.P, PC_WPC) : cl ? 2 * B : spec ? std::min(B, h->sm_count)
                          : group ? std::max(1, std::min((B + 2 * GROUP_GW - 1) / (2 * GROUP_GW), h->sm_count)) : grid_for(h, B);
    if (tcs) { if ((rc = ensure_tcs_ws(h, grid, rs))) return rc; }
    else if (group) { if ((rc = ensure_mtape_group(h, grid))) return rc; }
    else if ((rc = ensure_mtape(h, grid))) return rc;
    if (a->trace) {
        const size_t tb = (size_t)B * h->cfg.max_iter * SDEMPC_TRACE_W * 4;
        if (tb > h->trace_cap) {
            if (h->d_trace) cudaFree(h->d_trace);
            h->d_trace = nullptr;
            CUDA_TRY(cudaMalloc(&h->d_trace, tb));
            h->trace_cap = tb;
        }
        CUDA_TRY(cudaMemsetAsync(h->d_trace, 0, tb, h->stream));
    }
    KParams k = group ? h->kp_group : h->kp;
    Packer pk{h->h_in, h->d_in};
    // zero-copy staging: the large inputs page-locked by the caller (sdempc_host_register) go by DMA straight from its memory
    const bool any_pinned = (a->x && is_pinned(a->x)) || (use_win && is_pinned(a->xref_win)) || (a->xi_override && is_pinned(a->xi_override));
    if (any_pinned) pk.direct = h->stream;
    k.B = B;
    k.x = pk.put(a->x, (size_t)B * NX);
    k.u_plan = const_cast<float*>(pk.put(a->u_plan, (size_t)B * n));
    k.info = const_cast<sdempc_info*>(pk.put(a->info, (size_t)B));
    k.xref_win = use_win ? pk.put(a->xref_win, (size_t)B * (H + 1) * NX) : nullptr;
    k.curr_t = use_t ? pk.put(a->curr_t, (size_t)B) : nullptr;
    k.xdes = (!use_win && !use_t) ? pk.put(a->xdes, (size_t)B * NX) : nullptr;
    k.xi_override = a->xi_override ? pk.put(a->xi_override, (size_t)B * P * H * 6) : nullptr;
    k.rng = a->xi_override ? nullptr : reinterpret_cast<const unsigned long long*>(pk.put(a->rng, (size_t)B * 2));
    if (pk.failed) { CUDA_TRY(cudaGetLastError()); return fail(SDEMPC_ECUDA, "copy from page-locked caller memory failed"); }
    h->staged_direct_in = any_pinned;
    if (!any_pinned) CUDA_TRY(cudaMemcpyAsync(h->d_in, h->h_in, pk.off, cudaMemcpyHostToDevice, h->stream));
    Packer po{nullptr, h->d_out};
    k.x_evol = po.reserve<float>((size_t)B * (H + 1) * NX);
    k.u_plan_out = po.reserve<float>((size_t)B * n);
    k.info_out = po.reserve<sdempc_info>((size_t)B);
    k.trace = a->trace ? h->d_trace : nullptr;
    k.wimg = h->d_wimg; k.traj = h->d_traj; k.T = h->T; k.mtape_g = group ? h->d_mtape_group : h->d_mtape;
    if (tcs) { k.wimg = h->d_wimg_tc; k.tcs_ws = h->d_tcs_ws; k.tcs_ppc = ppc; k.tcs_rs = rs; k.tcs_sms = h->sm_count; }
    h->staged = k; h->staged_B = B; h->staged_ok = true; h->last_grid = grid; h->staged_spec = spec; h->staged_group = group; h->staged_cl = cl; h->staged_pc = pcl; h->staged_pcw = pclw;
    h->staged_tc = tcs; h->staged_tc_lat = tcs && tcs_lat;
    h->staged_tc_rate = tcs && h->cfg.u_slew_constr_coeff != 0.0f;   // the one build that evaluates the input-rate constraint
    h->staged_tc_spec = tcs && !h->staged_tc_rate && tcs_lat && h->tcs_spec && ppc * (h->tcs_spec == 2 ? 8 : 4) <= 128 / P;   // (few problems per CTA: speculative gradient passes, mpc_tcsolve.cuh)
    return 0;
}

static int launch(sdempc_handle* h, void (*fn)(KParams), const KParams& k, int grid) {
    if (fn != nullptr && (fn == h->kc.solve_cl || fn == h->kc.closed_cl || fn == h->kc.solve_pc || fn == h->kc.solve_pcw)) {   // one cluster per problem
        const unsigned cs = (fn == h->kc.solve_pcw) ? (unsigned)pc_cluster_ctas(h->kc.P, pcw_wpc(h->kc.P)) : (fn == h->kc.solve_pc) ? (unsigned)pc_cluster_ctas(h->kc.P, PC_WPC) : 2u;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3((fn == h->kc.solve_pcw ? pcw_wpc(h->kc.P) : SPEC_LSW) * 32);
        cfg.dynamicSmemBytes = h->smem_bytes_cl;
        cfg.stream = h->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, fn, k));
        h->launches += 1;
        return 0;
    }
    const bool spec = (fn == h->kc.solve_spec || fn == h->kc.closed_spec) && fn != nullptr;
    const bool group = (fn == h->kc.solve_group) && fn != nullptr;
    const bool tcs = (fn == h->kc.solve_tc || fn == h->kc.solve_tc_lat || fn == h->kc.solve_tc_spec || fn == h->kc.solve_tc_rate) && fn != nullptr;
    const int threads = tcs ? 128 : spec ? (SPEC_LSW + SPEC_SGW) * 32 : group ? GROUP_GW * 32 : h->kc.G * h->kc.P * 32;
    const size_t smem = tcs ? (size_t)h->kc.tc_bytes_solve : spec ? h->smem_bytes_spec : group ? h->smem_bytes_group : h->smem_bytes;
    void* args[] = {const_cast<KParams*>(&k)};
    CUDA_TRY(cudaLaunchKernel(reinterpret_cast<const void*>(fn), dim3(grid), dim3(threads), args, smem, h->stream));
    h->launches += 1;
    return 0;
}

static void (*staged_kernel(const sdempc_handle* h))(KParams) {
    return h->staged_tc ? (h->staged_tc_rate ? h->kc.solve_tc_rate : h->staged_tc_spec ? h->kc.solve_tc_spec : h->staged_tc_lat ? h->kc.solve_tc_lat : h->kc.solve_tc) : h->staged_pcw ? h->kc.solve_pcw : h->staged_pc ? h->kc.solve_pc : h->staged_cl ? h->kc.solve_cl : h->staged_spec ? h->kc.solve_spec
           : h->staged_group ? h->kc.solve_group : h->kc.solve;
}

static int fetch_solve(sdempc_handle* h, const sdempc_solve_args* a, bool direct = false) {
    if (!h->staged_ok || a->B != h->staged_B) return fail(SDEMPC_ESTATE, "fetch without a matching stage");
    const int B = a->B, H = h->cfg.horizon, NU = h->cfg.nu, n = H * NU;
    const size_t o_x = 0, o_p = a16((size_t)B * (H + 1) * NX * 4), o_i = o_p + a16((size_t)B * n * 4);
    const size_t total = o_i + a16((size_t)B * sizeof(sdempc_info));
    if (direct) {   // the caller's buffers are page-locked (cudaHostRegister): copy straight into them, no staging copy
        CUDA_TRY(cudaMemcpyAsync(a->x_evol, h->d_out + o_x, (size_t)B * (H + 1) * NX * 4, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaMemcpyAsync(a->u_plan, h->d_out + o_p, (size_t)B * n * 4, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaMemcpyAsync(a->info, h->d_out + o_i, (size_t)B * sizeof(sdempc_info), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        for (int b = 0; b < B; ++b) a->info[b].solve_time_us = h->last_ms * 1000.f;
        return 0;
    }
    CUDA_TRY(cudaMemcpyAsync(h->h_out, h->d_out, total, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    memcpy(a->x_evol, h->h_out + o_x, (size_t)B * (H + 1) * NX * 4);
    memcpy(a->u_plan, h->h_out + o_p, (size_t)B * n * 4);
    memcpy(a->info, h->h_out + o_i, (size_t)B * sizeof(sdempc_info));
    for (int b = 0; b < B; ++b) a->info[b].solve_time_us = h->last_ms * 1000.f;
    if (a->trace) {
        CUDA_TRY(cudaMemcpy(a->trace, h->d_trace, (size_t)B * h->cfg.max_iter * SDEMPC_TRACE_W * 4, cudaMemcpyDeviceToHost));
    }
    return 0;
}

extern "C" {

const char* sdempc_last_error(void) { return g_err.c_str(); }
const char* sdempc_version(void) { return "sdempc 0.1 (sm_100a, SPEC-ARITH 1)"; }

int sdempc_create(const sdempc_config* cfg, const void* model_blob, size_t nbytes, int device, sdempc_t** out) {
    if (!cfg || !model_blob || !out) return fail(SDEMPC_EINVAL, "null argument");
    if (nbytes < sizeof(sdempc_model_header)) return fail(SDEMPC_EINVAL, "model blob too small");
    sdempc_model_header mh;
    memcpy(&mh, model_blob, sizeof mh);
    if (mh.magic != SDEMPC_MODEL_MAGIC || mh.version != SDEMPC_MODEL_VERSION) return fail(SDEMPC_EINVAL, "bad model magic/version");
    if (mh.n_hidden != 2 || mh.n_out != 6 || mh.n_in != 6 + mh.nu) return fail(SDEMPC_EINVAL, "unsupported network shape");
    if (cfg->nu != mh.nu) return fail(SDEMPC_EINVAL, "config nu=%d does not match model nu=%d", cfg->nu, mh.nu);
    // the kernels build the reference window and write x_evol with one lane per row (rows 0..H): H + 1 <= 32
    if (cfg->horizon < 1 || cfg->horizon > SDEMPC_MAX_H - 1) return fail(SDEMPC_EINVAL, "horizon out of range 1..%d", SDEMPC_MAX_H - 1);
    if (cfg->max_iter < 1 || cfg->maxls < 0) return fail(SDEMPC_EINVAL, "max_iter must be >= 1 and maxls >= 0");
    if (!(cfg->moment_scale >= 0.f && cfg->moment_scale <= 1.f) || (cfg->moment_scale > 0.f && !(cfg->beta_init >= 0.f && cfg->beta_init <= 1.f)))
        return fail(SDEMPC_EINVAL, "moment_scale must be 0 (classical momentum) or in (0, 1], with beta_init in [0, 1]");
    const int W = mh.width, NIN = mh.n_in;
    const size_t per_net = (size_t)W * NIN + W + (size_t)W * W + W + 6 * (size_t)W + 6;
    if (nbytes < sizeof mh + 2 * per_net * 4) return fail(SDEMPC_EINVAL, "model blob truncated");
    const KernelChoice* kc = nullptr;
    for (const auto& c : choices())
        if (c.nu == mh.nu && c.W == W && c.P == cfg->num_particles) kc = &c;
    if (!kc)
        return fail(SDEMPC_EINVAL, "no compiled kernel for nu=%d width=%d particles=%d (see choices() in sdempc_api.cu)", mh.nu, W,
                    cfg->num_particles);
    sdempc_handle* h = new sdempc_handle();
    h->cfg = *cfg; h->mh = mh; h->device = device; h->kc = *kc;
    h->weights.resize(2 * per_net);
    memcpy(h->weights.data(), (const char*)model_blob + sizeof mh, 2 * per_net * 4);
    pack_weights(h);
    if (h->kc.rollout_tc) pack_weights_tc(h);
    if (const char* e = getenv("SDEMPC_TC_SPEC")) h->tcs_spec = atoi(e);
    if (const char* e = getenv("SDEMPC_TC_PPC")) h->tcs_ppc_override = atoi(e);   // experiments: problems per CTA of the tensor-core solve
    build_kparams(h);
    *out = h;
    return 0;
}

void sdempc_destroy(sdempc_t* h) {
    if (!h) return;
    if (h->dev_ready) {
        cudaSetDevice(h->device);
        if (h->stream) cudaStreamSynchronize(h->stream);
        cudaFree(h->d_wimg); cudaFree(h->d_beta); cudaFree(h->d_traj); cudaFree(h->d_mtape); cudaFree(h->d_mtape_group); cudaFree(h->d_in); cudaFree(h->d_out);
        cudaFree(h->d_trace); cudaFree(h->d_flush); cudaFree(h->d_wimg_tc); cudaFree(h->d_tape_tc); cudaFree(h->d_tcs_ws);
        if (h->h_in) cudaFreeHost(h->h_in);
        if (h->h_out) cudaFreeHost(h->h_out);
        if (h->ev0) cudaEventDestroy(h->ev0);
        if (h->ev1) cudaEventDestroy(h->ev1);
        if (h->stream) cudaStreamDestroy(h->stream);
    }
    delete h;
}

int sdempc_set_trajectory(sdempc_t* h, const float* table, int T) {
    if (!h || !table || T < 2) return fail(SDEMPC_EINVAL, "trajectory needs at least 2 rows");
    std::vector<float> ext(table, table + (size_t)T * 14), in((size_t)T * 14);
    const bool enu = (h->cfg.flags & SDEMPC_F_FRAME_ENU) != 0;
    for (int r = 0; r < T; ++r) {
        const float* a = table + (size_t)r * 14;
        float* b = in.data() + (size_t)r * 14;
        if (r > 0 && !(a[0] > a[-14])) return fail(SDEMPC_EINVAL, "trajectory times must be strictly increasing (row %d)", r);
        b[0] = a[0];
        if (enu) {
            const float s = 0.70710678118654752440f;
            float c0 = s * (a[7] + a[10]), c1 = s * (a[8] + a[9]), c2 = s * (a[8] - a[9]), c3 = s * (a[7] - a[10]);
            if (c0 < 0) { c0 = -c0; c1 = -c1; c2 = -c2; c3 = -c3; }
            b[1] = a[2]; b[2] = a[1]; b[3] = -a[3];
            b[4] = a[5]; b[5] = a[4]; b[6] = -a[6];
            b[7] = c0; b[8] = c1; b[9] = c2; b[10] = c3;
            b[11] = a[11]; b[12] = -a[12]; b[13] = -a[13];
        } else {
            for (int i = 1; i < 14; ++i) b[i] = a[i];
        }
        if (r > 0) {
            float d = 0;
            for (int i = 7; i < 11; ++i) d += b[i] * b[i - 14];
            if (d < 0) for (int i = 7; i < 11; ++i) b[i] = -b[i];
        }
    }
    h->traj_ext.swap(ext); h->traj_int.swap(in); h->T = T;
    if (h->dev_ready) {
        CUDA_TRY(cudaSetDevice(h->device));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        if (h->d_traj) cudaFree(h->d_traj);
        h->d_traj = nullptr;
        CUDA_TRY(cudaMalloc(&h->d_traj, h->traj_int.size() * 4));
        CUDA_TRY(cudaMemcpy(h->d_traj, h->traj_int.data(), h->traj_int.size() * 4, cudaMemcpyHostToDevice));
    }
    return 0;
}

int sdempc_state_from_traj(const sdempc_t* h, const float* t, int n, float* out) {
    if (!h || !t || !out) return fail(SDEMPC_EINVAL, "null argument");
    if (h->traj_ext.empty()) return fail(SDEMPC_ESTATE, "no trajectory set");
    const float* tab = h->traj_ext.data();
    const int T = h->T;
    for (int q = 0; q < n; ++q) {
        float* o = out + (size_t)q * NX;
        const float tt = t[q];
        int lo = 0, hi = T - 1;
        if (tt <= tab[0]) { for (int i = 0; i < NX; ++i) o[i] = tab[1 + i]; }
        else if (tt >= tab[(size_t)hi * 14]) { for (int i = 0; i < NX; ++i) o[i] = tab[(size_t)hi * 14 + 1 + i]; }
        else {
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tab[(size_t)mid * 14] <= tt) lo = mid; else hi = mid; }
            const float *a = tab + (size_t)lo * 14, *b = a + 14;
            const float al = (tt - a[0]) / (b[0] - a[0]);
            for (int i = 0; i < NX; ++i) o[i] = fmaf(al, b[1 + i] - a[1 + i], a[1 + i]);
        }
        const float n2 = fmaf(o[9], o[9], fmaf(o[8], o[8], fmaf(o[7], o[7], o[6] * o[6])));
        const float inv = 1.0f / sqrtf(n2);
        for (int i = 6; i < 10; ++i) o[i] = o[i] * inv;
    }
    return 0;
}

int sdempc_reset(sdempc_t* h, int B, const float* x, const float* xdes, float* u_plan, sdempc_info* info) {
    (void)x; (void)xdes;
    if (!h || B < 1 || !u_plan || !info) return fail(SDEMPC_EINVAL, "bad argument");
    const sdempc_config& c = h->cfg;
    for (int b = 0; b < B; ++b) {
        for (int t = 0; t < c.horizon; ++t)
            for (int i = 0; i < c.nu; ++i) {
                float v = c.uref[i];
                v = v < c.u_lo[i] ? c.u_lo[i] : v;
                v = v > c.u_hi[i] ? c.u_hi[i] : v;
                u_plan[((size_t)b * c.horizon + t) * c.nu + i] = v;
            }
        memset(&info[b], 0, sizeof(sdempc_info));
        info[b].stepsize = c.init_stepsize;
    }
    return 0;
}

int sdempc_stage(sdempc_t* h, const sdempc_solve_args* args) {
    if (!h) return fail(SDEMPC_EINVAL, "null handle");
    const int rc = stage_solve(h, args);
    // page-locked inputs are read by asynchronous copies: the caller may reuse them as soon as this returns
    if (rc == 0 && h->staged_direct_in) CUDA_TRY(cudaStreamSynchronize(h->stream));
    return rc;
}

int sdempc_launch_timed(sdempc_t* h, int n, int flush_l2, float* ms) {
    if (!h || !h->staged_ok) return fail(SDEMPC_ESTATE, "sdempc_stage() first");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t FL = (size_t)256 << 20;
    if (flush_l2 && !h->d_flush) CUDA_TRY(cudaMalloc(&h->d_flush, FL));
    for (int i = 0; i < n; ++i) {
        if (flush_l2) CUDA_TRY(cudaMemsetAsync(h->d_flush, i & 0xff, FL, h->stream));
        CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
        int rc = launch(h, staged_kernel(h), h->staged, h->last_grid);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
        CUDA_TRY(cudaEventSynchronize(h->ev1));
        float t = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&t, h->ev0, h->ev1));
        h->last_ms = t;
        if (ms) ms[i] = t;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int sdempc_sync(sdempc_t* h) {
    if (!h || !h->dev_ready) return fail(SDEMPC_ESTATE, "no device state yet");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int sdempc_device_out(sdempc_t* h, void** dev_ptr, size_t* nbytes) {
    if (!h || !dev_ptr || !nbytes) return fail(SDEMPC_EINVAL, "null argument");
    if (!h->staged_ok) return fail(SDEMPC_ESTATE, "sdempc_stage() first");
    const int B = h->staged_B, H = h->cfg.horizon, NU = h->cfg.nu;
    *dev_ptr = h->d_out;
    *nbytes = a16((size_t)B * (H + 1) * NX * 4) + a16((size_t)B * H * NU * 4) + a16((size_t)B * sizeof(sdempc_info));
    return 0;
}

int sdempc_fetch(sdempc_t* h, const sdempc_solve_args* args) {
    if (!h || !args) return fail(SDEMPC_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    return fetch_solve(h, args);
}

int sdempc_host_register(void* p, size_t nbytes) {
    if (!p || nbytes == 0) return fail(SDEMPC_EINVAL, "null range");
    const cudaError_t e = cudaHostRegister(p, nbytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();   // leave no sticky error behind: the caller falls back to staged copies
        return fail(SDEMPC_ECUDA, "cudaHostRegister: %s", cudaGetErrorString(e));
    }
    return 0;
}

int sdempc_host_unregister(void* p) {
    if (!p) return fail(SDEMPC_EINVAL, "null range");
    const cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(SDEMPC_ECUDA, "cudaHostUnregister: %s", cudaGetErrorString(e));
    }
    return 0;
}

int sdempc_fetch_direct(sdempc_t* h, const sdempc_solve_args* args) {
    if (!h || !args) return fail(SDEMPC_EINVAL, "null argument");
    if (!args->x_evol || !args->u_plan || !args->info) return fail(SDEMPC_EINVAL, "x_evol, u_plan and info are required");
    CUDA_TRY(cudaSetDevice(h->device));
    return fetch_solve(h, args, true);
}

int sdempc_solve_ex(sdempc_t* h, const sdempc_solve_args* a) {
    if (!h) return fail(SDEMPC_EINVAL, "null handle");
    int rc = stage_solve(h, a);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    if ((rc = launch(h, staged_kernel(h), h->staged, h->last_grid))) return rc;
    CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    // results straight into the caller's memory when it is page-locked (and no trace is asked for), through the staging block otherwise
    const bool direct_out = !a->trace && is_pinned(a->x_evol) && is_pinned(a->u_plan) && is_pinned(a->info);
    if ((rc = fetch_solve(h, a, direct_out))) return rc;   // synchronises the stream
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, h->ev0, h->ev1));
    h->last_ms = t;
    for (int b = 0; b < a->B; ++b) a->info[b].solve_time_us = t * 1000.f;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int sdempc_solve(sdempc_t* h, int B, const float* x, const float* curr_t, const float* xdes, const uint64_t* rng, float* u_plan,
                 float* x_evol, sdempc_info* info, const float* xi_override) {
    sdempc_solve_args a;
    memset(&a, 0, sizeof a);
    a.B = B; a.x = x; a.curr_t = curr_t; a.xdes = xdes; a.rng = rng; a.u_plan = u_plan; a.x_evol = x_evol; a.info = info;
    a.xi_override = xi_override;
    return sdempc_solve_ex(h, &a);
}

int sdempc_rollout(sdempc_t* h, int B, const float* x, const float* curr_t, const float* xdes, const float* xref_win,
                   const uint64_t* rng, const float* xi_override, const float* u, const float* u_prev, float* cost, float* grad,
                   float* x_evol) {
    if (!h || B < 1 || !x || !u || !u_prev || !cost) return fail(SDEMPC_EINVAL, "bad argument");
    if (!xref_win && !curr_t && !xdes) return fail(SDEMPC_EINVAL, "one of xref_win, curr_t, xdes is required");
    if (!xref_win && curr_t && h->traj_int.empty()) return fail(SDEMPC_ESTATE, "trajectory mode needs sdempc_set_trajectory() first");
    if (!rng && !xi_override) return fail(SDEMPC_EINVAL, "rng is required unless xi_override is given");
    int rc = ensure_device(h);
    if (rc) return rc;
    const int H = h->cfg.horizon, NU = h->cfg.nu, P = h->cfg.num_particles, n = H * NU;
    const bool use_win = xref_win != nullptr, use_t = !use_win && curr_t != nullptr;
    size_t in_bytes = a16((size_t)B * NX * 4) + a16((size_t)B * n * 4) + a16((size_t)B * NU * 4) + 64 +
                      a16((size_t)B * (H + 1) * NX * 4) + a16((size_t)B * P * H * 6 * 4 + (size_t)B * 16);
    const size_t out_bytes = a16((size_t)B * (H + 1) * NX * 4) + a16((size_t)B * n * 4) + a16((size_t)B * 4) + 64;
    if ((rc = ensure_io(h, in_bytes, out_bytes))) return rc;
    const int grid = grid_for(h, B);
    if ((rc = ensure_mtape(h, grid))) return rc;
    KParams k = h->kp;
    Packer pk{h->h_in, h->d_in};
    k.B = B;
    k.x = pk.put(x, (size_t)B * NX);
    k.u_in = pk.put(u, (size_t)B * n);
    k.uprev_in = pk.put(u_prev, (size_t)B * NU);
    k.xref_win = use_win ? pk.put(xref_win, (size_t)B * (H + 1) * NX) : nullptr;
    k.curr_t = use_t ? pk.put(curr_t, (size_t)B) : nullptr;
    k.xdes = (!use_win && !use_t) ? pk.put(xdes, (size_t)B * NX) : nullptr;
    k.xi_override = xi_override ? pk.put(xi_override, (size_t)B * P * H * 6) : nullptr;
    k.rng = xi_override ? nullptr : reinterpret_cast<const unsigned long long*>(pk.put(rng, (size_t)B * 2));
    CUDA_TRY(cudaMemcpyAsync(h->d_in, h->h_in, pk.off, cudaMemcpyHostToDevice, h->stream));
    Packer po{nullptr, h->d_out};
    k.x_evol = po.reserve<float>((size_t)B * (H + 1) * NX);
    float* d_grad = po.reserve<float>((size_t)B * n);
    k.grad_out = grad ? d_grad : nullptr;
    k.cost_out = po.reserve<float>((size_t)B);
    k.wimg = h->d_wimg; k.traj = h->d_traj; k.T = h->T; k.mtape_g = h->d_mtape;
    h->staged_ok = false;
    CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    if (h->cfg.flags & SDEMPC_F_TENSOR) {
        // tensor-core path: explicit opt-in, never a silent substitute (it is not SPEC-ARITH: TF32 products)
        // rows = problems x particles; the particles of a problem are adjacent lanes of one warp
        if (!h->kc.rollout_tc || P > 32 || (P & (P - 1)) != 0 || h->cfg.u_slew_constr_coeff != 0.0f)
            return fail(SDEMPC_EINVAL, "SDEMPC_F_TENSOR: the tensor-core rollout supports 1, 2, 4, ... 32 particles and no input-rate constraint");
        k.wimg = h->d_wimg_tc;
        const int tgrid = (int)(((size_t)B * P + 127) / 128);
        if (grad) {   // activation / step tape of the adjoint: [CTA][step][granule][128 rows] float4
            const size_t need = (size_t)tgrid * H * h->kc.tc_tape_granules * 128 * 16;
            if (need > h->tape_tc_bytes) {
                if (h->d_tape_tc) cudaFree(h->d_tape_tc);
                h->d_tape_tc = nullptr; h->tape_tc_bytes = 0;
                CUDA_TRY(cudaMalloc(&h->d_tape_tc, need));
                h->tape_tc_bytes = need;
            }
            k.mtape_g = reinterpret_cast<float2*>(h->d_tape_tc);
        }
        void* args[] = {&k};
        CUDA_TRY(cudaLaunchKernel(reinterpret_cast<const void*>(grad ? h->kc.rollout_tc_grad : h->kc.rollout_tc), dim3(tgrid), dim3(128), args,
                                  (size_t)(grad ? h->kc.tc_bytes_grad : h->kc.tc_bytes), h->stream));
        h->launches += 1;
        h->last_grid = tgrid;
    } else {
        if ((rc = launch(h, h->kc.rollout, k, grid))) return rc;
        h->last_grid = grid;
    }
    CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_out, h->d_out, po.off, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    const size_t o_p = a16((size_t)B * (H + 1) * NX * 4), o_c = o_p + a16((size_t)B * n * 4);
    if (x_evol) memcpy(x_evol, h->h_out, (size_t)B * (H + 1) * NX * 4);
    if (grad) memcpy(grad, h->h_out + o_p, (size_t)B * n * 4);
    memcpy(cost, h->h_out + o_c, (size_t)B * 4);
    return 0;
}

int sdempc_closed_loop(sdempc_t* h, int R, int ticks, const float* x0, const float* t0, const uint64_t* rng, float* x_hist,
                       float* u_hist, float* stats) {
    if (!h || R < 1 || ticks < 1 || !x0 || !t0 || !rng || !stats) return fail(SDEMPC_EINVAL, "bad argument");
    if (h->traj_int.empty()) return fail(SDEMPC_ESTATE, "closed loop needs sdempc_set_trajectory() first");
    int rc = ensure_device(h);
    if (rc) return rc;
    const int NU = h->cfg.nu;
    const size_t in_bytes = a16((size_t)R * NX * 4) + a16((size_t)R * 4) + a16((size_t)R * 16) + 64;
    const size_t xh = x_hist ? a16((size_t)R * (ticks + 1) * NX * 4) : 0, uh = u_hist ? a16((size_t)R * ticks * NU * 4) : 0;
    const size_t out_bytes = a16((size_t)R * 16) + xh + uh + 64;
    if ((rc = ensure_io(h, in_bytes, out_bytes))) return rc;
    const bool spec = use_spec(h, R), cl = use_cluster(h, R);
    const int grid = cl ? 2 * R : spec ? std::min(R, h->sm_count) : grid_for(h, R);
    if ((rc = ensure_mtape(h, grid))) return rc;
    KParams k = h->kp;
    Packer pk{h->h_in, h->d_in};
    k.B = R; k.ticks = ticks;
    k.x = pk.put(x0, (size_t)R * NX);
    k.t0 = pk.put(t0, (size_t)R);
    k.rng = reinterpret_cast<const unsigned long long*>(pk.put(rng, (size_t)R * 2));
    CUDA_TRY(cudaMemcpyAsync(h->d_in, h->h_in, pk.off, cudaMemcpyHostToDevice, h->stream));
    Packer po{nullptr, h->d_out};
    k.stats = po.reserve<float>((size_t)R * 4);
    k.x_hist = x_hist ? po.reserve<float>((size_t)R * (ticks + 1) * NX) : nullptr;
    k.u_hist = u_hist ? po.reserve<float>((size_t)R * ticks * NU) : nullptr;
    k.wimg = h->d_wimg; k.traj = h->d_traj; k.T = h->T; k.mtape_g = h->d_mtape;
    h->staged_ok = false;
    CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    if ((rc = launch(h, cl ? h->kc.closed_cl : spec ? h->kc.closed_spec : h->kc.closed, k, grid))) return rc;
    CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    h->last_grid = grid;
    CUDA_TRY(cudaMemcpyAsync(h->h_out, h->d_out, po.off, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    size_t off = 0;
    memcpy(stats, h->h_out + off, (size_t)R * 16); off += a16((size_t)R * 16);
    if (x_hist) { memcpy(x_hist, h->h_out + off, (size_t)R * (ticks + 1) * NX * 4); off += xh; }
    if (u_hist) { memcpy(u_hist, h->h_out + off, (size_t)R * ticks * NU * 4); off += uh; }
    return 0;
}

int sdempc_probe_fp32(int device, float* tflops) {
    if (!tflops) return fail(SDEMPC_EINVAL, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(SDEMPC_ECUDA, "no usable CUDA device %d", device);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int grid = prop.multiProcessorCount, threads = 512, iters = 20000;
    float* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)grid * threads * 4));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {   // first launch warms the clocks up; best of the rest
        CUDA_TRY(cudaEventRecord(e0, 0));
        fp32_probe_kernel<<<grid, threads>>>(d, iters);
        CUDA_TRY(cudaEventRecord(e1, 0));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CUDA_TRY(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    const double flop = 2.0 * 64.0 * iters * (double)grid * threads;
    *tflops = (float)(flop / (best * 1e-3) / 1e12);
    return 0;
}

int64_t sdempc_launch_count(const sdempc_t* h) { return h ? h->launches : 0; }
float sdempc_last_launch_ms(const sdempc_t* h) { return h ? h->last_ms : 0.f; }

int sdempc_kernel_info(sdempc_t* h, int32_t out[6]) {
    if (!h || !out) return fail(SDEMPC_EINVAL, "null argument");
    if (h->staged_tc) {
        out[0] = 128; out[1] = h->kc.tc_bytes_solve; out[2] = h->staged.tcs_ppc; out[3] = h->staged_tc_rate ? h->regs_tc_rate : h->staged_tc_spec ? h->regs_tc_spec : h->staged_tc_lat ? h->regs_tc_lat : h->regs_tc; out[4] = h->last_grid; out[5] = h->sm_count;
        return 0;
    }
    out[0] = h->staged_pcw ? pcw_wpc(h->kc.P) * 32 : (h->staged_cl || h->staged_pc) ? SPEC_LSW * 32 : h->staged_spec ? (SPEC_LSW + SPEC_SGW) * 32 : h->staged_group ? GROUP_GW * 32 : h->kc.G * h->kc.P * 32;
    out[1] = (int32_t)(h->staged_spec ? h->smem_bytes_spec : h->staged_group ? h->smem_bytes_group : h->smem_bytes);
    out[2] = (h->staged_spec || h->staged_pc) ? 1 : h->staged_group ? GROUP_GW * h->kc.gp : h->kc.G;
    out[3] = h->regs;
    out[4] = h->last_grid;
    out[5] = h->sm_count;
    return 0;
}

}  // extern "C"
