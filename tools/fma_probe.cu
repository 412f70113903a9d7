// fma_probe.cu — what does the FP32 pipe of one SM sustain for the instruction forms the MPC kernels use?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false tools/fma_probe.cu -o gpurun_out/fma_probe && gpurun_out/fma_probe
// Prints warp-instructions per cycle per SM sub-partition and the equivalent chip TFLOP/s for
// FFMA (3 distinct register operands), FFMA with a shared multiplicand, FFMA2 (fma.rn.f32x2), FMUL+FADD pairs,
// with 1, 2 and 4 resident warps per sub-partition.  Evidence for DESIGN.md section 5 "bounding roofline".
#include <cuda_runtime.h>
#include <cstdio>

template <int KIND>
__global__ void probe(float* out, int iters, long long* cyc) {
    float a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 1.0f + threadIdx.x * 1e-6f + i; b[i] = 0.999f + i * 1e-7f; c[i] = 0.5f * i; }
    float2 p[4], q[4], r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { p[i] = make_float2(a[i], a[i + 4]); q[i] = make_float2(b[i], b[i + 4]); r[i] = make_float2(c[i], c[i + 4]); }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
            if (KIND == 0) {          // FFMA, three distinct registers per instruction
#pragma unroll
                for (int i = 0; i < 8; ++i) c[i] = __fmaf_rn(a[i], b[i], c[i]);
            } else if (KIND == 1) {   // FFMA, one operand shared by consecutive instructions (operand reuse)
#pragma unroll
                for (int i = 0; i < 8; ++i) c[i] = __fmaf_rn(a[i], b[0], c[i]);
            } else if (KIND == 2) {   // FFMA2
#pragma unroll
                for (int i = 0; i < 4; ++i) r[i] = __ffma2_rn(p[i], q[i], r[i]);
            } else if (KIND == 3) {   // FMUL + FADD (no contraction)
#pragma unroll
                for (int i = 0; i < 8; i += 2) { c[i] = __fmul_rn(a[i], c[i]); c[i + 1] = __fadd_rn(b[i], c[i + 1]); }
            } else if (KIND == 4) {   // ONE dependent FFMA chain: latency
#pragma unroll
                for (int i = 0; i < 8; ++i) c[0] = __fmaf_rn(c[0], b[i], a[i]);
            } else if (KIND == 5) {   // ONE dependent FFMA2 chain: latency
#pragma unroll
                for (int i = 0; i < 8; ++i) r[0] = __ffma2_rn(r[0], q[i & 3], p[i & 3]);
            } else if (KIND == 6) {   // two independent FFMA chains (what a scalar two-network tanh would be)
#pragma unroll
                for (int i = 0; i < 8; ++i) { c[0] = __fmaf_rn(c[0], b[i], a[i]); c[1] = __fmaf_rn(c[1], b[i], a[i]); }
            } else {                  // two independent FFMA2 chains
#pragma unroll
                for (int i = 0; i < 8; ++i) { r[0] = __ffma2_rn(r[0], q[i & 3], p[i & 3]); r[1] = __ffma2_rn(r[1], q[i & 3], p[i & 3]); }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += r[i].x + r[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMallocManaged(&cyc, sizeof(long long));
    const int iters = 20000;
    const char* names[8] = {"FFMA  (3 distinct regs)", "FFMA  (shared operand) ", "FFMA2 (f32x2)          ", "FMUL+FADD pairs        ",
                            "FFMA  1 dependent chain", "FFMA2 1 dependent chain", "FFMA  2 chains         ", "FFMA2 2 chains         "};
    const int per_iter[8] = {32, 32, 16, 32, 32, 32, 64, 64};        // warp-instructions per loop iteration
    const double flop_per_inst[8] = {64, 64, 128, 32, 64, 128, 64, 128};   // flop per warp-instruction (32 lanes)
    for (int kind = 0; kind < 8; ++kind)
        for (int wps = 1; wps <= 4; wps *= 2) {
            const int threads = 4 * wps * 32;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (kind == 0) probe<0><<<148, threads>>>(out, iters, cyc);
                if (kind == 1) probe<1><<<148, threads>>>(out, iters, cyc);
                if (kind == 2) probe<2><<<148, threads>>>(out, iters, cyc);
                if (kind == 3) probe<3><<<148, threads>>>(out, iters, cyc);
                if (kind == 4) probe<4><<<148, threads>>>(out, iters, cyc);
                if (kind == 5) probe<5><<<148, threads>>>(out, iters, cyc);
                if (kind == 6) probe<6><<<148, threads>>>(out, iters, cyc);
                if (kind == 7) probe<7><<<148, threads>>>(out, iters, cyc);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double inst_per_smsp = (double)iters * per_iter[kind] * wps;
            const double ipc = inst_per_smsp / (double)*cyc;
            const double tflops = 148.0 * 4 * inst_per_smsp * flop_per_inst[kind] / (ms * 1e-3) / 1e12;
            printf("%s  warps/sub-partition %d : %.3f warp-inst/cycle/sub-partition, %.1f TFLOP/s chip-wide (%.2f ms)\n", names[kind], wps, ipc, tflops, ms);
        }
    return 0;
}
