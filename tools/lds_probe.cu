// lds_probe.cu — shared-memory cost of the LDS.128 address patterns the MPC kernels use (one SM, W warps).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/lds_probe.cu -o gpurun_out/lds_probe && gpurun_out/lds_probe
// Prints cycles per warp-level LDS.128 (throughput, many independent loads in flight) for:
//   0 uniform            : all 32 lanes read the same 16 bytes
//   1 lane-distinct      : 32 consecutive 16-byte chunks (512 B)
//   2 8 roles x 4 groups : lane reads chunk (lane & 7)  [quad kernel, 8 lanes per problem: weights]
//   3 4 roles x 8 groups : lane reads chunk (lane & 3)  [quad kernel, 4 lanes per problem: weights]
//   4 one per quarter    : lane reads chunk (lane >> 3) * 9   [quad kernel activations, 4 problems]
//   5 one per quad       : lane reads chunk (lane >> 2) * 9   [quad kernel activations, 8 problems]
//   6 2 per quarter      : lane reads chunk ((lane >> 3) * 2 + ((lane >> 2) & 1))
//   7 LDS.64 lane-distinct (256 B)
//   8 LDS.32 uniform
#include <cuda_runtime.h>
#include <cstdio>

template <int KIND>
__global__ void probe(float* out, int iters, long long* cyc) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int idx;
    if (KIND == 0 || KIND == 8) idx = 0;
    else if (KIND == 1 || KIND == 7) idx = lane;
    else if (KIND == 2) idx = lane & 7;
    else if (KIND == 3) idx = lane & 3;
    else if (KIND == 4) idx = (lane >> 3) * 9;
    else if (KIND == 5) idx = (lane >> 2) * 9;
    else idx = (lane >> 3) * 2 + ((lane >> 2) & 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int a = (idx + 64 * u + it) & 4095;
            if (KIND == 7) { const float2 v = *reinterpret_cast<const float2*>(&sm[a]); acc.x += v.x; acc.y += v.y; }
            else if (KIND == 8) { const float v = *reinterpret_cast<const float*>(&sm[a]); acc.x += v; }
            else { const float4 v = sm[a]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int KIND>
void run(const char* name, float* out, long long* cyc) {
    for (int warps : {1, 4, 8, 16}) {
        const int iters = 2000;
        cudaFuncSetAttribute(probe<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        probe<KIND><<<1, warps * 32, 65536>>>(out, iters, cyc);
        probe<KIND><<<1, warps * 32, 65536>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
        const double loads = (double)iters * 16 * warps;
        printf("%-22s warps %2d : %.2f cycles per warp-LDS (SM-wide)\n", name, warps, (double)c / loads);
    }
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    run<0>("uniform .128", out, cyc);
    run<1>("lane-distinct .128", out, cyc);
    run<2>("8 roles x 4 .128", out, cyc);
    run<3>("4 roles x 8 .128", out, cyc);
    run<4>("one per quarter .128", out, cyc);
    run<5>("one per quad .128", out, cyc);
    run<6>("two per quarter .128", out, cyc);
    run<7>("lane-distinct .64", out, cyc);
    run<8>("uniform .32", out, cyc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
