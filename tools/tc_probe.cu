// tc_probe.cu — bring-up of the tcgen05 pieces a tensor-core MLP path needs (round-2 design):
// D[128 x N] = A[128 x K] * B[N x K]^T with kind::tf32, one CTA of 128 threads (thread = row = TMEM lane),
// B in shared memory (K-major, no swizzle: 8 x 16-byte core matrices), A either in shared memory (SS) or in
// tensor memory (TS, written by the row's own thread with tcgen05.st), accumulator read back with
// tcgen05.ld 32x32b.  Checks both against a CPU product and prints the error.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/tc_probe.cu -o /tmp/tc_probe && /tmp/tc_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int M = 128, N = 64, K = 72;          // K multiple of 8 (one tf32 MMA consumes K = 8)
constexpr int KC = K / 4;                       // 16-byte chunks along K
constexpr int SBO = KC * 128;                   // bytes between 8-row groups (all K chunks of a group contiguous)
constexpr int LBO = 128;                        // bytes between adjacent K chunks (core matrices)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
    return d;                                    // base offset 0, layout type 0 = no swizzle
}

// element (row, k) of a K-major operand tile -> byte offset
__device__ __host__ __forceinline__ int tile_off(int row, int k) { return (row / 8) * SBO + (k / 4) * LBO + (row % 8) * 16 + (k % 4) * 4; }

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem), "l"(da), "l"(db),
                 "r"(idesc), "r"(acc), "r"(0u));
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(db),
                 "r"(idesc), "r"(acc), "r"(0u));
}
__device__ __forceinline__ void ld16(uint32_t taddr, float* o) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])));
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(parity)
                 : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* B, float* Dss, float* Dts) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                       // 16 row groups x SBO
    unsigned char* sB = smem + 16 * SBO;            // N / 8 row groups x SBO
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int t = threadIdx.x, warp = t >> 5;

    for (int k = 0; k < K; ++k) *reinterpret_cast<float*>(sA + tile_off(t, k)) = A[t * K + k];
    for (int i = t; i < N * K; i += 128) *reinterpret_cast<float*>(sB + tile_off(i / K, i % K)) = B[i];
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    const uint32_t lane_addr = tb + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

    // ---- SS: A and B from shared memory -> D at columns [0, N) ----
    if (t == 0) {
        for (int k8 = 0; k8 < K / 8; ++k8)
            mma_ss(tb, make_desc(smem_u32(sA) + k8 * 2 * LBO), make_desc(smem_u32(sB) + k8 * 2 * LBO), idesc, k8 > 0);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait_parity(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N; c += 16) {
        float v[16];
        ld16(lane_addr + c, v);
        for (int i = 0; i < 16; ++i) Dss[t * N + c + i] = v[i];
    }
    // ---- TS: A from tensor memory (columns [128, 128 + K)), written by the row's own thread ----
    for (int k = 0; k < K; k += 8) {
        float v[8];
        for (int i = 0; i < 8; ++i) v[i] = A[t * K + k + i];
        st8(lane_addr + 128 + k, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (t == 0) {
        for (int k8 = 0; k8 < K / 8; ++k8) mma_ts(tb + 64, tb + 128 + k8 * 8, make_desc(smem_u32(sB) + k8 * 2 * LBO), idesc, k8 > 0);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait_parity(smem_u32(&bar), 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N; c += 16) {
        float v[16];
        ld16(lane_addr + 64 + c, v);
        for (int i = 0; i < 16; ++i) Dts[t * N + c + i] = v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u));
}

int main() {
    std::vector<float> A(M * K), B(N * K), D(M * N), Dss(M * N), Dts(M * N);
    srand(1);
    for (auto& v : A) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    for (auto& v : B) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[i * K + k] * B[j * K + k];
            D[i * N + j] = (float)s;
        }
    float *dA, *dB, *dS, *dT;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dS, D.size() * 4); cudaMalloc(&dT, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dS, 0, D.size() * 4); cudaMemset(dT, 0, D.size() * 4);
    const int smem = (16 + N / 8) * SBO;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(dA, dB, dS, dT);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(Dss.data(), dS, D.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(Dts.data(), dT, D.size() * 4, cudaMemcpyDeviceToHost);
    double ess = 0, ets = 0, ref = 0;
    for (int i = 0; i < M * N; ++i) { ess = fmax(ess, fabs(Dss[i] - D[i])); ets = fmax(ets, fabs(Dts[i] - D[i])); ref = fmax(ref, fabs(D[i])); }
    printf("max |D| %.4f   SS max abs err %.3e   TS max abs err %.3e   (tf32 products: expect ~1e-3 * sqrt(K))\n", ref, ess, ets);
    printf("D[5][7] ref %.5f ss %.5f ts %.5f ; D[100][63] ref %.5f ss %.5f ts %.5f\n", D[5 * N + 7], Dss[5 * N + 7], Dts[5 * N + 7],
           D[100 * N + 63], Dss[100 * N + 63], Dts[100 * N + 63]);
    return 0;
}
