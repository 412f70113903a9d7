// sdempc_tc.cu — translation unit of the tensor-core kernels (tcgen05 + TMEM + TMA): the batched rollout /
// value_and_grad of mpc_tc.cuh and the batched APG solve of mpc_tcsolve.cuh.  Host code reaches them through tc_api.h.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -c sdempc_tc.cu
#include <algorithm>

#include "mpc_tcsolve.cuh"
#include "tc_api.h"

using namespace sdempc;

// Resident CTAs per SM are bounded by tensor memory (512 columns): the register budget is set to allow exactly that many.
template <int NU, int W, bool GRAD>
__global__ void __launch_bounds__(128, 512 / TCLayout<NU, W>::COLS) mpc_tc_rollout_kernel(const __grid_constant__ KParams P) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t tc_bar[2];
    tc_rollout_body<NU, W, GRAD>(P, tc_smem, &tmem_slot, tc_bar);
}

// MINB = resident CTAs per SM the register budget is set for.  Two builds of the width-32 solve: 4 CTAs per SM (128 registers,
// what tensor memory allows: large batches) and 2 CTAs per SM (164 registers, no spill: 11 % faster per CTA, used while the
// batch needs at most two CTAs per SM).  SPECG: the build with the speculative gradient pass and per-problem phases
// (mpc_tcsolve.cuh), for CTAs that own at most a quarter of their slots in problems.
// RATE: with the soft input-rate constraint (one more build, of the plain two-CTA kernel).
template <int NU, int W, int MINB, bool SPECG, bool RATE = false>
__global__ void __launch_bounds__(128, MINB) mpc_tc_solve_kernel(const __grid_constant__ KParams P) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t tc_bars[2];
    __shared__ TCSShared sh;
    tc_solve_body<NU, W, SPECG, (MINB <= 2), RATE>(P, tc_smem, &tmem_slot, tc_bars, sh);
}

template <int NU, int W>
static TCKernels make_tc() {
    using L = TCSLayout<NU, W>;
    TCKernels k;
    k.rollout = mpc_tc_rollout_kernel<NU, W, false>;
    k.rollout_grad = mpc_tc_rollout_kernel<NU, W, true>;
    k.solve = mpc_tc_solve_kernel<NU, W, 512 / L::COLS, false>;
    k.solve_lat = (512 / L::COLS > 2) ? mpc_tc_solve_kernel<NU, W, 2, false> : k.solve;
    k.solve_spec = mpc_tc_solve_kernel<NU, W, 2, true>;
    k.solve_rate = mpc_tc_solve_kernel<NU, W, 2, false, true>;
    // Residency must be bounded by tensor memory (512 / COLS CTAs per SM), never exceed it: a CTA that the block
    // scheduler places beyond that spins in tcgen05.alloc while holding its slot (width 64, forward variant:
    // registers and shared memory allowed three CTAs, tensor memory two -- launches were bimodal, 0.30 / 0.41 ms).
    // Where registers do not already impose the bound, the dynamic shared-memory request is padded so that one
    // more CTA cannot fit (228 KB per SM, 1 KB reserved per CTA).
    constexpr int tmem_ctas = 512 / L::COLS;
    constexpr int pad = tmem_ctas < 4 ? (228 * 1024) / (tmem_ctas + 1) + 1024 : 0;
    k.bytes = std::max((int)L::BYTES, pad);
    k.bytes_grad = std::max((int)L::BYTES_GRAD, pad);
    k.bytes_solve = std::max((int)L::BYTES_GRAD, pad > (int)sizeof(TCSShared) ? pad - (int)sizeof(TCSShared) : 0);
    k.tape_granules = TCLayout<NU, W>::TG;
    k.solve_tape_granules = L::STG;
    k.cols = L::COLS;
    return k;
}

namespace sdempc {
const TCKernels* tc_kernels(int nu, int width) {
    static const TCKernels k4_32 = make_tc<4, 32>();
#ifndef SDEMPC_DEV_IRIS_ONLY
    static const TCKernels k6_32 = make_tc<6, 32>(), k4_64 = make_tc<4, 64>(), k6_64 = make_tc<6, 64>();
    if (nu == 6 && width == 32) return &k6_32;
    if (nu == 4 && width == 64) return &k4_64;
    if (nu == 6 && width == 64) return &k6_64;
#endif
    if (nu == 4 && width == 32) return &k4_32;
    return nullptr;
}
}  // namespace sdempc
