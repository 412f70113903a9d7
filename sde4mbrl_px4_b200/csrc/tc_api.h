// tc_api.h — what sdempc_api.cu needs to know about the tensor-core translation unit (sdempc_tc.cu).
// The tensor-core kernels are compiled separately from the FP32 kernels: they are not SPEC-ARITH (no -fmad=false
// requirement), and two translation units compile in parallel.
#pragma once
#include "mpc_kernels.cuh"

namespace sdempc {

struct TCKernels {
    void (*rollout)(KParams);        // batched cost evaluation (mpc_tc.cuh)
    void (*rollout_grad)(KParams);   // ... with the adjoint sweep
    void (*solve)(KParams);          // batched APG solve (mpc_tcsolve.cuh), register budget for 512 / cols CTAs per SM
    void (*solve_spec)(KParams);     // two-CTA register budget + speculative gradient pass: at most slots / 4 problems per CTA
    void (*solve_rate)(KParams);     // two-CTA register budget + the soft input-rate constraint (u_slew_constr configurations)
    void (*solve_lat)(KParams);      // same kernel with the register budget of two CTAs per SM (== solve when cols = 256)
    int bytes, bytes_grad, bytes_solve;   // dynamic shared memory of each (already padded so that residency <= tensor memory)
    int tape_granules;               // rollout_grad: 16-byte tape granules per row and step
    int solve_tape_granules;         // solve: same for the GRAD rows
    int cols;                        // tensor-memory columns per CTA
};

// nullptr when (nu, width) is not compiled
const TCKernels* tc_kernels(int nu, int width);

}  // namespace sdempc
