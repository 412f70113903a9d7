// kern_b.cu — instantiation unit of the FP32 (SPEC-ARITH) kernels: (4, 32, 2), (4, 32, 4), (4, 32, 8) (nu, width, particles).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -c kern_b.cu
#include "mpc_entry.cuh"

namespace sdempc {
KernelChoice choice_4_32_2() { return make_choice<4, 32, 2, 4>(); }
KernelChoice choice_4_32_4() { return make_choice<4, 32, 4, 2>(); }
KernelChoice choice_4_32_8() { return make_choice<4, 32, 8, 1>(); }
}  // namespace sdempc
