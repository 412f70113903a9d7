// kern_c.cu — instantiation unit of the FP32 (SPEC-ARITH) kernels: (6, 32, 1), (6, 32, 8) (nu, width, particles).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -c kern_c.cu
#include "mpc_entry.cuh"

namespace sdempc {
KernelChoice choice_6_32_1() { return make_choice<6, 32, 1, 8>(); }
KernelChoice choice_6_32_8() { return make_choice<6, 32, 8, 1>(); }
}  // namespace sdempc
