// kern_a.cu — instantiation unit of the FP32 (SPEC-ARITH) kernels: (4, 32, 1) (nu, width, particles).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -c kern_a.cu
#include "mpc_entry.cuh"

namespace sdempc {
KernelChoice choice_4_32_1() { return make_choice<4, 32, 1, 8>(); }
}  // namespace sdempc
