// kern_e.cu — instantiation unit of the FP32 (SPEC-ARITH) kernels: (6, 64, 1), (6, 64, 8) (nu, width, particles).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -c kern_e.cu
#include "mpc_entry.cuh"

namespace sdempc {
KernelChoice choice_6_64_1() { return make_choice<6, 64, 1, 8>(); }
KernelChoice choice_6_64_8() { return make_choice<6, 64, 8, 1>(); }
}  // namespace sdempc
