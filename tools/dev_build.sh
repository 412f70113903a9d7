#!/bin/bash
# Quick experiment build: only the iris P=1 kernels, into gpurun_dev/libsdempc_dev.so (extra -D flags as arguments).
#   tools/dev_build.sh -DSOME_EXPERIMENT && SDEMPC_LIB=gpurun_dev/libsdempc_dev.so python tools/profile_solve.py
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_dev
env -u CC -u CXX /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
    -shared -Xcompiler -fPIC -cudart static -DSDEMPC_DEV_IRIS_ONLY "$@" \
    sde4mbrl_px4_b200/csrc/sdempc_api.cu -o gpurun_dev/libsdempc_dev${DEV_TAG}.so
