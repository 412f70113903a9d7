#!/bin/bash
# Build a variant of ONE translation unit from a scratch copy of csrc (edited by the caller) and link it with the current
# objects of the other units into build/ab/lib<name>.so, for A/B runs inside one gpurun call (SDEMPC_LIB=...).
#   tools/ab_variant.sh <name> <unit.cu> <scratch csrc dir> [extra nvcc flags]
set -e
name=$1; unit=$2; src=$3; shift 3
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/build/ab
obj=$root/build/ab/${name}_${unit%.cu}.o
flags="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
case $unit in sdempc_tc.cu) ;; *) flags="$flags -fmad=false";; esac
/usr/local/cuda/bin/nvcc $flags "$@" -I$root/sde4mbrl_px4_b200/csrc -c $src/$unit -o $obj
others=$(ls $root/build/*.o | grep -v "/${unit%.cu}.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $root/build/ab/lib$name.so $others $obj
echo built $root/build/ab/lib$name.so
