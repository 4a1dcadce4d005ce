#!/bin/bash
# Developer tool: builds an A/B variant of libbsw_gpu.so with extra -D flags into
# genarchbench_b200/lib/variants/libbsw_gpu_<name>.so (select with BSW_GPU_LIB=... in scripts/).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
mkdir -p "$ROOT/genarchbench_b200/lib/variants"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I"$ROOT/include" \
  -ccbin /usr/bin/g++ -Xcompiler -fPIC,-Wall,-pthread,-fopenmp "$@" -shared \
  -o "$ROOT/genarchbench_b200/lib/variants/libbsw_gpu_$name.so" "$ROOT/genarchbench_b200/csrc/bsw_gpu.cu"
echo built $name
