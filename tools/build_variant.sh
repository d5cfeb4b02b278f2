#!/bin/bash
# dev helper: builds a kernel VARIANT of the library for A/B runs on the GPU box (bench.py / tests pick it up through FULGOR_GPU_LIB):
#   tools/build_variant.sh NAME [SRC_ROOT] [nvcc flags, e.g. -DFG_MIN_BLOCKS=5]
# -> fulgor_b200/variants/libfulgor_gpu_NAME.so (git-ignored, travels to the box). SRC_ROOT defaults to the working tree; a checkout of
# another commit (git archive HEAD fulgor_b200/csrc include | tar -x -C DIR) gives the "before" arm.
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
SRC=${1:-.}; [ $# -gt 0 ] && shift
mkdir -p fulgor_b200/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --cudart static "$@" -shared \
  -o fulgor_b200/variants/libfulgor_gpu_$NAME.so $SRC/fulgor_b200/csrc/engine.cu $SRC/fulgor_b200/csrc/fur_reader.cpp
echo fulgor_b200/variants/libfulgor_gpu_$NAME.so
