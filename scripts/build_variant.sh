#!/bin/bash
# Build a variant of libkasf.so with extra -D flags into variants/<name>.so (A/B measurements; selected with KASF_LIB).
# usage: build_variant.sh <name> [-DKASF_...=..]...
set -e
NAME=$1; shift
mkdir -p variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
  -Xcompiler -O2 --expt-relaxed-constexpr "$@" -o variants/$NAME.so kasportsformer_b200/csrc/*.cu
echo built variants/$NAME.so
