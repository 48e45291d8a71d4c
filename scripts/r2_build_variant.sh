#!/bin/bash
# build/variants/<name>.so = libbsgs_b200.so compiled with extra -D flags (A/B experiments; selected with BSGS_B200_LIB)
# usage: scripts/r2_build_variant.sh name [-DFOO=1 ...]
set -e
cd "$(dirname "$0")/../bsgs-cuda_b200"
name=$1; shift
mkdir -p ../build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared "$@" -o ../build/variants/$name.so csrc/bsgs_b200.cu
