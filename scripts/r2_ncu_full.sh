#!/bin/bash
# on the GPU box: one `ncu --set full` capture (with source counters) of the giant-step kernel for build/variants/<name>.so
# usage: scripts/r2_ncu_full.sh name [bench args]   -> gpurun_out/<name>.ncu-rep
cd "$(dirname "$0")/.."
n=$1; shift
lib=$PWD/build/variants/$n.so; [ "$n" = "default" ] && lib=$PWD/bsgs-cuda_b200/libbsgs_b200.so
BSGS_B200_LIB=$lib timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_giant_step -s 3 -c 1 -f -o gpurun_out/$n \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-solve --no-extras "$@" > gpurun_out/ncu_full_$n.log 2>&1
ls -la gpurun_out/$n.ncu-rep
