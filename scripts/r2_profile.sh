#!/bin/bash
# on the GPU box: the ncu evidence of the shipped kernel (bsgs-cuda_b200/libbsgs_b200.so), config 2, 16 windows per launch:
#   gpurun_out/final.ncu-rep        one `ncu --set full --import-source on` capture of k_giant_step (read with scripts/ncu_summary.py)
#   gpurun_out/launches_final.csv   per-launch durations of `python bench.py --steps 2 --warmup 1` (all kernels of a bench run)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_giant_step -s 3 -c 1 -f -o gpurun_out/final \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-solve --no-extras > gpurun_out/ncu_full_final.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/launches_final.log 2>&1
ls -la gpurun_out/final.ncu-rep gpurun_out/launches_final.csv
