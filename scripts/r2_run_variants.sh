#!/bin/bash
# on the GPU box: device-resident throughput of every build/variants/*.so (or the names given), config 2;
# with NCU=1 also the executed-instruction count of one 16-window launch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
names=("$@"); [ ${#names[@]} -eq 0 ] && names=($(ls build/variants/*.so | xargs -n1 basename | sed 's/\.so$//'))
for n in "${names[@]}"; do
  lib=$PWD/build/variants/$n.so
  [ "$n" = "default" ] && lib=$PWD/bsgs-cuda_b200/libbsgs_b200.so
  BSGS_B200_LIB=$lib timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu --no-solve --no-extras $BENCH_ARGS 2>>gpurun_out/variants_err.log |
    python -c "import sys,json; j=json.loads(sys.stdin.readlines()[-1]); print('VARIANT $n:', round(j['value']/1e9,3), 'G steps/s', round(j['roofline']['frac'],4), 'e2e', round(j['e2e']['value']/1e9,3), j['clocks']['sm_mhz'], 'hits', j['hits_in_e2e'])" | tee -a gpurun_out/variants.log
  if [ -n "$NCU" ]; then
    BSGS_B200_LIB=$lib timeout 600 ncu --metrics smsp__inst_executed.sum,sm__cycles_elapsed.max,gpu__time_duration.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_alu.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_giant_step -s 3 -c 1 --csv --log-file gpurun_out/ncu_$n.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-solve --no-extras $BENCH_ARGS > /dev/null 2>>gpurun_out/variants_err.log
    python - <<PY | tee -a gpurun_out/variants.log
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_$n.csv')) if len(r)>5]
h=rows[0]; mi=h.index('Metric Name'); vi=h.index('Metric Value')
d={r[mi]:r[vi] for r in rows[1:]}
inst=float(d.get('smsp__inst_executed.sum','0').replace(',',''))
pw=444*8*171*16   # warps x point-windows per thread (config 2, 16 windows)
print('  NCU $n: inst/pw=%.1f'%(inst/pw), {k:v for k,v in d.items()})
PY
  fi
done
