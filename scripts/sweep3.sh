# geometry sweep of the giant-step kernel on config 2 (device-resident throughput only)
run() { timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-solve "$@" 2>>gpurun_out/sweep_err.log | python -c "import sys,json; j=json.loads(sys.stdin.readlines()[-1]); print('RESULT $*:', round(j['value']/1e9,2), round(j['roofline']['frac'],3), j['clocks']['sm_mhz'], j['config']['grid'], j['config']['block'], j['config']['points_per_thread'])"; }
mkdir -p gpurun_out
run
run --block 320
run --blocks-per-sm 2
run --nwin 8
run --block 320 --nwin 8
