# A/B of giant-step kernel builds on config 2 (device-resident throughput only), then the microbenchmark and the GPU tests
mkdir -p gpurun_out
run() { lib=$1; shift; BSGS_B200_LIB=$PWD/bsgs-cuda_b200/$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-solve "$@" 2>gpurun_out/sweep_err.log | python -c "import sys,json; j=json.loads(sys.stdin.readlines()[-1]); print('RESULT $lib $*:', round(j['value']/1e9,2), round(j['roofline']['frac'],3), j['clocks']['sm_mhz'])"; }
for lib in "$@"; do run $lib; done
run libbsgs_b200.so --blocks-per-sm 2
timeout 120 bsgs-cuda_b200/pipe_mix > gpurun_out/pipe_mix.json; tail -12 gpurun_out/pipe_mix.json
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
