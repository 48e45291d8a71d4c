timeout 600 python -m pytest tests -m gpu -q -x -k "not reference_kernel" 2>&1 | tail -3
run() { python bench.py --steps 8 --warmup 3 --no-cpu --no-solve "$@" 2>&1 | python -c "import sys,json; j=json.loads(sys.stdin.readlines()[-1]); print('RESULT $*:', round(j['value']/1e9,2), round(j['roofline']['frac'],3))"; }
run --nwin 16 --blocks-per-sm 3
run --nwin 16 --blocks-per-sm 2
run --nwin 8 --blocks-per-sm 3
