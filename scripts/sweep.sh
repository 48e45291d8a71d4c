BSGS_KERNEL=3 timeout 600 python -m pytest tests -m gpu -q -x -k "not reference_kernel" 2>&1 | tail -3
run() { python bench.py --steps 5 --warmup 3 --no-cpu --kernel 3 "$@" 2>&1 | python -c "import sys,json; j=json.loads(sys.stdin.readlines()[-1]); print('RESULT $*:', round(j['value']/1e9,2), round(j['roofline']['frac'],3))"; }
run --variant 5 --nwin 16 --blocks-per-sm 3
run --variant 5 --nwin 16 --blocks-per-sm 3 --l2-fetch 32
run --variant 5 --nwin 16 --blocks-per-sm 3 --l2-fetch 128
run --variant 6 --nwin 16 --blocks-per-sm 3
run --variant 6 --nwin 16 --blocks-per-sm 3 --l2-fetch 32
run --variant 7 --nwin 16 --blocks-per-sm 3
run --variant 8 --nwin 16 --blocks-per-sm 3
run --variant 6 --nwin 16 --blocks-per-sm 2
run --variant 7 --nwin 16 --blocks-per-sm 2
run --variant 8 --nwin 16 --blocks-per-sm 2
run --variant 8 --nwin 16 --blocks-per-sm 2 --l2-fetch 32
