m="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
run() { label=$1; shift; ncu --metrics $m --clock-control none -k regex:k_giant_step -s 3 -c 1 --csv "$@" 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' -v l="$label" '{print l, $(NF-2), $(NF-1), $NF}'; }
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-solve"
run base $B
run l2fetch32 $B --l2-fetch 32
run l2fetch64 $B --l2-fetch 64
BSGS_B200_LIB=$PWD/bsgs-cuda_b200/libexp_ld1.so run ld_64B_hint $B
BSGS_B200_LIB=$PWD/bsgs-cuda_b200/libexp_ld2.so run ld_plain_ldg $B
BSGS_B200_LIB=$PWD/bsgs-cuda_b200/libexp_ld3.so run ld_noalloc_2x128 $B
for v in 1 2 3; do BSGS_B200_LIB=$PWD/bsgs-cuda_b200/libexp_ld$v.so python bench.py --steps 8 --no-cpu --no-solve | python -c "import sys,json; j=json.loads(sys.stdin.readlines()[-1]); print('RESULT ld$v', round(j['value']/1e9,2))"; done
