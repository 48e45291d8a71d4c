# DRAM bytes and duration of one 16-window launch of the giant-step kernel (ncu; never a bench value)
m="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
run() { label=$1; shift; ncu --metrics $m --clock-control none -k regex:k_giant_step -s 3 -c 1 --csv "$@" 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' -v l="$label" '{print l, $(NF-2), $(NF-1), $NF}'; }
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-solve"
run base $B
run l2fetch32 $B --l2-fetch 32
