#!/bin/bash
# Round-2 ncu evidence in one gpurun call: every command first plain (must exit 0), then under ncu.
# Outputs under gpurun_out/r02/: launch list (csv), .ncu-rep per kernel group, text summaries.
set -u
O=gpurun_out/r02
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
run() {  # name, kernel regex, skip, count, command...
	local name=$1 k=$2 s=$3 c=$4; shift 4
	echo "== $name: $*"
	if ! timeout 600 "$@" > $O/$name.plain.log 2>&1; then echo "plain run failed"; tail -5 $O/$name.plain.log; return; fi
	tail -3 $O/$name.plain.log
	timeout 900 $NCU -k "regex:$k" -s $s -c $c -f -o $O/$name "$@" > $O/$name.ncu.log 2>&1 || { echo "ncu failed"; tail -5 $O/$name.ncu.log; }
	ls -la $O/$name.ncu-rep 2>/dev/null
}
B="python bench.py --steps 3 --warmup 3 --only-main --no-cpu-baseline --no-e2e"
echo "== launch list"
timeout 600 $B > $O/bench_plain.json 2> $O/bench_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $O/launches_r02.csv $B > $O/launches.log 2>&1
tail -2 $O/launches.log | cut -c1-200
run sampled "k_scan_sampled|k_resolve_queue" 8 2 $B
run sent "k_scan_rd|k_rd_expand" 2 2 python tools/sent_one.py 4 1024 2
run xd_u8 "k_scan_xd" 1 1 python tools/quick_bench.py 1024 dfa 2
run xd_u16 "k_scan_xd" 3 1 python tools/flow_bench.py 256 2000 0
run dense "k_dense_walk" 2 1 python tools/density_sweep.py 512 10000 --only zero:0.2
echo "== cli_bench"
timeout 600 python tools/cli_bench.py 512 4 "-w 1" "-w 4" "-w 4 -G 65536" "-w 8 -G 65536" 2>&1 | tail -6 | tee $O/cli_bench.log
du -sh $O
