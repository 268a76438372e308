set -u
O=gpurun_out/r02
mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --only-main --no-cpu-baseline --no-e2e"
python bench.py --impl reference > $O/bench_n1_reference.json 2>/dev/null
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
timeout 600 $B > $O/bench_plain.json 2> $O/bench_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $O/launches_r02.csv $B > $O/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_scan_sampled|k_resolve_queue" -s 8 -c 2 -f -o $O/sampled $B > $O/sampled.ncu.log 2>&1
ls -la $O/sampled.ncu-rep $O/launches_r02.csv
