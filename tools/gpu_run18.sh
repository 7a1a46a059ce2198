#!/bin/bash
# run 18: ncu launch list + one full capture of the dominant kernel, both on the bench.py command itself
mkdir -p gpurun_out/run18
R=gpurun_out/run18
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > $R/bench_under_ncu.log 2>&1
wc -l $R/launches_bench.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_snippet_fwd --launch-skip 24 -c 1 -o $R/prof_bench_snippet_fwd python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > $R/bench_under_ncu_full.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $R/bench_n1.json 2> $R/bench_n1.err; cut -c1-300 $R/bench_n1.json
ls -la $R
