#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
for r in init local; do
  timeout 300 python tools/opbench.py --iters 30 --regime $r --cases snip_enc_N1,snip_dec_N1,enc_N1,enc_N8,dec_N1 >> gpurun_out/opbench_run7.jsonl 2>> gpurun_out/opbench.err
done
timeout 300 python tools/opbench.py --iters 30 --regime init --pairs 32 --snip-pairs 32 --cases snip_enc_N1,enc_N1 >> gpurun_out/opbench_run7.jsonl 2>> gpurun_out/opbench.err
timeout 300 python tools/opbench.py --iters 30 --regime init --pairs 8 --snip-pairs 8 --cases snip_enc_N1,enc_N1 >> gpurun_out/opbench_run7.jsonl 2>> gpurun_out/opbench.err
cut -c1-170 gpurun_out/opbench_run7.jsonl; tail -3 gpurun_out/opbench.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 9 -c 1 -o gpurun_out/prof_run7_fwd python tools/opbench.py --iters 1 --warmup 0 --regime init --cases enc_N1 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_bwd -s 9 -c 1 -o gpurun_out/prof_run7_bwd python tools/opbench.py --iters 1 --warmup 0 --regime init --cases enc_N1 >> gpurun_out/ncu_full.log 2>&1
