#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
for r in init local; do
  timeout 300 python tools/opbench.py --iters 30 --regime $r --cases snip_enc_N1,snip_dec_N1,enc_N1,dec_N1 >> gpurun_out/opbench_run5.jsonl 2>> gpurun_out/opbench.err
done
timeout 300 python tools/opbench.py --iters 30 --regime init --pairs 32 --snip-pairs 32 --cases snip_enc_N1,enc_N1 >> gpurun_out/opbench_run5.jsonl 2>> gpurun_out/opbench.err
cut -c1-170 gpurun_out/opbench_run5.jsonl; tail -3 gpurun_out/opbench.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-1500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
