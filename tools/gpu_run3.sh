#!/bin/bash
mkdir -p gpurun_out
for r in init local; do
for p in 16 32; do
  timeout 300 python tools/opbench.py --iters 30 --regime $r --pairs $p --snip-pairs $p --cases snip_enc_N1,enc_N1 >> gpurun_out/opbench_run3.jsonl 2>> gpurun_out/opbench.err
done; done
cut -c1-175 gpurun_out/opbench_run3.jsonl
tail -5 gpurun_out/opbench.err
# ncu: warmup 0, iters 1 (inner=10 -> launches: fwd x10, [memset+bwd] x10, bwd_nomemset x10)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 9 -c 3 -o gpurun_out/prof_run3_local python tools/opbench.py --iters 1 --warmup 0 --cases enc_N1 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 9 -c 3 -o gpurun_out/prof_run3_init python tools/opbench.py --iters 1 --warmup 0 --regime init --cases enc_N1 >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head
