#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for p in 8 16 32; do
  timeout 300 python tools/opbench.py --iters 30 --pairs $p --snip-pairs $p --cases snip_enc_N1,snip_dec_N1,enc_N1,enc_N8,dec_N1 > gpurun_out/opbench_p$p.jsonl 2>> gpurun_out/opbench.err
  timeout 300 python tools/opbench.py --iters 20 --flush --pairs $p --snip-pairs $p --cases snip_enc_N1,enc_N1 > gpurun_out/opbench_flush_p$p.jsonl 2>> gpurun_out/opbench.err
done
timeout 300 python tools/opbench.py --iters 30 --regime uniform --cases enc_N1 > gpurun_out/opbench_uniform.jsonl 2>> gpurun_out/opbench.err
cat gpurun_out/opbench_p*.jsonl gpurun_out/opbench_flush_p*.jsonl gpurun_out/opbench_uniform.jsonl | cut -c1-200
tail -5 gpurun_out/opbench.err
# full ncu capture: per-call fwd+bwd (enc_N1) and fused layer fwd+bwd
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_ -c 6 -o gpurun_out/prof_run2 python tools/opbench.py --iters 1 --cases snip_enc_N1 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_ -c 6 -o gpurun_out/prof_run2_percall python tools/opbench.py --iters 1 --cases enc_N1 >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -30
