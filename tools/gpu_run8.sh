#!/bin/bash
# run 8: parity after the dot-first backward / padded smem records / small-grid tiles; op bench incl.
# deterministic backward; where an end-to-end step goes (torch.profiler); ncu of the four hot kernels.
mkdir -p gpurun_out
R=gpurun_out/run8
mkdir -p $R
timeout 900 python -m pytest tests -m gpu -x -q > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -3 $R/pytest_gpu.log
for regime in init local; do
  timeout 300 python tools/opbench.py --iters 30 --regime $regime --cases snip_enc_N1,snip_dec_N1,enc_N1,enc_N8,dec_N1 >> $R/opbench.jsonl 2>> $R/opbench.err
done
cut -c1-170 $R/opbench.jsonl
timeout 300 python tools/e2e_profile.py --steps 3 > $R/e2e_profile_infer.json 2> $R/e2e_profile.err
timeout 300 python tools/e2e_profile.py --steps 2 --train --batch 2 > $R/e2e_profile_train.json 2>> $R/e2e_profile.err
cut -c1-1500 $R/e2e_profile_infer.json; cut -c1-1500 $R/e2e_profile_train.json; tail -3 $R/e2e_profile.err
timeout 600 python bench.py --steps 10 --warmup 3 > $R/bench_n1.json 2> $R/bench_n1.err; cut -c1-400 $R/bench_n1.json; tail -2 $R/bench_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_ -c 3 -o $R/prof_percall python tools/opbench.py --iters 1 --warmup 0 --inner 1 --regime local --cases enc_N1 > $R/ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_snippet -c 2 -o $R/prof_snip python tools/opbench.py --iters 1 --warmup 0 --inner 1 --regime local --cases snip_enc_N1 >> $R/ncu.log 2>&1
ls -la $R
