#!/bin/bash
# Standard GPU validation (run on a B200 box from the repo root, e.g. `gpurun -- bash tools/gpu_check.sh [outdir]`):
# parity tests, smoke, op-level timings in both location regimes (fp32 + bf16), headline bench.
# Everything is written under gpurun_out/<outdir> (default: check).
R=gpurun_out/${1:-check}
mkdir -p $R
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $R/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -p timeout --timeout=180 > $R/pytest_gpu.log 2>&1   # per-test timeout: a hung kernel must not eat the job; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -15 $R/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $R/smoke.log 2>&1; tail -3 $R/smoke.log
rm -f $R/opbench.jsonl
for regime in init local; do
  timeout 300 python tools/opbench.py --iters 30 --regime $regime --bf16 --cases snip_enc_N1,snip_dec_N1,enc_N1,enc_N8,dec_N1 >> $R/opbench.jsonl 2>> $R/opbench.err
done
tail -5 $R/opbench.err
python - $R <<'PY'
import json, sys
for l in open(sys.argv[1] + '/opbench.jsonl'):
    d = json.loads(l)
    print("%-12s %-22s %-6s %-36s %9.2f us %7.1f GB/s %.4f" % (d['case'], d['impl'], d['regime'], d['pass'], d['us_median'], d['GBps'], d['frac_of_measured_hbm']))
PY
timeout 600 python bench.py --steps 20 --warmup 5 > $R/bench_n1.json 2> $R/bench_n1.err; cat $R/bench_n1.json; tail -3 $R/bench_n1.err
