#!/bin/bash
# run 9: deterministic backward after the rank-sort redesign; forward variants; head-major layout experiment
mkdir -p gpurun_out/run9
R=gpurun_out/run9
timeout 900 python -m pytest tests -m gpu -x -q > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -3 $R/pytest_gpu.log
for regime in init local; do
  for v in 0 1 2; do
    timeout 300 python tools/opbench.py --iters 30 --regime $regime --fwd-variant $v --head-major --cases enc_N1,enc_N8 >> $R/opbench.jsonl 2>> $R/opbench.err
  done
done
timeout 300 python tools/opbench.py --iters 30 --regime local --cases dec_N1,snip_dec_N1 >> $R/opbench.jsonl 2>> $R/opbench.err
python - <<'PY'
import json
for l in open('gpurun_out/run9/opbench.jsonl'):
    d=json.loads(l)
    print("%-8s %-16s %-6s v%d %-18s %9.2f us %7.1f GB/s" % (d['case'],d['impl'],d['regime'],d.get('fwd_variant',0),d['pass'],d['us_median'],d['GBps']))
PY
tail -3 $R/opbench.err
