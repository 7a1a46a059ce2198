#!/bin/bash
# run 19: single 16-byte sample record + chunk rotation -- parity and timing
mkdir -p gpurun_out/run19
R=gpurun_out/run19
timeout 900 python -m pytest tests -m gpu -x -q > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -4 $R/pytest_gpu.log
for regime in init local; do
timeout 300 python tools/opbench.py --iters 30 --regime $regime --bf16 --cases snip_enc_N1,snip_dec_N1,enc_N1,enc_N8,dec_N1 >> $R/opbench.jsonl 2>> $R/opbench.err
done
python - <<'PY'
import json
for l in open('gpurun_out/run19/opbench.jsonl'):
    d=json.loads(l)
    if d['pass'] in ('fwd','bwd'):
        print("%-12s %-22s %-6s %-6s %9.2f us %7.1f GB/s %.4f" % (d['case'],d['impl'],d['regime'],d['pass'],d['us_median'],d['GBps'],d['frac_of_measured_hbm']))
PY
