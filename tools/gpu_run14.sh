#!/bin/bash
# run 14: config 4 (training step) single GPU fp32 + bf16 autocast; config 5 (op sweep vs vendored op and CPU)
mkdir -p gpurun_out/run14
R=gpurun_out/run14
timeout 600 python tools/trainbench.py --steps 8 --warmup 3 > $R/train_n1_fp32.json 2> $R/train.err; cut -c1-600 $R/train_n1_fp32.json
timeout 600 python tools/trainbench.py --steps 8 --warmup 3 --bf16 > $R/train_n1_bf16.json 2>> $R/train.err; cut -c1-600 $R/train_n1_bf16.json
tail -5 $R/train.err
timeout 1500 python tools/opsweep.py --iters 15 --ref --cpu > $R/opsweep.jsonl 2> $R/opsweep.err
python - <<'PY'
import json
for l in open('gpurun_out/run14/opsweep.jsonl'):
    d=json.loads(l)
    print("%-14s N=%-3d Lq=%-5d L=%-2d P=%d %-22s %-4s %11.2f us %8.1f GB/s %s" % (d['config'],d['N'],d['Lq'],d['levels'],d['P'],d['impl'],d['pass'],d['us'],d['GBps'],d.get('frac_of_measured_hbm','')))
PY
tail -5 $R/opsweep.err
