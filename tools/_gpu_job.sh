R=gpurun_out/r2ag
mkdir -p $R
timeout 600 python -m pytest tests/test_bf16_gpu.py tests/test_msda_gpu.py tests/test_fuzz_gpu.py -m gpu -q -p timeout --timeout=180 > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -4 $R/pytest_gpu.log
timeout 300 python tools/opsweep.py --iters 20 --only dec_N1_k4_P8,dec_N1,q300_N1,dec_N64 > $R/opsweep_dec.jsonl 2> $R/opsweep.err; tail -2 $R/opsweep.err
python - $R <<'PY'
import json, sys
for l in open(sys.argv[1] + '/opsweep_dec.jsonl'):
    d = json.loads(l)
    if d['pass'] == 'fwd': print(d['config'], d['impl'], d['pass'], d['us'])
PY
