R=gpurun_out/r2al
mkdir -p $R
timeout 1500 python -m pytest tests -m gpu -x -q -p timeout --timeout=180 > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -3 $R/pytest_gpu.log
MSDA_FUZZ_SCALE=8 timeout 900 python -m pytest tests/test_fuzz_gpu.py -m gpu -q -p timeout --timeout=180 > $R/pytest_fuzz_soak.log 2>&1; echo "pytest exit $?" >> $R/pytest_fuzz_soak.log; tail -3 $R/pytest_fuzz_soak.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-gpu-baseline --no-cpu-baseline 2> $R/bench.err | grep '^{' > $R/bench.json; python - $R <<'PY'
import json, sys
d = json.load(open(sys.argv[1] + '/bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'])
PY
