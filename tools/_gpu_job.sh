R=gpurun_out/r2u
mkdir -p $R
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_model_gpu.py -m gpu -q -p timeout --timeout=180 > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -4 $R/pytest_gpu.log
for i in 1 2; do timeout 900 python bench.py --steps 40 --warmup 5 --no-train --no-gpu-baseline --no-cpu-baseline 2>> $R/bench.err | grep '^{' >> $R/bench_infer_40steps.jsonl; done
python - $R <<'PY'
import json, sys
for l in open(sys.argv[1] + '/bench_infer_40steps.jsonl'):
    d = json.loads(l); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['traffic'], d['roofline']['frac'])
PY
