R=gpurun_out/r2z
mkdir -p $R
for i in 1 2; do timeout 900 python bench.py --steps 20 --warmup 5 2>> $R/bench.err | grep '^{' >> $R/bench_n1_x2.jsonl; done
python - $R <<'PY'
import json, sys
for l in open(sys.argv[1] + '/bench_n1_x2.jsonl'):
    d = json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['extra_untimed_warmup_steps'], d['train']['value'], d['gpu_baseline']['value'])
PY
