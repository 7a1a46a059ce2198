R=gpurun_out/r2an
mkdir -p $R
timeout 900 python bench.py > $R/bench_default.json 2> $R/bench_default.err; python - $R <<'PY'
import json, sys
d = json.load(open(sys.argv[1] + '/bench_default.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline_backward']['frac'], d['roofline_backward']['traffic'])
print(d['train']['value'], d['train']['ms_per_step'], d['gpu_baseline']['value'], d['cpu_baseline']['value'], d['config']['extra_untimed_warmup_steps'])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $R/bench_ref.json 2>> $R/bench_default.err; cut -c1-160 $R/bench_ref.json
