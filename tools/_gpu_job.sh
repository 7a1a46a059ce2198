R=gpurun_out/r2final
mkdir -p $R
timeout 1500 python -m pytest tests -m gpu -x -q -p timeout --timeout=180 > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -3 $R/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $R/smoke.log 2>&1; tail -2 $R/smoke.log
timeout 600 python bench.py > $R/bench_default.json 2> $R/bench_default.err; python - $R <<'PY'
import json, sys
d = json.load(open(sys.argv[1] + '/bench_default.json'))
print({k: d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','higher_is_better','scaling','vs_baseline','dtype','data','gpu_launches')})
print(d['e2e'], d['clocks']); print({k: d['roofline'][k] for k in ('bound','achieved','peak','unit','frac','traffic')}); print(d['cpu_baseline'])
PY
