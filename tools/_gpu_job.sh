R=gpurun_out/r2m
mkdir -p $R
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p timeout --timeout=180 > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -6 $R/pytest_gpu.log
for regime in init local; do
  timeout 120 python tools/opbench.py --iters 30 --regime $regime --cases snip_dec_N1 --only fwd_direct | sed "s/\"pairs\": 16/\"variant\": \"split\"/" >> $R/opbench_dec.jsonl 2>> $R/opbench.err
  MSDA_FWD_SPLIT=0 timeout 120 python tools/opbench.py --iters 30 --regime $regime --cases snip_dec_N1 --only fwd_direct | sed "s/\"pairs\": 16/\"variant\": \"tiles\"/" >> $R/opbench_dec.jsonl 2>> $R/opbench.err
done
tail -3 $R/opbench.err
python - $R <<'PY'
import json, sys
for l in open(sys.argv[1] + '/opbench_dec.jsonl'):
    d = json.loads(l)
    print("%-6s %-40s %9.2f us %7.1f GB/s %.4f %s" % (d['regime'], d['pass'], d['us_median'], d['GBps'], d['frac_of_measured_hbm'], d.get('variant', '')))
PY
timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-gpu-baseline --no-cpu-baseline > $R/bench_n1.json 2> $R/bench_n1.err; cut -c1-300 $R/bench_n1.json; tail -2 $R/bench_n1.err
python - $R <<'PY'
import json, sys
d = json.load(open(sys.argv[1] + '/bench_n1.json'))
for k in d['roofline']['kernels']: print(k)
PY
