R=gpurun_out/r2s
mkdir -p $R
timeout 600 python -m pytest tests/test_fullsize_gpu.py tests/test_fuzz_gpu.py -m gpu -q -k "planar" --maxfail=10 -p timeout --timeout=120 > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -12 $R/pytest_gpu.log
for regime in init local; do
  timeout 120 python tools/opbench.py --iters 30 --regime $regime --cases snip_enc_N1 --only fwd_planar,"analytic ref" >> $R/opbench_win.jsonl 2>> $R/opbench.err
done
tail -3 $R/opbench.err
python - $R <<'PY'
import json, sys
for l in open(sys.argv[1] + '/opbench_win.jsonl'):
    d = json.loads(l)
    print("%-6s %-40s %9.2f us %7.1f GB/s %.4f" % (d['regime'], d['pass'], d['us_median'], d['GBps'], d['frac_of_measured_hbm']))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"winfwd" -s 2 -c 1 -o $R/ncu_op_winfwd_init python tools/opbench.py --iters 2 --warmup 1 --inner 1 --regime init --cases snip_enc_N1 --only windowed > $R/ncu_op_win.log 2>&1; tail -2 $R/ncu_op_win.log
