R=gpurun_out/r2n
mkdir -p $R
timeout 900 python -m pytest tests/test_msda_gpu.py tests/test_fuzz_gpu.py tests/test_bf16_gpu.py tests/test_capi_c_gpu.py tests/test_reference_dropin_gpu.py -m gpu -q --maxfail=40 -p timeout --timeout=180 > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -5 $R/pytest_gpu.log
for fl in "" "--flush"; do
  timeout 120 python tools/opbench.py --iters 30 --regime local $fl --ref --cases dec_N1,dec_N2 | sed "s/\"pairs\": 16/\"variant\": \"split\"/" >> $R/opbench_dec.jsonl 2>> $R/opbench.err
  MSDA_FWD_SPLIT=0 timeout 120 python tools/opbench.py --iters 30 --regime local $fl --cases dec_N1,dec_N2 | sed "s/\"pairs\": 16/\"variant\": \"tiles\"/" >> $R/opbench_dec.jsonl 2>> $R/opbench.err
done
tail -3 $R/opbench.err
python - $R <<'PY'
import json, sys
for l in open(sys.argv[1] + '/opbench_dec.jsonl'):
    d = json.loads(l)
    if d['pass'] == 'fwd': print("%-8s %-10s flush=%-5s %9.2f us  %s" % (d['case'], d['impl'], d['l2_flush'], d['us_median'], d.get('variant', '')))
PY
