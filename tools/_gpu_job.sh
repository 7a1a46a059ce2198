R=gpurun_out/r2b
mkdir -p $R
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -30 $R/pytest_gpu.log
for regime in init; do
ncu --set full --clock-control none --import-source on -k regex:"msda_snippet_(fwd|bwd)_kernel" -s 12 -c 4 -o $R/ncu_presum_$regime python tools/opbench.py --iters 2 --warmup 1 --inner 1 --regime $regime --cases snip_enc_N1 > $R/ncu_$regime.log 2>&1
done
tail -3 $R/ncu_init.log
python tools/opbench.py --iters 20 --regime init --cases snip_enc_N1,snip_dec_N1 > $R/opbench.jsonl 2> $R/opbench.err; tail -3 $R/opbench.err
python - $R <<'PY'
import json, sys
for l in open(sys.argv[1] + '/opbench.jsonl'):
    d = json.loads(l)
    print("%-12s %-22s %-6s %-36s %9.2f us %7.1f GB/s %.4f" % (d['case'], d['impl'], d['regime'], d['pass'], d['us_median'], d['GBps'], d['frac_of_measured_hbm']))
PY
