R=gpurun_out/r2b
mkdir -p $R
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -40 $R/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $R/smoke.log 2>&1; tail -4 $R/smoke.log
python tools/opbench.py --iters 20 --regime init --cases snip_enc_N1,snip_dec_N1 > $R/opbench.jsonl 2> $R/opbench.err; tail -3 $R/opbench.err
python tools/opbench.py --iters 20 --regime local --cases snip_enc_N1 >> $R/opbench.jsonl 2>> $R/opbench.err
python - $R <<'PY'
import json, sys
for l in open(sys.argv[1] + '/opbench.jsonl'):
    d = json.loads(l)
    print("%-12s %-22s %-6s %-36s %9.2f us %7.1f GB/s %.4f" % (d['case'], d['impl'], d['regime'], d['pass'], d['us_median'], d['GBps'], d['frac_of_measured_hbm']))
PY
# matching launches: [0-2] fwd_direct [3-5] bwd_direct [6-8] fwd_presummed [9-11] layer_fwd [12-14] bwd_presummed ...
ncu --set full --clock-control none --import-source on -k regex:"msda_snippet_(fwd|bwd)_kernel" -s 7 -c 1 -o $R/ncu_fwd_presum_init python tools/opbench.py --iters 2 --warmup 1 --inner 1 --regime init --cases snip_enc_N1 > $R/ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"msda_snippet_(fwd|bwd)_kernel" -s 13 -c 1 -o $R/ncu_bwd_presum_init python tools/opbench.py --iters 2 --warmup 1 --inner 1 --regime init --cases snip_enc_N1 > $R/ncu_bwd.log 2>&1
tail -2 $R/ncu_fwd.log $R/ncu_bwd.log
timeout 900 python bench.py --steps 20 --warmup 5 > $R/bench_n1.json 2> $R/bench_n1.err; cat $R/bench_n1.json; tail -5 $R/bench_n1.err
