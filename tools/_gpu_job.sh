R=gpurun_out/r2f
mkdir -p $R
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $R/smi.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -12 $R/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $R/smoke.log 2>&1; tail -2 $R/smoke.log
./tools/micro/gather_shapes > $R/gather_shapes.json 2>&1; cat $R/gather_shapes.json
rm -f $R/opbench.jsonl
for regime in init local; do
  timeout 300 python tools/opbench.py --iters 30 --regime $regime --cases snip_enc_N1,snip_dec_N1 >> $R/opbench.jsonl 2>> $R/opbench.err
done
for pairs in 16 64; do
  MSDA_PLANAR_PAIRS=$pairs timeout 300 python tools/opbench.py --iters 30 --regime init --cases snip_enc_N1 --only fwd_planar,bwd_planar | sed "s/\"pairs\": 16/\"planar_pairs\": $pairs/" >> $R/opbench_planar_pairs.jsonl 2>> $R/opbench.err
  MSDA_PLANAR_PAIRS=$pairs timeout 300 python tools/opbench.py --iters 30 --regime local --cases snip_enc_N1 --only fwd_planar,bwd_planar | sed "s/\"pairs\": 16/\"planar_pairs\": $pairs/" >> $R/opbench_planar_pairs.jsonl 2>> $R/opbench.err
done
tail -5 $R/opbench.err
python - $R <<'PY'
import json, sys
for f in ('/opbench.jsonl', '/opbench_planar_pairs.jsonl'):
  for l in open(sys.argv[1] + f):
    d = json.loads(l)
    print("%-12s %-6s %-36s %9.2f us %7.1f GB/s %.4f %s" % (d['case'], d['regime'], d['pass'], d['us_median'], d['GBps'], d['frac_of_measured_hbm'], d.get('planar_pairs', '')))
PY
timeout 900 python bench.py --steps 20 --warmup 5 > $R/bench_n1.json 2> $R/bench_n1.err; cut -c1-400 $R/bench_n1.json; tail -3 $R/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-train --no-gpu-baseline --no-cpu-baseline > $R/ncu_list.log 2>&1; tail -2 $R/ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"msda_planar_fwd_kernel" -s 2 -c 1 -o $R/ncu_bench_planar_fwd python bench.py --steps 1 --warmup 1 --no-graph --no-train --no-gpu-baseline --no-cpu-baseline > $R/ncu_bench_fwd.log 2>&1; tail -2 $R/ncu_bench_fwd.log
for k in msda_planar_fwd_kernel msda_planar_bwd_kernel frame_sum_planar_kernel frame_unsum_planar_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 3 -c 1 -o $R/ncu_op_init_$k python tools/opbench.py --iters 2 --warmup 1 --inner 1 --regime init --cases snip_enc_N1 --only planar > $R/ncu_op_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"msda_planar_(fwd|bwd)_kernel" -s 3 -c 1 -o $R/ncu_op_local_planar_fwd python tools/opbench.py --iters 2 --warmup 1 --inner 1 --regime local --cases snip_enc_N1 --only fwd_planar > $R/ncu_op_local_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"msda_planar_(fwd|bwd)_kernel" -s 3 -c 1 -o $R/ncu_op_local_planar_bwd python tools/opbench.py --iters 2 --warmup 1 --inner 1 --regime local --cases snip_enc_N1 --only bwd_planar > $R/ncu_op_local_bwd.log 2>&1
du -sh $R; ls -la $R
