R=gpurun_out/r2t
mkdir -p $R
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $R/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p timeout --timeout=180 > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -4 $R/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $R/smoke.log 2>&1; tail -2 $R/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $R/bench_n1.json 2> $R/bench_n1.err; cut -c1-300 $R/bench_n1.json; tail -2 $R/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $R/bench_reference_arm.json 2> $R/bench_reference_arm.err; cut -c1-200 $R/bench_reference_arm.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"msda_snippet_fwd_kernel|frame_sum_kernel|fwd_split" -s 4 -c 3 -o $R/ncu_bench_fwd python bench.py --steps 1 --warmup 1 --no-graph --no-train --no-gpu-baseline --no-cpu-baseline > $R/ncu_bench_fwd.log 2>&1; tail -2 $R/ncu_bench_fwd.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"msda_snippet_bwd_kernel|frame_unsum_kernel" -s 8 -c 2 -o $R/ncu_op_bwd_init python tools/opbench.py --iters 2 --warmup 1 --inner 1 --regime init --cases snip_enc_N1 --only presummed,frame_unsum > $R/ncu_op_bwd.log 2>&1
rm -f $R/opbench.jsonl $R/opbench_flushed.jsonl
for regime in init local; do
  timeout 300 python tools/opbench.py --iters 30 --regime $regime --bf16 --ref --cases snip_enc_N1,snip_dec_N1,enc_N1,enc_N8,dec_N1 >> $R/opbench.jsonl 2>> $R/opbench.err
  timeout 300 python tools/opbench.py --iters 30 --regime $regime --flush --ref --cases snip_enc_N1,snip_dec_N1,enc_N1,dec_N1 >> $R/opbench_flushed.jsonl 2>> $R/opbench.err
done
tail -3 $R/opbench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3400 --csv --log-file $R/launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-train --no-gpu-baseline --no-cpu-baseline > $R/ncu_list.log 2>&1; tail -2 $R/ncu_list.log; wc -l $R/launches.csv
du -sh $R
