#!/bin/bash
# first GPU session: parity tests, op-level bench vs vendored op, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python tools/opbench.py --ref --iters 50 > gpurun_out/opbench_noflush.jsonl 2> gpurun_out/opbench.err
timeout 600 python tools/opbench.py --ref --iters 30 --flush > gpurun_out/opbench_flush.jsonl 2>> gpurun_out/opbench.err
timeout 600 python tools/opbench.py --ref --iters 30 --regime uniform > gpurun_out/opbench_uniform.jsonl 2>> gpurun_out/opbench.err
cat gpurun_out/opbench_noflush.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_opbench.csv python tools/opbench.py --iters 2 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 14 -c 4 -o gpurun_out/prof_percall python tools/opbench.py --iters 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
