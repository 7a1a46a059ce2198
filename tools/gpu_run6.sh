#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 9 -c 1 -o gpurun_out/prof_run6_fwd python tools/opbench.py --iters 1 --warmup 0 --regime init --cases enc_N1 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_bwd -s 9 -c 1 -o gpurun_out/prof_run6_bwd python tools/opbench.py --iters 1 --warmup 0 --regime init --cases enc_N1 >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
