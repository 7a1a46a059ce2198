#!/bin/bash
# run 12: bf16 backward with paired-lane reductions; compute-sanitizer memcheck + racecheck on the op tests
mkdir -p gpurun_out/run12
R=gpurun_out/run12
timeout 900 python -m pytest tests -m gpu -x -q > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -4 $R/pytest_gpu.log
timeout 300 python tools/opbench.py --iters 30 --regime local --bf16 --cases snip_enc_N1,snip_dec_N1,enc_N1,dec_N1 >> $R/opbench.jsonl 2>> $R/opbench.err
python - <<'PY'
import json
for l in open('gpurun_out/run12/opbench.jsonl'):
    d=json.loads(l)
    print("%-12s %-22s %-6s %-18s %9.2f us %7.1f GB/s %.4f" % (d['case'],d['impl'],d['regime'],d['pass'],d['us_median'],d['GBps'],d['frac_of_measured_hbm']))
PY
SEL='not full_size and not large_channels and not gradcheck and not vendored and not golden_fp64 and not deterministic_full'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_bf16_gpu.py tests/test_msda_gpu.py tests/test_module_gpu.py -m gpu -x -q -k "$SEL" > $R/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> $R/sanitizer_memcheck.log; tail -6 $R/sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_bf16_gpu.py tests/test_msda_gpu.py -m gpu -x -q -k "fast_path_channels or bf16_percall or snipper_golden_fp32 or deterministic_backward or fused_snippet or pile_up" > $R/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> $R/sanitizer_racecheck.log; tail -6 $R/sanitizer_racecheck.log
