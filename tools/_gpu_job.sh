R=gpurun_out/r2am
mkdir -p $R
MSDA_FUZZ_SCALE=3 timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_fuzz_gpu.py -m gpu -q -p timeout --timeout=800 > $R/sanitizer_memcheck_fuzz_soak.log 2>&1; tail -4 $R/sanitizer_memcheck_fuzz_soak.log
