R=gpurun_out/r2ae
mkdir -p $R
timeout 600 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -k "decoder" -p timeout --timeout=180 > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -4 $R/pytest_gpu.log
