R=gpurun_out/r2r
mkdir -p $R
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_fuzz_gpu.py tests/test_msda_gpu.py tests/test_fullsize_gpu.py -m gpu -q -k "planar or packed or levels_points or decoder or ragged or config3" -p timeout --timeout=800 > $R/sanitizer_racecheck.log 2>&1; tail -5 $R/sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_fuzz_gpu.py tests/test_msda_gpu.py tests/test_layers_gpu.py tests/test_module_gpu.py -m gpu -q -p timeout --timeout=800 > $R/sanitizer_memcheck.log 2>&1; tail -5 $R/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool synccheck python -m pytest tests/test_fuzz_gpu.py -m gpu -q -k "planar or packed" -p timeout --timeout=500 > $R/sanitizer_synccheck.log 2>&1; tail -4 $R/sanitizer_synccheck.log
