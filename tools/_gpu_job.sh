R=gpurun_out/r2c
mkdir -p $R
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 > $R/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $R/pytest_gpu.log; tail -25 $R/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $R/bench_n1.json 2> $R/bench_n1.err; cat $R/bench_n1.json | cut -c1-1500; tail -3 $R/bench_n1.err
timeout 600 python tools/e2e_profile.py --steps 3 > $R/e2e_profile.json 2> $R/e2e_profile.err; cut -c1-3000 $R/e2e_profile.json
