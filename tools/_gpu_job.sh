R=gpurun_out/r2ai
mkdir -p $R
timeout 600 python tools/e2e_profile.py --steps 3 > $R/e2e_profile_infer.json 2> $R/e2e.err; cut -c1-300 $R/e2e_profile_infer.json
timeout 600 python tools/e2e_profile.py --steps 2 --train > $R/e2e_profile_train.json 2>> $R/e2e.err; cut -c1-300 $R/e2e_profile_train.json; tail -2 $R/e2e.err
