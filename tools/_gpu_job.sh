R=gpurun_out/r2x
mkdir -p $R
timeout 300 python tools/debug/replay_modes.py > $R/replay_modes.json 2> $R/replay_modes.err; cat $R/replay_modes.json; tail -2 $R/replay_modes.err
