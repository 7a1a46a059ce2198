R=gpurun_out/r2o
mkdir -p $R
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > $R/bench_n2.json 2> $R/bench_n2.err; echo "exit $?"; cut -c1-300 $R/bench_n2.json; tail -5 $R/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $R/bench_ref_n2.json 2> $R/bench_ref_n2.err; echo "exit $?"; cut -c1-200 $R/bench_ref_n2.json; tail -3 $R/bench_ref_n2.err
python - $R <<'PY'
import json, sys
d = json.load(open(sys.argv[1] + '/bench_n2.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e'])
t = d.get('train'); print({k: t[k] for k in t if k != 'msda_kernels'} if t else None)
PY
