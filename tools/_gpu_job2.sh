R=gpurun_out/r2aj
mkdir -p $R
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 5 > $R/bench_n4.json 2> $R/bench_n4.err; echo "exit $?"
python - $R <<'PY'
import json, sys
lines = [l for l in open(sys.argv[1] + '/bench_n4.json') if l.startswith('{')]
d = json.loads(lines[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e'], d['clocks'])
t = d.get('train'); print({k: t[k] for k in t if k != 'msda_kernels'} if t else None)
PY
