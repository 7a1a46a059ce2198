R=gpurun_out/r2ac
mkdir -p $R
ob() { tag=$1; for regime in init local; do timeout 120 python tools/opbench.py --iters 30 --regime $regime --cases snip_enc_N1 --only bwd_presummed | grep -v deterministic | sed "s/\"pairs\": 16/\"variant\": \"$tag\"/" >> $R/opbench.jsonl 2>> $R/opbench.err; done; }
for thr in 768 1152 1344; do
  cd snipper_b200/csrc
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -DMSDA_BWD_PRESUM_THREADS=$thr -Xptxas=-v -c msda_snippet.cu -o ../lib/obj/msda_snippet.o 2>&1 | grep -A1 "bwd_kernelIfLi12ELi16ELi1536ELi2ELb1" | grep -E "Used|spill" | head -2
  nvcc -shared -Xcompiler -fPIC -o ../lib/libmsda_b200.so ../lib/obj/*.o
  cd ../..
  ob threads_$thr
done
python - $R <<'PY'
import json, sys
for l in open(sys.argv[1] + '/opbench.jsonl'):
    d = json.loads(l)
    if d['pass'] == 'bwd_presummed': print("%-6s %-20s %9.2f us %s" % (d['regime'], d['pass'], d['us_median'], d.get('variant', '')))
PY
