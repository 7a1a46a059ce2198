R=gpurun_out/r2ab
mkdir -p $R
ob() { tag=$1; for regime in init local; do timeout 120 python tools/opbench.py --iters 30 --regime $regime --cases snip_enc_N1,enc_N1 --only presummed | grep -v deterministic | sed "s/\"pairs\": 16/\"variant\": \"$tag\"/" >> $R/opbench.jsonl 2>> $R/opbench.err; done; }
ob default
MSDA_NVCC_EXTRA="-DMSDA_LOAD_EVICT_LAST" python -c "from snipper_b200 import build; build.build_library(force=True)" > $R/build.log 2>&1; tail -1 $R/build.log
ob evict_last
timeout 300 ncu --set full --clock-control none -k regex:"msda_snippet_bwd_kernel" -s 1 -c 1 -o $R/ncu_bwd_evict_last python tools/opbench.py --iters 1 --warmup 1 --inner 1 --regime init --cases snip_enc_N1 --only bwd_presummed > $R/ncu.log 2>&1
python - $R <<'PY'
import json, sys
for l in open(sys.argv[1] + '/opbench.jsonl'):
    d = json.loads(l)
    print("%-12s %-6s %-40s %9.2f us %s" % (d['case'], d['regime'], d['pass'], d['us_median'], d.get('variant', '')))
PY
