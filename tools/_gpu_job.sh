R=gpurun_out/r2j
mkdir -p $R
timeout 300 python tools/debug/planar_probe.py multi_level_fwd multi_level_bwd encoder_fwd encoder_bwd > $R/probe.log 2>&1; cat $R/probe.log
ob() { # tag, extra env...
  tag=$1; shift
  for regime in init local; do
    env "$@" timeout 120 python tools/opbench.py --iters 30 --regime $regime --cases snip_enc_N1 --only fwd_planar,bwd_planar | sed "s/\"pairs\": 16/\"variant\": \"$tag\"/" >> $R/opbench_variants.jsonl 2>> $R/opbench.err
  done
}
ob tile2d_p32 A=1
ob rows_p32 MSDA_PLANAR_TILE2D=0
ob tile2d_p16 MSDA_PLANAR_PAIRS=16
ob tile2d_p64 MSDA_PLANAR_PAIRS=64
tools/debug/variant.sh -DMSDA_PLANAR_FWD_MIN_BLOCKS=5 -DMSDA_PLANAR_BWD_MIN_BLOCKS=5
ob tile2d_p32_blocks5 A=1
ob tile2d_p16_blocks5 MSDA_PLANAR_PAIRS=16
tools/debug/variant.sh -DMSDA_PLANAR_FWD_MIN_BLOCKS=6 -DMSDA_PLANAR_BWD_MIN_BLOCKS=6
ob tile2d_p32_blocks6 A=1
ob tile2d_p16_blocks6 MSDA_PLANAR_PAIRS=16
tools/debug/variant.sh -DMSDA_PLANAR_FWD_MIN_BLOCKS=8 -DMSDA_PLANAR_BWD_MIN_BLOCKS=4
ob tile2d_p32_blocks8 A=1
tail -5 $R/opbench.err
python - $R <<'PY'
import json, sys
for l in open(sys.argv[1] + '/opbench_variants.jsonl'):
    d = json.loads(l)
    if d['pass'] in ('fwd_planar', 'bwd_planar'):
        print("%-6s %-12s %9.2f us %7.1f GB/s %.4f %s" % (d['regime'], d['pass'], d['us_median'], d['GBps'], d['frac_of_measured_hbm'], d.get('variant', '')))
PY
