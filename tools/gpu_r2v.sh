#!/bin/bash
# round-2 GPU call V (1 GPU): forms of the K = H (attention-output) LayerNorm launch after the shared-memory parameters — ABAB
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
for r in a b; do
  timeout 600 $B > gpurun_out/r2v_default_$r.json 2> gpurun_out/r2v_default_$r.err
  PLLB_LN_STAGED=0 timeout 600 $B > gpurun_out/r2v_direct_$r.json 2> gpurun_out/r2v_direct_$r.err
  PLLB_LN_PAIR=2 timeout 600 $B > gpurun_out/r2v_pairstaged_$r.json 2> gpurun_out/r2v_pairstaged_$r.err
  PLLB_LN_STAGED=0 PLLB_LN_PAIR=2 timeout 600 $B > gpurun_out/r2v_pairdirect_$r.json 2> gpurun_out/r2v_pairdirect_$r.err
done
for f in gpurun_out/r2v_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); bk=d['roofline']['by_kind']
    print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],1), {k:round(v['ms'],1) for k,v in bk.items()}, d['clocks']['sm_mhz'], d['pll_checksum'])
except Exception as e: print(sys.argv[1], 'ERR', e)
P
done
