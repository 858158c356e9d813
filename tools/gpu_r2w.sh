#!/bin/bash
# round-2 GPU call W (1 GPU): FFN2 + LayerNorm with the staged (TMA box) 16-bit output instead of direct 32-byte stores — ABAB
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
for r in a b; do
  timeout 600 $B > gpurun_out/r2w_default_$r.json 2> gpurun_out/r2w_default_$r.err
  PLLB_LN_STAGED=1 timeout 600 $B > gpurun_out/r2w_staged_$r.json 2> gpurun_out/r2w_staged_$r.err
done
for f in gpurun_out/r2w_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); bk=d['roofline']['by_kind']
    print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],1), {k:round(v['ms'],1) for k,v in bk.items()}, d['clocks']['sm_mhz'], d['pll_checksum'])
except Exception as e: print(sys.argv[1], 'ERR', e)
P
done
