#!/bin/bash
# round-2 GPU call U (1 GPU): L2 prefetch of the next row block's A boxes in the fused GEMM+LayerNorm kernels — ABAB
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
for r in a b; do
  PLLB_LN_APF=1 timeout 600 $B > gpurun_out/r2u_apf1_$r.json 2> gpurun_out/r2u_apf1_$r.err
  PLLB_LN_APF=0 timeout 600 $B > gpurun_out/r2u_apf0_$r.json 2> gpurun_out/r2u_apf0_$r.err
  PLLB_LN_APF=2 timeout 600 $B > gpurun_out/r2u_apf2_$r.json 2> gpurun_out/r2u_apf2_$r.err
done
for f in gpurun_out/r2u_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); bk=d['roofline']['by_kind']
    print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],1), {k:round(v['ms'],1) for k,v in bk.items()}, d['clocks']['sm_mhz'], d['pll_checksum'])
except Exception as e: print(sys.argv[1], 'ERR', e)
P
done
