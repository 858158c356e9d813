#!/bin/bash
# round-2 GPU call GG (1 GPU): chunk size of the scoring pipeline (expanded rows per chunk) — 2^19 / 2^20 (default) / 2^21
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
timeout 600 $B > gpurun_out/r2gg_c20_a.json 2> gpurun_out/r2gg_c20_a.err
timeout 600 $B --chunk-tokens 2097152 > gpurun_out/r2gg_c21_a.json 2> gpurun_out/r2gg_c21_a.err
timeout 600 $B --chunk-tokens 524288 > gpurun_out/r2gg_c19_a.json 2> gpurun_out/r2gg_c19_a.err
timeout 600 $B > gpurun_out/r2gg_c20_b.json 2> gpurun_out/r2gg_c20_b.err
timeout 600 $B --chunk-tokens 2097152 > gpurun_out/r2gg_c21_b.json 2> gpurun_out/r2gg_c21_b.err
for f in gpurun_out/r2gg_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); bk=d['roofline']['by_kind']
    print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],1), {k:round(v['ms'],1) for k,v in bk.items()}, d['clocks']['sm_mhz'], d['pll_checksum'])
except Exception as e: print(sys.argv[1], 'ERR', e)
P
done
