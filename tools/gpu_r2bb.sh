#!/bin/bash
# round-2 GPU call BB (1 GPU): the FFN intermediate in the K-blocked layout — parity, then ABAB against the row-major form
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "not config1_100x10 and not combiner and not levenshtein and not text_front_end" > gpurun_out/r2bb_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2bb_pytest.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
for r in a b; do
  timeout 600 $B > gpurun_out/r2bb_blocked_$r.json 2> gpurun_out/r2bb_blocked_$r.err
  PLLB_FFN_BLOCKED=0 timeout 600 $B > gpurun_out/r2bb_rowmajor_$r.json 2> gpurun_out/r2bb_rowmajor_$r.err
done
tail -4 gpurun_out/r2bb_pytest.log
for f in gpurun_out/r2bb_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); bk=d['roofline']['by_kind']
    print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],1), {k:round(v['ms'],1) for k,v in bk.items()}, d['clocks']['sm_mhz'], d['pll_checksum'])
except Exception as e: print(sys.argv[1], 'ERR', e)
P
done
tail -3 gpurun_out/r2bb_blocked_a.err
