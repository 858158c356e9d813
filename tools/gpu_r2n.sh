#!/bin/bash
# round-2 GPU call N (1 GPU): full GPU suite on HEAD, default bench line, per-layer operand types (accuracy + speed)
mkdir -p gpurun_out
PLLB_C4_GOLDEN_MODES="fp16from:12,fp16from:6" PLLB_C2_GOLDEN_MODES="bf16+fp16head,fp16from:8,bf16+fp16tail,fp16" timeout 1200 python -m pytest tests -m gpu -x -q -s --durations=15 > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2n_bench_c2.json 2> gpurun_out/r2n_bench_c2.err
timeout 600 $B --operand-dtype bf16+fp16tail > gpurun_out/r2n_tail_a.json 2> gpurun_out/r2n_tail_a.err
timeout 600 $B --operand-dtype fp16from:8 > gpurun_out/r2n_from8_a.json 2> gpurun_out/r2n_from8_a.err
timeout 600 $B --operand-dtype bf16+fp16head > gpurun_out/r2n_head_a.json 2> gpurun_out/r2n_head_a.err
timeout 600 $B --operand-dtype fp16 > gpurun_out/r2n_fp16_a.json 2> gpurun_out/r2n_fp16_a.err
grep "golden\|passed\|failed\|rc=" gpurun_out/r2n_pytest.log | cut -c1-220
for f in gpurun_out/r2n_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d['dtype'], round(d['value'],1), round(d['ms_per_step'],1), d['clocks'])
except Exception as e: print(sys.argv[1], 'ERR', e)
P
done
