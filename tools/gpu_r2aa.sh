#!/bin/bash
# round-2 GPU call AA (1 GPU): LayerNorm epilogue pass B with double-buffered accumulator chunks — parity + ABAB against the previous build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_gemm_layernorm or paired_layernorm or pll_vs_reference_golden or config4 or fp16_operand" > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aa_pytest.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
OLD=$PWD/asr-rescoring_b200/libpllb200_prev.so
for r in a b; do
  timeout 600 $B > gpurun_out/r2aa_new_$r.json 2> gpurun_out/r2aa_new_$r.err
  PLLB_LIB=$OLD timeout 600 $B > gpurun_out/r2aa_old_$r.json 2> gpurun_out/r2aa_old_$r.err
done
tail -3 gpurun_out/r2aa_pytest.log
for f in gpurun_out/r2aa_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); bk=d['roofline']['by_kind']
    print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],1), {k:round(v['ms'],1) for k,v in bk.items()}, d['clocks']['sm_mhz'], d['pll_checksum'])
except Exception as e: print(sys.argv[1], 'ERR', e)
P
done
