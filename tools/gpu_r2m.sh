#!/bin/bash
# round-2 GPU call M (1 GPU): per-layer operand types — accuracy on the full C2 golden and speed
mkdir -p gpurun_out
PLLB_C2_GOLDEN_MODES="bf16+fp16head,fp16from:9,fp16from:8,bf16+fp16tail,fp16from:4,fp16" timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "c2_2000_utterances or pll_vs_reference_golden or config4 or saturate" > gpurun_out/r2m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
timeout 600 $B --operand-dtype bf16+fp16head > gpurun_out/r2m_head_a.json 2> gpurun_out/r2m_head_a.err
timeout 600 $B --operand-dtype bf16+fp16tail > gpurun_out/r2m_tail_a.json 2> gpurun_out/r2m_tail_a.err
timeout 600 $B --operand-dtype fp16 > gpurun_out/r2m_fp16_a.json 2> gpurun_out/r2m_fp16_a.err
timeout 600 $B --operand-dtype fp16from:8 > gpurun_out/r2m_from8_a.json 2> gpurun_out/r2m_from8_a.err
timeout 600 $B --operand-dtype bf16+fp16head > gpurun_out/r2m_head_b.json 2> gpurun_out/r2m_head_b.err
timeout 600 $B --operand-dtype bf16+fp16tail > gpurun_out/r2m_tail_b.json 2> gpurun_out/r2m_tail_b.err
grep "c2 golden\|passed\|failed" gpurun_out/r2m_pytest.log | cut -c1-200
