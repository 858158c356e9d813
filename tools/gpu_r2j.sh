#!/bin/bash
# round-2 GPU call J (1 GPU): early accumulator release in the 16-bit GEMM epilogues, FFN1 in cta_group::2 mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "gemm or gelu or pll_vs_reference_golden or determin or layers_vs_oracle" > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
PLLB_GEMM_EARLY_RELEASE=0 timeout 600 $B > gpurun_out/r2j_late_a.json 2> gpurun_out/r2j_late_a.err
timeout 600 $B > gpurun_out/r2j_early_a.json 2> gpurun_out/r2j_early_a.err
PLLB_GEMM_MODE_GELU=2 timeout 600 $B > gpurun_out/r2j_early_gelu2_a.json 2> gpurun_out/r2j_early_gelu2_a.err
PLLB_GEMM_EARLY_RELEASE=0 timeout 600 $B > gpurun_out/r2j_late_b.json 2> gpurun_out/r2j_late_b.err
timeout 600 $B > gpurun_out/r2j_early_b.json 2> gpurun_out/r2j_early_b.err
PLLB_GEMM_MODE_GELU=2 timeout 600 $B > gpurun_out/r2j_early_gelu2_b.json 2> gpurun_out/r2j_early_gelu2_b.err
tail -n 3 gpurun_out/r2j_pytest.log
