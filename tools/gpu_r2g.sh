#!/bin/bash
# round-2 GPU call G (1 GPU): paired LayerNorm clusters — tests and ABAB
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "paired or cluster_sizes or pll_vs_reference_golden or layers_vs_oracle or rescorebert or determin" > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
PLLB_LN_PAIR=0 timeout 600 $B > gpurun_out/r2g_pair0_a.json 2> gpurun_out/r2g_pair0_a.err
PLLB_LN_PAIR=1 timeout 600 $B > gpurun_out/r2g_pair1_a.json 2> gpurun_out/r2g_pair1_a.err
PLLB_LN_PAIR=2 timeout 600 $B > gpurun_out/r2g_pair2_a.json 2> gpurun_out/r2g_pair2_a.err
PLLB_LN_PAIR=0 timeout 600 $B > gpurun_out/r2g_pair0_b.json 2> gpurun_out/r2g_pair0_b.err
PLLB_LN_PAIR=1 timeout 600 $B > gpurun_out/r2g_pair1_b.json 2> gpurun_out/r2g_pair1_b.err
tail -n 5 gpurun_out/r2g_pytest.log
