#!/bin/bash
# round-2 GPU call B: operand-type probe, full GPU test suite, first bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2b_gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "fp16_operands or gelu_epilogue or saturate" > gpurun_out/r2b_mixed.log 2>&1
echo "mixed probe rc=$?" >> gpurun_out/r2b_mixed.log
timeout 1500 python -m pytest tests -m gpu -x -q -s --deselect tests/test_gpu_parity.py::test_c2_2000_utterances_pll_and_one_best_vs_reference_golden > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
PLLB_C2_GOLDEN_MIN_UTTS=1000 timeout 900 python -m pytest tests/test_gpu_parity.py -q -s -k "c2_2000_utterances" > gpurun_out/r2b_c2golden.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err
timeout 600 python bench.py --steps 3 --warmup 2 --operand-dtype bf16+fp16head --no-cpu-baseline > gpurun_out/r2b_bench_c2_mixed.json 2> gpurun_out/r2b_bench_c2_mixed.err
timeout 600 python bench.py --steps 3 --warmup 2 --operand-dtype fp16 --no-cpu-baseline > gpurun_out/r2b_bench_c2_fp16.json 2> gpurun_out/r2b_bench_c2_fp16.err
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2b_bench_c2_b.json 2> gpurun_out/r2b_bench_c2_b.err
timeout 600 python bench.py --workload c1 --steps 5 --warmup 3 > gpurun_out/r2b_bench_c1.json 2> gpurun_out/r2b_bench_c1.err
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r2b_bench_c5.json 2> gpurun_out/r2b_bench_c5.err
tail -n 3 gpurun_out/r2b_mixed.log gpurun_out/r2b_pytest.log gpurun_out/r2b_c2golden.log
head -c 600 gpurun_out/r2b_bench_c2.json
