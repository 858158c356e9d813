#!/bin/bash
# round-2 GPU call A: operand-type probe, full GPU test suite, first bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "operand_types or gelu_epilogue" > gpurun_out/r2a_mixed.log 2>&1
echo "mixed probe rc=$?" >> gpurun_out/r2a_mixed.log
timeout 1500 python -m pytest tests -m gpu -x -q -s --deselect tests/test_gpu_parity.py::test_c2_2000_utterances_pll_and_one_best_vs_reference_golden > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
timeout 600 python bench.py --steps 3 --warmup 2 --operand-dtype mixed --no-cpu-baseline > gpurun_out/r2a_bench_c2_mixed.json 2> gpurun_out/r2a_bench_c2_mixed.err
timeout 600 python bench.py --steps 3 --warmup 2 --operand-dtype fp16 --no-cpu-baseline > gpurun_out/r2a_bench_c2_fp16.json 2> gpurun_out/r2a_bench_c2_fp16.err
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2a_bench_c2_b.json 2> gpurun_out/r2a_bench_c2_b.err
timeout 600 python bench.py --workload c1 --steps 5 --warmup 3 > gpurun_out/r2a_bench_c1.json 2> gpurun_out/r2a_bench_c1.err
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r2a_bench_c5.json 2> gpurun_out/r2a_bench_c5.err
tail -3 gpurun_out/r2a_mixed.log gpurun_out/r2a_pytest.log
head -c 600 gpurun_out/r2a_bench_c2.json
