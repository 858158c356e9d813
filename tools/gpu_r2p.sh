#!/bin/bash
# round-2 GPU call P (1 GPU): fine-tuning path at the bert-base shape, its tests, throughput next to the reference loop
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -s > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
timeout 300 python tools/train_probe.py tiny > gpurun_out/r2p_probe_tiny.log 2>&1
timeout 900 python tools/bench_train.py --batch 32 --steps 30 --warmup 3 --cpu-batches 2 > gpurun_out/r2p_train_b32.json 2> gpurun_out/r2p_train_b32.err
timeout 600 python tools/bench_train.py --batch 256 --steps 10 --warmup 2 --cpu-batches 0 > gpurun_out/r2p_train_b256.json 2> gpurun_out/r2p_train_b256.err
grep -v "^$" gpurun_out/r2p_pytest.log | tail -25 | cut -c1-250
head -3 gpurun_out/r2p_probe_tiny.log; grep -c "<<<<" gpurun_out/r2p_probe_tiny.log
cat gpurun_out/r2p_train_b32.json gpurun_out/r2p_train_b256.json; tail -3 gpurun_out/r2p_train_b32.err
