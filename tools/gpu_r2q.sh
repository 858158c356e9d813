#!/bin/bash
# round-2 GPU call Q (1 GPU): fine-tuning step after the row-sliced colsum; launch lists of the training step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -x > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
timeout 600 python tools/bench_train.py --batch 32 --steps 30 --warmup 3 --cpu-batches 0 > gpurun_out/r2q_train_b32.json 2> gpurun_out/r2q_train_b32.err
timeout 600 python tools/bench_train.py --batch 256 --steps 10 --warmup 2 --cpu-batches 0 > gpurun_out/r2q_train_b256.json 2> gpurun_out/r2q_train_b256.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/r2q_launches_train_b32.csv python tools/bench_train.py --batch 32 --steps 2 --warmup 1 --cpu-batches 0 > gpurun_out/r2q_ncu_b32.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/r2q_launches_train_b256.csv python tools/bench_train.py --batch 256 --steps 2 --warmup 1 --cpu-batches 0 > gpurun_out/r2q_ncu_b256.log 2>&1
tail -3 gpurun_out/r2q_pytest.log
cat gpurun_out/r2q_train_b32.json gpurun_out/r2q_train_b256.json
wc -l gpurun_out/r2q_launches_train_b32.csv gpurun_out/r2q_launches_train_b256.csv
