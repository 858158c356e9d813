#!/bin/bash
# round-2 GPU call O (1 GPU): MLM fine-tuning path — per-tensor gradient report, then its GPU tests
mkdir -p gpurun_out
timeout 300 python tools/train_probe.py tiny > gpurun_out/r2o_probe_tiny.log 2>&1; echo "probe rc=$?" >> gpurun_out/r2o_probe_tiny.log
timeout 600 python -m pytest tests/test_gpu_train.py -q -s > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -60 gpurun_out/r2o_probe_tiny.log
grep -v "^$" gpurun_out/r2o_pytest.log | tail -60 | cut -c1-250
