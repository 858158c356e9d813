#!/bin/bash
# round-2 GPU call R (1 GPU): fine-tuning step as a replayed CUDA graph
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -x -s > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
timeout 600 python tools/bench_train.py --batch 32 --steps 60 --warmup 40 --cpu-batches 0 > gpurun_out/r2r_train_b32.json 2> gpurun_out/r2r_train_b32.err
PLLB_TRAIN_GRAPH=0 timeout 600 python tools/bench_train.py --batch 32 --steps 60 --warmup 40 --cpu-batches 0 > gpurun_out/r2r_train_b32_eager.json 2> gpurun_out/r2r_train_b32_eager.err
timeout 600 python tools/bench_train.py --batch 256 --steps 30 --warmup 30 --cpu-batches 0 > gpurun_out/r2r_train_b256.json 2> gpurun_out/r2r_train_b256.err
PLLB_TRAIN_GRAPH=0 timeout 600 python tools/bench_train.py --batch 256 --steps 30 --warmup 30 --cpu-batches 0 > gpurun_out/r2r_train_b256_eager.json 2> gpurun_out/r2r_train_b256_eager.err
grep -v "^$" gpurun_out/r2r_pytest.log | tail -16 | cut -c1-220
for f in gpurun_out/r2r_train_*.json; do echo $f; python -c "
import json,sys; d=json.load(open('$f')); print({k: round(v,3) if isinstance(v,float) else v for k,v in d['gpu'].items()})"; done
tail -2 gpurun_out/r2r_train_b32.err
