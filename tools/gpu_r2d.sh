#!/bin/bash
# round-2 GPU call D (N GPUs): strong-scaling bench lines and the N-GPU == 1-GPU identity log
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r2d_bench_c2_${N}gpu.json 2> gpurun_out/r2d_bench_c2_${N}gpu.err
timeout 600 $TR bench.py --gpus $N --workload c5 --steps 5 --warmup 2 > gpurun_out/r2d_bench_c5_${N}gpu.json 2> gpurun_out/r2d_bench_c5_${N}gpu.err
timeout 900 python tools/e2e_dropin.py --gpus $N > gpurun_out/r2d_identity_${N}gpu.log 2>&1
echo "identity rc=$?" >> gpurun_out/r2d_identity_${N}gpu.log
tail -n 8 gpurun_out/r2d_identity_${N}gpu.log
head -c 300 gpurun_out/r2d_bench_c2_${N}gpu.json
