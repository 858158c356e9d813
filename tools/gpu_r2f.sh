#!/bin/bash
# round-2 GPU call F (8 GPUs): config 3 (c2 sharded, strong scaling) and config 4 (c4 on 8 GPUs, 150 utterances per GPU)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r2f_bench_c2_${N}gpu.json 2> gpurun_out/r2f_bench_c2_${N}gpu.err
timeout 900 $TR bench.py --gpus $N --workload c4 --utts $((150 * N)) --steps 1 --warmup 1 > gpurun_out/r2f_bench_c4_${N}gpu.json 2> gpurun_out/r2f_bench_c4_${N}gpu.err
head -c 400 gpurun_out/r2f_bench_c2_${N}gpu.json; echo; head -c 400 gpurun_out/r2f_bench_c4_${N}gpu.json
