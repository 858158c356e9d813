#!/bin/bash
# round-2 GPU call Y (8 GPUs): config 3 (the C2 list sharded over 8 ranks, strong scaling) with the final build
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577"
timeout 600 $TR bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2y_bench_c2_8gpu.json 2> gpurun_out/r2y_bench_c2_8gpu.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2y_bench_c2_8gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['scaling'], round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['config']['score_call_ms_by_rank'], d['config']['lpt_load_imbalance'], d['pll_checksum'], d['best_weight'], d['best_cer'], d['clocks'])
P
