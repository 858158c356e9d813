#!/bin/bash
# round-2 GPU call DD (2 GPUs): N GPUs == 1 GPU identity test and the 2-GPU strong-scaling line on the final build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -s -k "two_gpu" > gpurun_out/r2dd_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2dd_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29588"
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2dd_bench_c2_2gpu.json 2> gpurun_out/r2dd_bench_c2_2gpu.err
tail -3 gpurun_out/r2dd_pytest.log
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2dd_bench_c2_2gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['scaling'], round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['config']['score_call_ms_by_rank'], d['pll_checksum'], d['best_weight'], d['best_cer'])
P
