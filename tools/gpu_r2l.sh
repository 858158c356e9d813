#!/bin/bash
# round-2 GPU call L (1 GPU): TMA-fed attention with the byte-granular ring (T <= 32): bit-identity, ABAB, launch times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tma or attention" > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
timeout 600 $B > gpurun_out/r2l_att0_a.json 2> gpurun_out/r2l_att0_a.err
PLLB_ATT_TMA=1 timeout 600 $B > gpurun_out/r2l_att1_a.json 2> gpurun_out/r2l_att1_a.err
timeout 600 $B > gpurun_out/r2l_att0_b.json 2> gpurun_out/r2l_att0_b.err
PLLB_ATT_TMA=1 timeout 600 $B > gpurun_out/r2l_att1_b.json 2> gpurun_out/r2l_att1_b.err
CMD="python bench.py --utts 400 --steps 1 --warmup 1 --no-cpu-baseline"
PLLB_ATT_TMA=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'attention' -c 120 --csv --log-file gpurun_out/r2l_launches_att.csv $CMD > gpurun_out/r2l_ncu1.log 2>&1
tail -n 3 gpurun_out/r2l_pytest.log
