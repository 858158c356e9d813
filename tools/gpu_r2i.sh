#!/bin/bash
# round-2 GPU call I (1 GPU): attention instruction diet (MUFU-only exp2, masked key-tile skip) ABAB against the previous build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "attention or layers or layer0 or golden or determin or tma or duplicate or degenerate" > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
PLLB_LIB=$PWD/asr-rescoring_b200/libpllb200_prev.so timeout 600 $B > gpurun_out/r2i_prev_a.json 2> gpurun_out/r2i_prev_a.err
timeout 600 $B > gpurun_out/r2i_new_a.json 2> gpurun_out/r2i_new_a.err
PLLB_LIB=$PWD/asr-rescoring_b200/libpllb200_prev.so timeout 600 $B > gpurun_out/r2i_prev_b.json 2> gpurun_out/r2i_prev_b.err
timeout 600 $B > gpurun_out/r2i_new_b.json 2> gpurun_out/r2i_new_b.err
CMD="python bench.py --utts 400 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'attention' -c 120 --csv --log-file gpurun_out/r2i_launches_att.csv $CMD > gpurun_out/r2i_ncu1.log 2>&1
tail -n 3 gpurun_out/r2i_pytest.log
