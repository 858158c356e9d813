#!/bin/bash
# round-2 GPU call T (1 GPU): ncu --set full of the fused GEMM+LayerNorm kernels after the shared-memory parameters (source view)
mkdir -p gpurun_out
CMD="python bench.py --utts 400 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/r2t_plain.json 2> gpurun_out/r2t_plain.err || exit 1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_ln_kernel -s 24 -c 4 -f -o gpurun_out/r2t_ln $CMD > gpurun_out/r2t_ncu.log 2>&1
ls -la gpurun_out/r2t_ln.ncu-rep; tail -3 gpurun_out/r2t_ncu.log
