#!/bin/bash
# round-2 GPU call C (1 GPU): full GPU suite with the new default operand mode, SM-count sensitivity,
# ncu launch list + full captures
mkdir -p gpurun_out
PLLB_C2_GOLDEN_MIN_UTTS=1000 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
timeout 600 $B > gpurun_out/r2c_sm_default.json 2> gpurun_out/r2c_sm_default.err
PLLB_LN_MAX_CLUSTERS=40 timeout 600 $B > gpurun_out/r2c_sm_ln40.json 2> gpurun_out/r2c_sm_ln40.err
PLLB_GEMM_MAX_CTAS=132 timeout 600 $B > gpurun_out/r2c_sm_gemm132.json 2> gpurun_out/r2c_sm_gemm132.err
timeout 600 $B > gpurun_out/r2c_sm_default2.json 2> gpurun_out/r2c_sm_default2.err
CMD="python bench.py --utts 400 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/r2c_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2c_launches.csv $CMD > gpurun_out/r2c_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_ln_kernel|gemm_tcgen05_kernel|attention_mma_kernel' -s 30 -c 6 -o gpurun_out/r2c_prof_layer $CMD > gpurun_out/r2c_ncu2.log 2>&1
CMD2="python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'levenshtein|rescore_sweep|lse_finish|hyp_sum|embed_unique|expand_plan|row_src|rowmajor_to_t32|gather_rows|attention_row|ln_kernel' -c 14 -o gpurun_out/r2c_prof_small $CMD2 > gpurun_out/r2c_ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'gemm_tcgen05_kernel<4' -c 1 -o gpurun_out/r2c_prof_lse $CMD2 > gpurun_out/r2c_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'tokenize' -c 4 -o gpurun_out/r2c_prof_tok python tools/tokenize_probe.py > gpurun_out/r2c_ncu5.log 2>&1
tail -n 3 gpurun_out/r2c_pytest.log
ls -la gpurun_out/*.ncu-rep
