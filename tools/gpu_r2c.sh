#!/bin/bash
# round-2 GPU call C2 (1 GPU): TMA-fed attention kernel (tests + ABAB), SM-count sensitivity, ncu launch list + full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "attention or layers or layer0 or golden or determin or cluster_sizes or degenerate or duplicate" > gpurun_out/r2c_pytest_att.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest_att.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
PLLB_ATT_TMA=0 timeout 600 $B > gpurun_out/r2c_att0_a.json 2> gpurun_out/r2c_att0_a.err
timeout 600 $B > gpurun_out/r2c_att1_a.json 2> gpurun_out/r2c_att1_a.err
PLLB_ATT_TMA=0 timeout 600 $B > gpurun_out/r2c_att0_b.json 2> gpurun_out/r2c_att0_b.err
timeout 600 $B > gpurun_out/r2c_att1_b.json 2> gpurun_out/r2c_att1_b.err
PLLB_LN_MAX_CLUSTERS=40 timeout 600 $B > gpurun_out/r2c_sm_ln40.json 2> gpurun_out/r2c_sm_ln40.err
PLLB_GEMM_MAX_CTAS=132 timeout 600 $B > gpurun_out/r2c_sm_gemm132.json 2> gpurun_out/r2c_sm_gemm132.err
CMD="python bench.py --utts 400 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2c_launches.csv $CMD > gpurun_out/r2c_ncu1.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,l1tex__data_bank_conflicts_pipe_lsu.sum"
timeout 900 ncu --set full --clock-control none -k regex:'gemm_ln_kernel|gemm_tcgen05_kernel|attention_mma_kernel|attention_tma_kernel' -s 30 -c 7 -o gpurun_out/r2c_prof_layer $CMD > gpurun_out/r2c_ncu2.log 2>&1
ncu -i gpurun_out/r2c_prof_layer.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2c_prof_layer.csv 2>/dev/null
ncu -i gpurun_out/r2c_prof_layer.ncu-rep --page details --csv > gpurun_out/r2c_prof_layer_details.csv 2>/dev/null
rm -f gpurun_out/r2c_prof_layer.ncu-rep
CMD2="python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none -k regex:'levenshtein|rescore_sweep|lse_finish|hyp_sum|embed_unique|expand_plan|row_src|rowmajor_to_t32|gather_rows|attention_row|ln_kernel|gemm_tcgen05_kernel<4' -c 16 -o gpurun_out/r2c_prof_small $CMD2 > gpurun_out/r2c_ncu3.log 2>&1
ncu -i gpurun_out/r2c_prof_small.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2c_prof_small.csv 2>/dev/null
rm -f gpurun_out/r2c_prof_small.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:'tokenize' -c 4 -o gpurun_out/r2c_prof_tok python tools/tokenize_probe.py > gpurun_out/r2c_ncu5.log 2>&1
ncu -i gpurun_out/r2c_prof_tok.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2c_prof_tok.csv 2>/dev/null
rm -f gpurun_out/r2c_prof_tok.ncu-rep
tail -n 3 gpurun_out/r2c_pytest_att.log
du -sh gpurun_out
