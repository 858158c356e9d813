#!/bin/bash
# round-2 GPU call E (1 GPU): full GPU suite, bench lines of every workload, ncu rows of the small kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench_c2.json 2> gpurun_out/r2e_bench_c2.err
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/r2e_bench_c4.json 2> gpurun_out/r2e_bench_c4.err
timeout 600 python bench.py --workload c1 --steps 10 --warmup 3 > gpurun_out/r2e_bench_c1.json 2> gpurun_out/r2e_bench_c1.err
timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/r2e_bench_c5.json 2> gpurun_out/r2e_bench_c5.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2e_ref_c2.json 2> gpurun_out/r2e_ref_c2.err
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active"
CMD2="python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none -k regex:'levenshtein|rescore_sweep|lse_finish|hyp_sum|gather_rows|attention_row|^ln_kernel|gemm_tcgen05_kernel<4|gemm_tcgen05_kernel<3|expand_plan|embed_unique' -c 24 -o gpurun_out/r2e_prof_small $CMD2 > gpurun_out/r2e_ncu3.log 2>&1
ncu -i gpurun_out/r2e_prof_small.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2e_prof_small.csv 2>/dev/null
rm -f gpurun_out/r2e_prof_small.ncu-rep
CMD="python bench.py --utts 400 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2e_launches.csv $CMD > gpurun_out/r2e_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'gemm_ln_kernel|gemm_tcgen05_kernel|attention_mma_kernel' -s 30 -c 5 -o gpurun_out/r2e_prof_layer $CMD > gpurun_out/r2e_ncu2.log 2>&1
ncu -i gpurun_out/r2e_prof_layer.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2e_prof_layer.csv 2>/dev/null
rm -f gpurun_out/r2e_prof_layer.ncu-rep
tail -n 3 gpurun_out/r2e_pytest.log
du -sh gpurun_out
