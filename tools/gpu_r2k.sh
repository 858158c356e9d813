#!/bin/bash
# round-2 GPU call K (1 GPU): final suite (full C2 golden), final bench lines of every workload, ncu
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2k_bench_c2.json 2> gpurun_out/r2k_bench_c2.err
timeout 900 python bench.py --workload c4 --steps 3 --warmup 2 > gpurun_out/r2k_bench_c4.json 2> gpurun_out/r2k_bench_c4.err
timeout 600 python bench.py --workload c1 --steps 10 --warmup 3 > gpurun_out/r2k_bench_c1.json 2> gpurun_out/r2k_bench_c1.err
timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/r2k_bench_c5.json 2> gpurun_out/r2k_bench_c5.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_ref_c2.json 2> gpurun_out/r2k_ref_c2.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active"
CMD="python bench.py --utts 400 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2k_launches.csv $CMD > gpurun_out/r2k_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'gemm_ln_kernel|gemm_tcgen05_kernel|attention_mma_kernel' -s 30 -c 5 -o gpurun_out/r2k_prof_layer $CMD > gpurun_out/r2k_ncu2.log 2>&1
ncu -i gpurun_out/r2k_prof_layer.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2k_prof_layer.csv 2>/dev/null
rm -f gpurun_out/r2k_prof_layer.ncu-rep
tail -n 3 gpurun_out/r2k_pytest.log; cat gpurun_out/r2k_smoke.log | tail -2
