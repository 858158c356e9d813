#!/bin/bash
# round-2 GPU call X (1 GPU): final validation — smoke, the whole GPU suite, the default bench line and the reference arm, launch list
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2x_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
timeout 900 python bench.py > gpurun_out/r2x_bench_c2.json 2> gpurun_out/r2x_bench_c2.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/r2x_bench_c2_reference_arm.json 2> gpurun_out/r2x_bench_c2_reference_arm.err; echo "ref rc=$?"
timeout 600 python bench.py --workload c1 > gpurun_out/r2x_bench_c1.json 2> gpurun_out/r2x_bench_c1.err
timeout 900 python bench.py --workload c4 --steps 1 --warmup 1 > gpurun_out/r2x_bench_c4.json 2> gpurun_out/r2x_bench_c4.err
CMD="python bench.py --utts 400 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2x_launches.csv $CMD > gpurun_out/r2x_ncu_launches.log 2>&1
tail -2 gpurun_out/r2x_smoke.log; tail -12 gpurun_out/r2x_pytest.log
for f in gpurun_out/r2x_bench_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d.get('impl','b200'), d['dtype'], round(d['value'],1), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), d.get('clocks'), (d.get('roofline') or {}).get('frac'), d.get('cpu_baseline',{}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
P
done
