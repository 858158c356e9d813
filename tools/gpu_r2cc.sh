#!/bin/bash
# round-2 GPU call CC (1 GPU): final build — whole GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2cc_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2cc_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2cc_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2cc_pytest.log
timeout 900 python bench.py > gpurun_out/r2cc_bench_c2.json 2> gpurun_out/r2cc_bench_c2.err; echo "bench rc=$?"
PLLB_FFN_BLOCKED=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "pll_vs_reference_golden or deterministic or degenerate" > gpurun_out/r2cc_pytest_blocked.log 2>&1; echo "blocked rc=$?" >> gpurun_out/r2cc_pytest_blocked.log
tail -2 gpurun_out/r2cc_smoke.log; grep "bert-base-chinese, 12 layers\|passed\|failed\|rc=" gpurun_out/r2cc_pytest.log | cut -c1-250; tail -2 gpurun_out/r2cc_pytest_blocked.log
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2cc_bench_c2.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['roofline']['frac'], d['cpu_baseline'])
P
