#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/last_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/last_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/last_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/last_bench_cfg3.json 2> gpurun_out/last_bench_cfg3.err
bash tools/gpu_sanitize.sh > gpurun_out/last_sanitize_summary.txt 2>&1
tail -3 gpurun_out/last_tests.log; tail -2 gpurun_out/last_smoke.log; python -c "
import json; d=json.load(open('gpurun_out/last_bench_cfg3.json')); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['stages_ms'], d['roofline']['frac'])"
cat gpurun_out/last_sanitize_summary.txt
