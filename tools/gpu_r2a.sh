#!/bin/bash
# round-2 first GPU pass: parity tests of the two-pass approximate stage + A/B bench lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_api.py -x -q > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
run() { name=$1; shift; timeout 300 env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r2a_bench_$name.json 2> gpurun_out/r2a_bench_$name.err; }
EXTRA="" run def X=1
EXTRA="--approx direct" run direct X=1
EXTRA="" run minb5 FPB_K3_MINB=5
EXTRA="" run q40 FPB_K3_TAU_Q=0.4
EXTRA="" run q60 FPB_K3_TAU_Q=0.6
EXTRA="" run q30 FPB_K3_TAU_Q=0.3
EXTRA="--config cfg3c" run c_def X=1
EXTRA="--config cfg3c --approx direct" run c_direct X=1
EXTRA="--config cfg5" run cfg5_def X=1
EXTRA="--config cfg5 --approx direct" run cfg5_direct X=1
tail -5 gpurun_out/r2a_tests.log
for f in gpurun_out/r2a_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(round(d["value"]), d["ms_per_step"], d["stages_ms"], {k:v for k,v in d["approx_stage"].items() if "row" in k or "refined" in k})
except Exception as e: print("ERR", e)
PY
done
