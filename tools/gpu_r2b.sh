#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_api.py -x -q > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.log
run() { name=$1; shift; timeout 300 env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r2b_bench_$name.json 2> gpurun_out/r2b_bench_$name.err; }
EXTRA="" run def X=1
EXTRA="" run minb4 FPB_K3_MINB=4
EXTRA="" run minb6 FPB_K3_MINB=6
EXTRA="" run l10 FPB_K3_LAMBDA=1.0
EXTRA="" run l25 FPB_K3_LAMBDA=2.5
EXTRA="" run l40 FPB_K3_LAMBDA=4.0
EXTRA="--config cfg3c" run c_def X=1
EXTRA="--config cfg3c" run c_l10 FPB_K3_LAMBDA=1.0
EXTRA="--config cfg3c" run c_l30 FPB_K3_LAMBDA=3.0
EXTRA="--config cfg5" run cfg5_def X=1
EXTRA="--config cfg2" run cfg2_def X=1
# per-kernel durations of the default build (ncu launch list; numbers printed under ncu are not bench values)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_under_ncu.log 2>&1
tail -5 gpurun_out/r2b_tests.log
for f in gpurun_out/r2b_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(round(d["value"]), round(d["ms_per_step"],3), d["stages_ms"], {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["approx_stage"].items() if "row" in k or "refined" in k})
except Exception as e: print("ERR", e)
PY
done
