#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2e_tests.log
run() { name=$1; shift; timeout 300 env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r2e_bench_$name.json 2> gpurun_out/r2e_bench_$name.err; }
EXTRA="" run s0 FPB_K3_SHAPE=0
EXTRA="" run s2 FPB_K3_SHAPE=2
EXTRA="" run s3 FPB_K3_SHAPE=3
EXTRA="" run s0_l20 FPB_K3_SHAPE=0 FPB_K3_LAMBDA=2.0
EXTRA="" run s0_l13 FPB_K3_SHAPE=0 FPB_K3_LAMBDA=1.3
EXTRA="--config cfg3c" run c_s0 FPB_K3_SHAPE=0
EXTRA="--config cfg3c" run c_s0_l80 FPB_K3_SHAPE=0 FPB_K3_LAMBDA=8.0
EXTRA="--config cfg3c" run c_s0_l160 FPB_K3_SHAPE=0 FPB_K3_LAMBDA=16.0
EXTRA="--config cfg3c --approx direct" run c_direct X=1
EXTRA="--config cfg2" run cfg2_s0 FPB_K3_SHAPE=0
EXTRA="--config cfg5" run cfg5_s0 FPB_K3_SHAPE=0
EXTRA="--config cfg5" run cfg5_l40 FPB_K3_LAMBDA=4.0
EXTRA="--config cfg5 --approx direct" run cfg5_direct X=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_under_ncu.log 2>&1
tail -8 gpurun_out/r2e_tests.log
for f in gpurun_out/r2e_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["stages_ms"], {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["approx_stage"].items() if "row" in k or "refined" in k}, round(d["roofline"]["frac"],3))
except Exception as e: print("ERR", e)
PY
done
