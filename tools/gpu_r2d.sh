#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2d_tests.log
run() { name=$1; shift; timeout 300 env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r2d_bench_$name.json 2> gpurun_out/r2d_bench_$name.err; }
EXTRA="" run s0 FPB_K3_SHAPE=0
EXTRA="" run s1 FPB_K3_SHAPE=1
EXTRA="" run s2 FPB_K3_SHAPE=2
EXTRA="" run s3 FPB_K3_SHAPE=3
EXTRA="" run s4 FPB_K3_SHAPE=4
EXTRA="" run s0_l25 FPB_K3_SHAPE=0 FPB_K3_LAMBDA=2.5
EXTRA="" run s0_l20 FPB_K3_SHAPE=0 FPB_K3_LAMBDA=2.0
EXTRA="--config cfg3c" run c_s0 FPB_K3_SHAPE=0
EXTRA="--config cfg3c" run c_s0_l40 FPB_K3_SHAPE=0 FPB_K3_LAMBDA=4.0
EXTRA="--config cfg3c" run c_s0_l80 FPB_K3_SHAPE=0 FPB_K3_LAMBDA=8.0
EXTRA="--config cfg2" run cfg2_s0 FPB_K3_SHAPE=0
EXTRA="--config cfg5" run cfg5_s0 FPB_K3_SHAPE=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_under_ncu.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k3_bound_kernel -s 3 -c 1 -o gpurun_out/r2d_k3_bound python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_ncu_bound.log 2>&1
ncu -i gpurun_out/r2d_k3_bound.ncu-rep --page raw --csv > gpurun_out/r2d_k3_bound_raw.csv 2>/dev/null
ncu -i gpurun_out/r2d_k3_bound.ncu-rep --page source --csv > gpurun_out/r2d_k3_bound_source.csv 2>/dev/null
tail -5 gpurun_out/r2d_tests.log
for f in gpurun_out/r2d_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["stages_ms"], {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["approx_stage"].items() if "row" in k or "refined" in k})
except Exception as e: print("ERR", e)
PY
done
