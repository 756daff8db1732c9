#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2i_tests.log
run() { name=$1; shift; timeout 300 env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r2i_bench_$name.json 2> gpurun_out/r2i_bench_$name.err; }
EXTRA="" run s0 FPB_K3_SHAPE=0
EXTRA="" run s1 FPB_K3_SHAPE=1
EXTRA="" run s2 FPB_K3_SHAPE=2
EXTRA="" run s3 FPB_K3_SHAPE=3
EXTRA="" run s4 FPB_K3_SHAPE=4
FPB_K3_SHAPE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_under_ncu.log 2>&1
FPB_K3_SHAPE=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k3_bound_kernel -s 3 -c 1 -o gpurun_out/r2i_k3_bound python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_ncu_bound.log 2>&1
ncu -i gpurun_out/r2i_k3_bound.ncu-rep --page raw --csv > gpurun_out/r2i_k3_bound_raw.csv 2>/dev/null
ncu -i gpurun_out/r2i_k3_bound.ncu-rep --page source --csv > gpurun_out/r2i_k3_bound_source.csv 2>/dev/null
tail -6 gpurun_out/r2i_tests.log
for f in gpurun_out/r2i_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(round(d["value"],2), round(d["ms_per_step"],3), round(d["e2e"]["value"],2), d.get("stages_ms"), d.get("roofline",{}).get("frac"))
except Exception as e: print("ERR", e, open(sys.argv[1]).read()[:500])
PY
done
