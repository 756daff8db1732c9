#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2g_tests.log
run() { name=$1; shift; timeout 300 env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r2g_bench_$name.json 2> gpurun_out/r2g_bench_$name.err; }
EXTRA="" run def X=1
EXTRA="" run nowin FPB_K5_NO_L2_WINDOW=1
EXTRA="" run s3 FPB_K3_SHAPE=3
EXTRA="" run s1 FPB_K3_SHAPE=1
EXTRA="--config cfg3c" run cfg3c X=1
EXTRA="--config cfg5" run cfg5 X=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k5_maxsim_v4 -s 3 -c 2 --csv --log-file gpurun_out/r2g_k5_traffic_window.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
FPB_K5_NO_L2_WINDOW=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k5_maxsim_v4 -s 3 -c 2 --csv --log-file gpurun_out/r2g_k5_traffic_nowindow.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
bash tools/gpu_sanitize.sh > gpurun_out/r2g_sanitize_summary.txt 2>&1
tail -6 gpurun_out/r2g_tests.log
for f in gpurun_out/r2g_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(round(d["value"],2), round(d["ms_per_step"],3), round(d["e2e"]["value"],2), d.get("stages_ms"), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.get("approx_stage",{}).items() if "row" in k or "refined" in k}, d.get("roofline",{}).get("frac"))
except Exception as e: print("ERR", e, open(sys.argv[1]).read()[:500])
PY
done
cat gpurun_out/r2g_k5_traffic_window.csv | tail -9; cat gpurun_out/r2g_k5_traffic_nowindow.csv | tail -9
cat gpurun_out/r2g_sanitize_summary.txt
