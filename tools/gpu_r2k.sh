#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2k_tests.log
run() { name=$1; shift; timeout 300 env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r2k_bench_$name.json 2> gpurun_out/r2k_bench_$name.err; }
EXTRA="" run def X=1
EXTRA="" run unpaired FPB_K3_SHAPE=1
EXTRA="" run k5minb3 FPB_K5_MINB=3
EXTRA="" run l18 FPB_K3_LAMBDA=1.8
EXTRA="" run l22 FPB_K3_LAMBDA=2.2
EXTRA="--config cfg5" run cfg5 X=1
EXTRA="--config cfg2" run cfg2 X=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:k5_maxsim_v4 -s 3 -c 1 --csv --log-file gpurun_out/r2k_k5_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -6 gpurun_out/r2k_tests.log
for f in gpurun_out/r2k_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(round(d["value"],2), round(d["ms_per_step"],3), round(d["e2e"]["value"],2), d.get("stages_ms"), d.get("roofline",{}).get("frac"), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.get("approx_stage",{}).items() if "row" in k or "refined" in k})
except Exception as e: print("ERR", e, open(sys.argv[1]).read()[:500])
PY
done
cut -d, -f13-15 gpurun_out/r2k_k5_metrics.csv | tail -8
