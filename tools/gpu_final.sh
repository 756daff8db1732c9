#!/bin/bash
# what the driver runs at round end, plus the ncu captures committed under profiles/
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_cfg3.json 2> gpurun_out/final_bench_cfg3.err
run() { name=$1; shift; timeout 300 env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/final_bench_$name.json 2> gpurun_out/final_bench_$name.err; }
EXTRA="--config cfg2" run cfg2 X=1
EXTRA="--config cfg3c" run cfg3c X=1
EXTRA="--config cfg5" run cfg5 X=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_under_ncu.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k5_maxsim_v4 -s 3 -c 1 -o gpurun_out/final_k5v4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_k5.log 2>&1
ncu -i gpurun_out/final_k5v4.ncu-rep --page raw --csv > gpurun_out/final_k5v4_raw.csv 2>/dev/null
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k3_bound_kernel -s 3 -c 1 -o gpurun_out/final_k3_bound python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_bound.log 2>&1
ncu -i gpurun_out/final_k3_bound.ncu-rep --page raw --csv > gpurun_out/final_k3_bound_raw.csv 2>/dev/null
tail -4 gpurun_out/final_tests.log; tail -2 gpurun_out/final_smoke.log
for f in gpurun_out/final_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(round(d["value"],2), round(d["ms_per_step"],3), round(d["e2e"]["value"],2), d.get("stages_ms"), d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("traffic"))
    for k in ("cpu_baseline","parity_sample"):
        if k in d: print(k, json.dumps(d[k])[:700])
except Exception as e: print("ERR", e, open(sys.argv[1]).read()[:500])
PY
done
