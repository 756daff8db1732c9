#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2j_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2j_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench_cfg3.json 2> gpurun_out/r2j_bench_cfg3.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2j_bench_reference.json 2> gpurun_out/r2j_bench_reference.err
run() { name=$1; shift; timeout 300 env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r2j_bench_$name.json 2> gpurun_out/r2j_bench_$name.err; }
EXTRA="--approx direct" run cfg3_direct X=1
EXTRA="--config cfg2" run cfg2 X=1
EXTRA="--config cfg3c" run cfg3c X=1
EXTRA="--config cfg5" run cfg5 X=1
EXTRA="--config tiny" run tiny X=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_under_ncu.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k3_bound_kernel -s 3 -c 1 -o gpurun_out/r2j_k3_bound python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_ncu_bound.log 2>&1
ncu -i gpurun_out/r2j_k3_bound.ncu-rep --page raw --csv > gpurun_out/r2j_k3_bound_raw.csv 2>/dev/null
tail -6 gpurun_out/r2j_tests.log; tail -3 gpurun_out/r2j_smoke.log
for f in gpurun_out/r2j_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(round(d["value"],2), round(d["ms_per_step"],3), round(d["e2e"]["value"],2), d.get("stages_ms"), d.get("roofline",{}).get("frac"))
    for k in ("cpu_baseline","parity_sample"):
        if k in d: print(k, json.dumps(d[k])[:1200])
except Exception as e: print("ERR", e, open(sys.argv[1]).read()[:500])
PY
done
