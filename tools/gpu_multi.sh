#!/bin/bash
# multi-GPU pass: NCCL tests + sharded bench lines.  usage: gpu_multi_multi.sh <ngpus>
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/multi_tests_n$N.log 2>&1
echo "tests rc=$?" >> gpurun_out/multi_tests_n$N.log
tail -4 gpurun_out/multi_tests_n$N.log
run() { name=$1; shift; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/multi_bench_${name}_n$N.json 2> gpurun_out/multi_bench_${name}_n$N.err; echo "$name rc=$?"; }
if [ "$N" = "2" ]; then
  run cfg3 --no-cpu-baseline
  run cfg5 --config cfg5 --no-cpu-baseline
  run cfg3_par --parity-queries 8
elif [ "$N" = "4" ]; then
  run cfg3 --no-cpu-baseline
  run cfg3_par --parity-queries 4
else
  run cfg3 --no-cpu-baseline
  run cfg3_g1 --no-cpu-baseline --query-groups 1
  run cfg4 --config cfg4 --no-cpu-baseline
  run cfg3_par --parity-queries 4
fi
for f in gpurun_out/multi_bench_*_n$N.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d.get("stages_ms"), d["config"]["parallelism"][:60])
    if "parity_sample" in d: print("parity", json.dumps(d["parity_sample"])[:900])
    if "cpu_baseline" in d: print("cpu", json.dumps(d["cpu_baseline"])[:300])
except Exception as e: print("ERR", e, open(sys.argv[1]).read()[:300]); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
