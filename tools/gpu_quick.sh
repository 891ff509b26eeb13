#!/bin/bash
# Quick GPU visit: parity tests + our bench arm (optionally with extra env A/B legs).
# Usage: bash tools/gpu_quick.sh <tag> ["ENV=1 ENV2=1" ...]   each extra arg = one more bench leg with that env
TAG=${1:-quick}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_ours.json 2> $OUT/bench_ours.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_ours.json")); print("ours", d["ms_per_step"], d["e2e"], d["stage_ms"])
except Exception as e: print("bench failed", e); print(open("$OUT/bench_ours.err").read()[-2000:])
PY
n=0
for envs in "$@"; do
  n=$((n+1))
  env $envs timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench_leg$n.json 2> $OUT/bench_leg$n.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_leg$n.json")); print("leg$n [$envs]", d["ms_per_step"], d["stage_ms"])
except Exception as e: print("leg failed", e); print(open("$OUT/bench_leg$n.err").read()[-2000:])
PY
done
