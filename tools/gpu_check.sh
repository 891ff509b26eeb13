#!/bin/bash
# Standard visit: GPU tests (fail fast), a short bench with stage times, optionally the ncu launch list.
# Usage: bash tools/gpu_check.sh <tag> [launches]
TAG=${1:-check}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; grep -n "Error\|passed\|failed" $OUT/pytest_gpu.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${BENCH_FLAGS:---no-e2e} > $OUT/bench.json 2> $OUT/bench.err; python -c "
import json; d=json.load(open('$OUT/bench.json')); print(round(d['ms_per_step'],4), d['stage_ms'], 'iter_ms', round(d['iter_ms'],3), 'e2e', d.get('e2e'))" || tail -5 $OUT/bench.err
if [ "$2" = "launches" ]; then bash tools/gpu_launches.sh $TAG | head -30; fi
