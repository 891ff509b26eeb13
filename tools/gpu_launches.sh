#!/bin/bash
# ncu launch list (gpu__time_duration per launch, cold-cache and serialised: compare SHARES) of one short bench run.
# Usage: bash tools/gpu_launches.sh <tag>
TAG=${1:-launches}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launches.log 2>&1
python tools/ncu_summary.py launches $OUT/launches.csv 2>/dev/null | head -60 || tail -5 $OUT/ncu_launches.log
