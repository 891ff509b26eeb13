#!/bin/bash
# ncu --set full capture of kernels matching a regex during a short bench run (one launch per kernel by default).
# Usage: bash tools/gpu_ncu.sh <tag> <kernel-regex> [skip] [count]
TAG=$1; PAT=$2; SKIP=${3:-3}; CNT=${4:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$PAT" -s $SKIP -c $CNT \
    -o $OUT/prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-config5 > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log | cut -c1-300
ls -la $OUT
