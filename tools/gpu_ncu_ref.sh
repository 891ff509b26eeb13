#!/bin/bash
# ncu --set full capture of the REFERENCE arm's kernels (oracle/_ref, the reference's own CUDA code) on the bench workload.
# Usage: bash tools/gpu_ncu_ref.sh <tag>
TAG=$1
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/ref_launches.csv \
    python bench.py --impl reference --steps 2 --warmup 3 > $OUT/ncu_ref_launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'renderCUDA|preprocessCUDA|computeCov2DCUDA|duplicateWithKeys|identifyTileRanges' -s 10 -c 7 \
    -o $OUT/ref -f python bench.py --impl reference --steps 2 --warmup 3 > $OUT/ncu_ref_full.log 2>&1
tail -3 $OUT/ncu_ref_full.log | cut -c1-300
ls -la $OUT
