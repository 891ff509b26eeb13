#!/bin/bash
# Repeat the GPU suite to surface run-to-run flakiness (atomics order), then a randomised parity sweep.
# Usage: bash tools/gpu_flaky.sh <tag> [repeats] [fuzz cases] [fuzz seed]
TAG=${1:-flaky}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for i in $(seq 1 ${2:-4}); do
  timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_$i.log 2>&1
  echo "run $i: $(tail -1 $OUT/pytest_$i.log | cut -c1-120)"; grep -E "^FAILED" $OUT/pytest_$i.log | cut -c1-200
done
bash tools/gpu_fuzz.sh $TAG ${3:-2000} ${4:-4}
