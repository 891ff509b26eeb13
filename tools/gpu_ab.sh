#!/bin/bash
# A/B visit: parity tests on the product library, then one short bench leg per library variant
# (eogs2_b200/libeogs_raster_<name>.so, built with EOGS_LIB_SUFFIX=_<name>), optionally an ncu capture.
# Usage: bash tools/gpu_ab.sh <tag> [variant ...]      NCU_PAT=<kernel regex> adds a --set full capture of the product library
TAG=${1:-ab}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
leg() {  # name, env
  env $2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-config5 > $OUT/bench_$1.json 2> $OUT/bench_$1.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$1.json")); print("$1", round(d["ms_per_step"],4), d["stage_ms"])
except Exception as e: print("$1 failed", e); print(open("$OUT/bench_$1.err").read()[-1500:])
PY
}
leg product ""
for v in "$@"; do
  leg $v "EOGS_RASTER_LIB=$PWD/eogs2_b200/libeogs_raster_$v.so"
  if [ -n "$PARITY_VARIANTS" ]; then
    EOGS_RASTER_LIB=$PWD/eogs2_b200/libeogs_raster_$v.so timeout 600 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -2
  fi
done
leg product2 ""
if [ -n "$NCU_PAT" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$NCU_PAT" -s ${NCU_SKIP:-3} -c ${NCU_CNT:-1} \
      -o $OUT/prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
  tail -2 $OUT/ncu_full.log | cut -c1-200
fi
ls -la $OUT
