#!/bin/bash
# Small-scene visit: parity tests, the config-1 probe (tools/small_scene.py) and one short bench leg.
# Usage: bash tools/gpu_small.sh <tag>
TAG=${1:-small}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log | cut -c1-300
for kind in trained init; do python tools/small_scene.py 50000 512 512 $kind >> $OUT/small.json 2>> $OUT/small.err; done
python tools/small_scene.py 300000 1024 1024 trained >> $OUT/small.json 2>> $OUT/small.err
cat $OUT/small.json; tail -3 $OUT/small.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-config5 > $OUT/bench.json 2> $OUT/bench.err
python -c "
import json; d=json.load(open('$OUT/bench.json')); print(round(d['ms_per_step'],4), d['stage_ms'])"
