#!/bin/bash
# Randomised parity sweep against the compiled reference.  Usage: bash tools/gpu_fuzz.sh <tag> [cases] [seed]
TAG=${1:-fuzz}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python tools/fuzz_parity.py --cases ${2:-3000} --seed ${3:-2} > $OUT/fuzz_parity.json 2> $OUT/fuzz.err
tail -3 $OUT/fuzz.err | cut -c1-300; head -12 $OUT/fuzz_parity.json
