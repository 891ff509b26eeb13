#!/bin/bash
# Generic GPU visit: run the given shell snippets, teeing into gpurun_out/<tag>/.
# Usage: bash tools/gpu_cmd.sh <tag> '<cmd1>' '<cmd2>' ...   (each cmd's stdout+stderr -> step<N>.log)
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
export OUT
n=0
for c in "$@"; do
  n=$((n+1))
  echo "=== step $n: $c" | tee $OUT/step$n.log
  bash -c "$c" >> $OUT/step$n.log 2>&1
  echo "rc=$?" >> $OUT/step$n.log
  tail -25 $OUT/step$n.log | cut -c1-400
done
