#!/bin/bash
# compute-sanitizer over tools/sanitize_run.py with all four tools.  Usage: bash tools/gpu_sanitize.sh <tag>
TAG=${1:-sanitize}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > $OUT/$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard" $OUT/$tool.log | head -12
done
