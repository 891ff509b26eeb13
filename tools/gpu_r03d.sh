bash tools/gpu_cmd.sh r03d \
 'timeout 1200 python -m pytest tests -m gpu -x -q' \
 'timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_ours.json' \
 'timeout 900 python tools/bench_configs.py --configs 3 --no-ref --out $OUT/configs3.json'
