OUT=gpurun_out/r03m; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_ours.json 2> $OUT/bench.err
python -c "
import json; d=json.load(open('$OUT/bench_ours.json')); print('ms/step', d['ms_per_step'], d['stage_ms'], 'e2e', d['e2e']['value'], 'iter', d['iter_ms'])"
timeout 900 python tools/fuzz_parity.py --cases 300 --seed 5 > $OUT/fuzz_seed5.json 2> $OUT/fuzz_seed5.err; tail -2 $OUT/fuzz_seed5.err
timeout 900 python tools/bench_configs.py --configs 1,3 --no-ref --out $OUT/configs13.json 2>&1 | tail -8
