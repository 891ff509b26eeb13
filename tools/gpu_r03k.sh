OUT=gpurun_out/r03k; mkdir -p $OUT
timeout 600 python tools/fuzz_check.py tools/fuzz_r03f.json > $OUT/check.json 2> $OUT/check.err
python - <<PY
import json
d=json.load(open("$OUT/check.json"))
for r in d:
    print(r['case'], {k:('%.1e'%x['ours_vs_oracle']) for k,x in r.items() if isinstance(x,dict) and k!='dL_drotations'})
PY
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_ours.json 2> $OUT/bench.err
python -c "
import json; d=json.load(open('$OUT/bench_ours.json')); print('ms/step', d['ms_per_step'], d['stage_ms'], 'e2e', d['e2e']['value'], 'iter', d['iter_ms'])"
timeout 900 python tools/fuzz_parity.py --cases 1500 --seed 1 > $OUT/fuzz_seed1.json 2> $OUT/fuzz_seed1.err; tail -3 $OUT/fuzz_seed1.err
timeout 900 python tools/fuzz_parity.py --cases 1500 --seed 3 > $OUT/fuzz_seed3.json 2> $OUT/fuzz_seed3.err; tail -3 $OUT/fuzz_seed3.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
