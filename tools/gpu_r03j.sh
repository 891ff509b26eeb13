OUT=gpurun_out/r03j; mkdir -p $OUT
timeout 600 python tools/fuzz_check.py tools/fuzz_r03f.json > $OUT/check.json 2> $OUT/check.err
python - <<PY
import json
d=json.load(open("$OUT/check.json"))
for r in d:
    print(r['case'], {k:('%.1e'%x['ours_vs_oracle']) for k,x in r.items() if isinstance(x,dict) and k!='dL_drotations'})
PY
for v in "" _noborder; do
  EOGS_RASTER_LIB=$PWD/eogs2_b200/libeogs_raster$v.so timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench$v.json 2> $OUT/bench$v.err
  python -c "
import json; d=json.load(open('$OUT/bench$v.json')); print('variant [$v] ms/step', d['ms_per_step'], 'bwd', d['stage_ms']['blend_bwd'])"
done
timeout 900 python tools/fuzz_parity.py --cases 1500 --seed 1 > $OUT/fuzz_seed1.json 2> $OUT/fuzz_seed1.err; tail -3 $OUT/fuzz_seed1.err
timeout 900 python tools/fuzz_parity.py --cases 1500 --seed 2 > $OUT/fuzz_seed2.json 2> $OUT/fuzz_seed2.err; tail -3 $OUT/fuzz_seed2.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
