OUT=gpurun_out/r03h; mkdir -p $OUT
for v in "" _exact _shift _both; do
  echo "=== variant '$v'"
  EOGS_RASTER_LIB=$PWD/eogs2_b200/libeogs_raster$v.so timeout 600 python tools/fuzz_check.py tools/fuzz_r03f.json > $OUT/check$v.json 2> $OUT/check$v.err
  python - <<PY
import json
d=json.load(open("$OUT/check$v.json"))
for r in d:
    print(r['case'], {k:('%.1e'%x['ours_vs_oracle']) for k,x in r.items() if isinstance(x,dict) and k!='dL_drotations'})
PY
  EOGS_RASTER_LIB=$PWD/eogs2_b200/libeogs_raster$v.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench$v.json 2> $OUT/bench$v.err
  python -c "
import json; d=json.load(open('$OUT/bench$v.json')); print('ms/step', d['ms_per_step'], 'bwd', d['stage_ms']['blend_bwd'])"
done
