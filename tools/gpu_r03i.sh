OUT=gpurun_out/r03i; mkdir -p $OUT
for v in "" _xexp _xdiv; do
  echo "=== variant '$v'"
  EOGS_RASTER_LIB=$PWD/eogs2_b200/libeogs_raster$v.so timeout 600 python tools/fuzz_debug.py tools/fuzz_r03f.json 1319 2>&1 | tail -22
  EOGS_RASTER_LIB=$PWD/eogs2_b200/libeogs_raster$v.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench$v.json 2> $OUT/bench$v.err
  python -c "
import json; d=json.load(open('$OUT/bench$v.json')); print('ms/step', d['ms_per_step'], 'bwd', d['stage_ms']['blend_bwd'])"
done
