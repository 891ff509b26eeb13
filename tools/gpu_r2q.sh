mkdir -p gpurun_out/r2q
timeout 1200 python tools/fuzz_parity.py --cases 3000 --seed 2 > gpurun_out/r2q/fuzz_parity.json 2> gpurun_out/r2q/fuzz.err; tail -3 gpurun_out/r2q/fuzz.err | cut -c1-300; head -8 gpurun_out/r2q/fuzz_parity.json
for v in fb7 fb8; do EOGS_RASTER_LIB=$PWD/eogs2_b200/libeogs_raster_$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-config5 > gpurun_out/r2q/bench_$v.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2q/bench_$v.json')); print('$v', round(d['ms_per_step'],4), d['stage_ms']['blend_fwd'])"; done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-config5 > gpurun_out/r2q/bench_base.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2q/bench_base.json')); print('base', round(d['ms_per_step'],4), d['stage_ms']['blend_fwd'])"
timeout 600 python tools/bench_configs.py --configs 1 2>/dev/null | tail -6 | cut -c1-300
