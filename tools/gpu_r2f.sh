mkdir -p gpurun_out/r2f
# memcheck on a small forward first (binning is new): fail fast before the full suite
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py -q -x -k "cpu_oracle and 10000" > gpurun_out/r2f/memcheck.log 2>&1; tail -5 gpurun_out/r2f/memcheck.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2f/pytest_gpu.log 2>&1; grep -n "Error\|passed\|failed" gpurun_out/r2f/pytest_gpu.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2f/bench.json 2> gpurun_out/r2f/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2f/bench.json')); print(d['ms_per_step'], d['stage_ms'])" || tail -5 gpurun_out/r2f/bench.err
