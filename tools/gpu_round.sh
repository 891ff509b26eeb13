#!/bin/bash
# One GPU-box visit for the committed evidence: parity tests, both bench arms, the ncu launch list and full captures of
# the blend and binning kernels.   Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err
tail -c 400 $OUT/bench_ref.json; echo; tail -c 400 $OUT/bench_ours.json; echo
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-config5 > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'blend_(bwd|fwd)_kernel|bin_scatter_cols_kernel' -s 9 -c 3 \
    -o $OUT/blend -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-config5 > $OUT/ncu_full.log 2>&1
fi
ls -la $OUT
