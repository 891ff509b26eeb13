#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the multi-GPU tests (NVLS kernel vs NCCL, bands), then the bench under torchrun.
# Usage: bash tools/gpu_multi.sh <tag> <N>
TAG=$1; N=$2
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi.txt 2>&1
timeout 600 python -m pytest tests/test_nvls_gpu.py tests/test_bands_gpu.py -q -x > $OUT/pytest_multi.log 2>&1; tail -3 $OUT/pytest_multi.log | cut -c1-300
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_n$N.json").read().strip().splitlines()[-1])
    print("value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", d["e2e"], "iter_ms", round(d["iter_ms"],3))
    print("allreduce", d["config"].get("allreduce"))
    print("dp", json.dumps(d.get("dp"))[:1500])
    print("config5", d.get("config5"))
    print("clocks", d.get("clocks"))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_n$N.err").read()[-3000:])
PY
