"""distCUDA2 timing: this repo's eogs_knn_dist2 vs the compiled reference (oracle/_ref/libknn_ref.so) on the
same points.  CUDA events for ours (stream-ordered); wall clock around the reference (it synchronises itself).
    python tools/bench_knn.py > gpurun_out/<tag>/knn.json"""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from knn_cases import points                      # noqa: E402
from eogs2_b200.simple_knn import distCUDA2      # noqa: E402
from oracle import ref_knn                        # noqa: E402

rows = []
for kind, P in (("uniform", 100_000), ("uniform", 1_000_000), ("uniform", 5_000_000), ("clustered", 1_000_000)):
    p = torch.from_numpy(points(kind, P, 1337)).cuda()
    for _ in range(3):
        d = distCUDA2(p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        d = distCUDA2(p)
    e1.record(); torch.cuda.synchronize()
    ours_ms = e0.elapsed_time(e1) / reps
    row = {"kind": kind, "P": P, "ours_ms": round(ours_ms, 4)}
    if ref_knn.available():
        r = ref_knn.distCUDA2(p)
        t = time.perf_counter()
        for _ in range(3):
            r = ref_knn.distCUDA2(p)
        torch.cuda.synchronize()
        row["ref_ms"] = round((time.perf_counter() - t) / 3 * 1e3, 4)
        row["bit_exact"] = bool(torch.equal(r.view(torch.int32), d.view(torch.int32)))
        row["speedup"] = round(row["ref_ms"] / ours_ms, 2)
    rows.append(row)
    print(row, file=sys.stderr)
print(json.dumps(rows, indent=1))
