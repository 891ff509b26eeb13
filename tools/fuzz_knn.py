"""Randomised distCUDA2 sweep against the COMPILED REFERENCE (oracle/_ref/libknn_ref.so): random sizes and point
distributions, bit-exact.  Development tool:  python tools/fuzz_knn.py --cases 300 > gpurun_out/<tag>/fuzz_knn.json"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from knn_cases import points                      # noqa: E402
from eogs2_b200.simple_knn import distCUDA2      # noqa: E402
from oracle import ref_knn                        # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cases", type=int, default=300)
ap.add_argument("--seed", type=int, default=0)
a = ap.parse_args()
rng = np.random.default_rng(a.seed)
rows, nbad = [], 0
for i in range(a.cases):
    kind = str(rng.choice(["uniform", "clustered", "planar", "offset", "line"]))
    P = int(rng.choice([1, 2, 3, 4, 5, 31, 32, 33, 1023, 1024, 1025, int(rng.integers(6, 200_000))]))
    if kind == "clustered":
        P = max(P, 20)
    seed = int(rng.integers(0, 100_000))
    p = points(kind, P, seed)
    if rng.random() < 0.3 and P > 4:                      # heavy duplication
        idx = rng.integers(0, P, P // 2)
        p[rng.integers(0, P, P // 2)] = p[idx]
    t = torch.from_numpy(p).cuda()
    mine, ref = distCUDA2(t), ref_knn.distCUDA2(t)
    bad = int((mine.view(torch.int32) != ref.view(torch.int32)).sum())
    nbad += bad > 0
    rows.append(dict(case=i, kind=kind, P=P, seed=seed, mismatching_values=bad))
    if bad:
        print("MISMATCH", rows[-1], file=sys.stderr)
print(json.dumps({"cases": len(rows), "mismatching_cases": nbad, "rows": rows}))
print(f"fuzz_knn: {len(rows)} cases, {nbad} with mismatches", file=sys.stderr)
