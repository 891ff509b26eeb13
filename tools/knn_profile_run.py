"""Workload for the ncu captures of distCUDA2: 1 M uniform points, a few calls.  (tools/gpu_r03c.sh)"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from knn_cases import points                      # noqa: E402
from eogs2_b200.simple_knn import distCUDA2      # noqa: E402

p = torch.from_numpy(points("uniform", 1_000_000, 1337)).cuda()
for _ in range(3):
    d = distCUDA2(p)
torch.cuda.synchronize()
print(float(d.mean()))
