"""Generates tests/golden/knn_ref.npz by running the REFERENCE simple-knn (oracle/_ref/libknn_ref.so, built
from /root/reference by oracle/ref_build/Makefile) on a B200:

    gpurun -- python tests/golden/make_golden_knn.py     # writes gpurun_out/golden/knn_ref.npz
    cp gpurun_out/golden/knn_ref.npz tests/golden/

The reference ships no test of distCUDA2; this fixture pins oracle_dist2 (oracle/eogs_oracle.c, CPU test,
bit-exact) and the CUDA path (GPU test, bit-exact).  Inputs are stored next to the outputs.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from knn_cases import GOLDEN_CASES, points            # noqa: E402
from oracle import ref_knn                            # noqa: E402


def main():
    out_dir = ROOT / "gpurun_out" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    save = {}
    for name, (kind, P, seed) in GOLDEN_CASES.items():
        p = points(kind, P, seed)
        d = ref_knn.distCUDA2(torch.from_numpy(p).cuda()).cpu().numpy()
        save[f"{name}_points"], save[f"{name}_dist2"] = p, d
        print(name, P, d[:3])
    np.savez_compressed(out_dir / "knn_ref.npz", **save)


if __name__ == "__main__":
    main()
