"""Randomised sweep of the stages around the rasterizer: re-runs the parametrised GPU parity tests (tile bands,
virtual-camera resample, photometric loss, fused render glue) with random sizes / seeds / options and counts
assertion failures.  Development tool:   python tools/fuzz_misc.py --cases 40 > gpurun_out/<tag>/fuzz_misc.json"""
import argparse
import json
import sys
import traceback
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import test_bands_gpu as TB        # noqa: E402
import test_fused_gpu as TF        # noqa: E402
import test_losses_gpu as TL       # noqa: E402
import test_shadow_gpu as TS       # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cases", type=int, default=40)
ap.add_argument("--seed", type=int, default=0)
a = ap.parse_args()
rng = np.random.default_rng(a.seed)
dev = torch.device("cuda:0")
fails, counts = [], {}


def run(name, fn, **kw):
    counts[name] = counts.get(name, 0) + 1
    try:
        fn(dev, **kw)
    except Exception as e:                                   # noqa: BLE001
        fails.append({"test": name, "args": {k: (v if isinstance(v, (int, float, bool)) else str(v)) for k, v in kw.items()},
                      "error": "".join(traceback.format_exception_only(type(e), e)).strip()[:400]})
        print("FAIL", fails[-1], file=sys.stderr)


for i in range(a.cases):
    Hb = int(rng.integers(17, 700))
    run("bands", TB.test_bands_reproduce_the_whole_image, P=int(rng.choice([1, 50, 3000, 40000])),
        W=int(rng.integers(1, 700)), H=Hb, seed=int(rng.integers(0, 9999)),
        world=int(rng.integers(1, min(8, (Hb + 15) // 16) + 1)), weighted=bool(rng.random() < 0.5))
    run("resample", TS.test_resample_matches_torch_grid_sample, H=int(rng.integers(2, 300)), W=int(rng.integers(2, 300)),
        f=int(rng.choice([1, 2, 3])), seed=int(rng.integers(0, 9999)), spill=float(rng.choice([0.5, 1.0, 1.7, 3.0, 5.0])))
    run("loss", TL.test_photometric_loss_matches_torch, C=int(rng.choice([1, 3])), H=int(rng.integers(1, 400)),
        W=int(rng.integers(1, 400)), lam=float(rng.choice([0.0, 0.2, 1.0])), seed=int(rng.integers(0, 9999)))
    run("fused", TF.test_fused_render_matches_the_reference_sequence, P=int(rng.choice([1, 40, 5000, 30000])),
        W=int(rng.integers(1, 500)), H=int(rng.integers(1, 500)), seed=int(rng.integers(0, 9999)),
        aa=bool(rng.random() < 0.3), learn_last=bool(rng.random() < 0.5), mod=float(rng.choice([1.0, 0.7, 1.3])))
print(json.dumps({"runs": counts, "failures": fails}, indent=1))
print(f"fuzz_misc: {sum(counts.values())} runs, {len(fails)} failures", file=sys.stderr)
