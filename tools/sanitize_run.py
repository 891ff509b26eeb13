"""Workload for compute-sanitizer (tools/gpu_sanitize.sh): forward + backward of the rasterizer at sizes that exercise
every kernel variant — ragged image borders, tile bands, the small-image tile split of the backward, antialiasing,
precomputed covariances — small enough to finish under memcheck / racecheck / synccheck / initcheck."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import eogs2_b200 as E                      # noqa: E402
from eogs2_b200 import scene as S           # noqa: E402

dev = torch.device("cuda:0")
CASES = [  # P, W, H, kind, antialiasing, band
    (20_000, 300, 200, "trained", False, None),
    (3_000, 64, 48, "init", True, None),
    (12_000, 500, 333, "trained", False, (3, 11)),
    (40_000, 1100, 900, "trained", False, None),      # > 2368 tiles: the backward does not split tiles
    (1, 17, 1, "trained", False, None),
]
for P, W, H, kind, aa, band in CASES:
    sc = S.make_scene(P, kind, 7)
    view = S.make_camera(7)
    d = {k: getattr(sc, k).to(dev) for k in ("means3D", "scales", "rotations", "opacities")}
    colors = S.colors_precomp(sc, view).to(dev)
    view = view.to(dev)
    bg = S.background(7).to(dev)
    empty = torch.empty(0, device=dev)
    st = E.rasterize_forward_raw(bg, d["means3D"], colors, d["opacities"], d["scales"], d["rotations"], 1.0, empty, view,
                                 H, W, aa, False, band=band)
    hb = st.band_height
    dcol, dinv = (t.to(dev) for t in S.upstream_grads(5, hb, W, 7, False))
    g = E.rasterize_backward_raw(st, bg, d["means3D"], colors, d["opacities"], d["scales"], d["rotations"], 1.0, empty,
                                 view, view, dcol, dinv, aa)
    torch.cuda.synchronize()
    assert all(torch.isfinite(t).all() for t in g if t is not None)
    print("ok", P, W, H, kind, aa, band, "instances", st.num_rendered, flush=True)
