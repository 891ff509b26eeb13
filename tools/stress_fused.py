"""Developer probe (GPU): hunt for a rare mismatch between the fused-parameter path and the unfused path."""
import sys
from pathlib import Path
from types import SimpleNamespace
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import eogs2_b200 as E
import test_fused_gpu as T
from eogs2_b200 import fused as F, scene as S
dev = torch.device("cuda:0")
P, W, H, seed = 30000, 320, 240, 5
pipe = SimpleNamespace(debug=False, antialiasing=False, compute_cov3D_python=False, require_radii=True)
bg = S.background(seed).to(dev)
dcol = S.upstream_grads(5, H, W, seed, False)[0].to(dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
bad = 0
base_ref = base_fused = None
for it in range(N):
    pc, cam = T.FakeModel(dev, P, seed), T.FakeCamera(dev, seed, W, H, False)
    ref = T.reference_render(cam, pc, pipe, bg, 1.0)
    (ref["render"] * dcol).sum().backward()
    pc2, cam2 = T.FakeModel(dev, P, seed), T.FakeCamera(dev, seed, W, H, False)
    got = F.render_fused(cam2, pc2, pipe, bg, 1.0)
    (got["render"] * dcol).sum().backward()
    r, g = ref["render"].detach(), got["render"].detach()
    if base_ref is None:
        base_ref, base_fused = r.clone(), g.clone()
    e_rf = (r - g).abs()
    nb = int((e_rf > 1e-4).sum())
    d_ref = int((r != base_ref).sum()); d_fus = int((g != base_fused).sum())
    if nb or d_ref or d_fus:
        bad += 1
        ys, xs = torch.nonzero((e_rf > 1e-4).any(0), as_tuple=True)
        tiles = sorted(set(((ys // 16) * 20 + xs // 16).tolist()))
        print(f"it {it}: ref-vs-fused bad values {nb} per-channel max {[float(e_rf[c].max()) for c in range(5)]} "
              f"ref changed vs run 0: {d_ref} fused changed vs run 0: {d_fus} tiles {tiles[:20]}", flush=True)
print(f"done: {bad} bad iterations of {N}")
