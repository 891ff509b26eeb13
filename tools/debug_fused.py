import sys
from pathlib import Path
from types import SimpleNamespace
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import test_fused_gpu as T
import test_shadow_gpu as TS
from eogs2_b200 import fused as F, scene as S, shadow as SH
dev = torch.device("cuda:0")
print("tf32 matmul allowed:", torch.backends.cuda.matmul.allow_tf32, torch.get_float32_matmul_precision())
def check(tag):
    P, W, H, seed = 30000, 320, 240, 5
    pipe = SimpleNamespace(debug=False, antialiasing=False, compute_cov3D_python=False, require_radii=True)
    bg = S.background(seed).to(dev)
    pc, cam = T.FakeModel(dev, P, seed), T.FakeCamera(dev, seed, W, H, False)
    ref = T.reference_render(cam, pc, pipe, bg, 1.0)
    got = F.render_fused(cam, pc, pipe, bg, 1.0)
    err = (got["render"] - ref["render"]).abs()
    print(tag, "per-channel max err", [float(err[c].max()) for c in range(5)], "radii mismatch", int((got["radii"] != ref["radii"]).sum()))
    alt_t = cam.ECEF_to_UVA(pc._xyz)[..., 2]
    a = F.altitude_row(cam.affine)
    alt_k = pc._xyz[:, 0] * a[0] + pc._xyz[:, 1] * a[1] + pc._xyz[:, 2] * a[2] + a[3]
    alt_64 = (pc._xyz.double() @ cam.affine[:3, :3].double() + cam.affine[3, :3].double())[..., 2]
    print(tag, "altitude colour: torch matmul vs fp64", float((alt_t - alt_64).abs().max()), " elementwise vs fp64", float((alt_k - alt_64).abs().max()))
check("before")
a = TS.make_inputs(dev, 96, 128, 2, 1, 1.0)
TS.torch_reference(*a)
check("after einsum/grid_sample")
