"""GPU: the fused render glue (eogs2_b200/fused.py) against the reference's own
sequence — torch activations + colors_precomp + GaussianRasterizer (gaussian_renderer/renderer.py:
27-144) — on the same raw GaussianModel parameters.

Bars: images max-abs <= 1e-4 (the altitude colour is an fma chain in the kernel and a cuBLAS matmul in
torch: last-bit differences); radii equal except for a <= 1e-4 fraction (F.normalize's reduction order
is torch-internal); every gradient w.r.t. the raw parameters, the camera matrix and the altitude
affine row <= 1e-3 relative."""
import math
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from eogs2_b200 import fused as F
from eogs2_b200 import scene as S

pytestmark = pytest.mark.gpu
C0 = 0.28209479177387814


def rel(a, b):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


class FakeModel:
    """The slice of GaussianModel that render() touches (scene/gaussian_model.py:41-53,109-137)."""

    def __init__(self, dev, P, seed):
        sc = S.make_scene(P, "trained", seed)
        g = torch.Generator().manual_seed(seed + 1)
        self._xyz = sc.means3D.to(dev).requires_grad_(True)
        self._features_dc = ((sc.rgb - 0.5) / C0).unsqueeze(1).to(dev).requires_grad_(True)          # [P,1,3]
        self._opacity = torch.logit(sc.opacities).to(dev).requires_grad_(True)                        # [P,1]
        self._scaling = torch.log(sc.scales).to(dev).requires_grad_(True)
        self._rotation = (sc.rotations * (0.5 + torch.rand(P, 1, generator=g))).to(dev).requires_grad_(True)
        self.active_sh_degree = 0

    def params(self):
        return [self._xyz, self._features_dc, self._opacity, self._scaling, self._rotation]

    get_xyz = property(lambda s: s._xyz)
    get_scaling = property(lambda s: torch.exp(s._scaling))
    get_rotation = property(lambda s: torch.nn.functional.normalize(s._rotation))
    get_opacity = property(lambda s: torch.sigmoid(s._opacity))


class FakeCamera:
    def __init__(self, dev, seed, W, H, learn_last=False):
        v = S.make_camera(seed).to(dev)
        self.world_view_transform = v.clone().requires_grad_(True)
        self.full_proj_transform = self.world_view_transform
        self.affine = v.clone().requires_grad_(True)              # unrefined affine (ECEF_to_UVA uses it)
        self.image_width, self.image_height = W, H
        self.FoVx = self.FoVy = 0.5
        self.camera_center = torch.zeros(3, device=dev)
        self.learn_wv_only_lastparam = learn_last
        self.last_row = (torch.tensor([0.01, -0.02, 0.0, 0.0], device=dev)).requires_grad_(True)
        self.image_name = "synthetic"

    def ECEF_to_UVA(self, xyz):                                    # affine_cameras.py:432-438: xyz @ At + bt
        # written as broadcast multiplies instead of `@`: the same affine map without depending on which cuBLAS
        # kernel (and internal precision) a [P,3]x[3,3] matmul happens to get on the box
        At, bt = self.affine[:3, :3], self.affine[3, :3]
        return xyz[:, 0:1] * At[0] + xyz[:, 1:2] * At[1] + xyz[:, 2:3] * At[2] + bt

    def leaves(self):
        return [self.world_view_transform, self.affine, self.last_row]


def reference_render(cam, pc, pipe, bg, mod=1.0):
    """renderer.py:27-144, line by line, on top of the (unfused) drop-in rasterizer."""
    screenspace_points = torch.zeros_like(pc.get_xyz, requires_grad=True) + 0
    screenspace_points.retain_grad()
    viewmatrix, projmatrix = cam.world_view_transform, cam.full_proj_transform
    if cam.learn_wv_only_lastparam:
        viewmatrix = cam.world_view_transform.clone()
        projmatrix = cam.full_proj_transform.clone()
        viewmatrix[-1, :] = viewmatrix[-1, :] + cam.last_row
        projmatrix[-1, :] = projmatrix[-1, :] + cam.last_row
    rs = GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=math.tan(cam.FoVx * 0.5),
        tanfovy=math.tan(cam.FoVy * 0.5), bg=bg, scale_modifier=mod, viewmatrix=viewmatrix, projmatrix=projmatrix,
        sh_degree=0, campos=cam.camera_center, prefiltered=False, debug=pipe.debug, antialiasing=pipe.antialiasing)
    rgb = (pc._features_dc * C0 + 0.5).squeeze(1)
    altitude = cam.ECEF_to_UVA(pc._xyz)[..., 2].unsqueeze(-1)
    colors = torch.cat([rgb, altitude, torch.ones_like(altitude)], dim=-1)
    img, radii, _ = GaussianRasterizer(rs)(means3D=pc.get_xyz, means2D=screenspace_points, shs=None,
                                           colors_precomp=colors, opacities=pc.get_opacity,
                                           scales=pc.get_scaling, rotations=pc.get_rotation, cov3D_precomp=None)
    return {"render": img, "viewspace_points": screenspace_points, "radii": radii,
            "visibility_filter": (radii > 0).nonzero()}


@pytest.mark.parametrize("P,W,H,seed,aa,learn_last,mod", [
    (30_000, 320, 240, 5, False, False, 1.0),
    (20_000, 257, 190, 6, True, True, 1.0),          # antialiasing + learn_wv_only_lastparam
    (10_000, 128, 128, 7, False, False, 0.8),         # scaling_modifier
])
def test_fused_render_matches_the_reference_sequence(cuda_dev, P, W, H, seed, aa, learn_last, mod):
    dev = cuda_dev
    pipe = SimpleNamespace(debug=False, antialiasing=aa, compute_cov3D_python=False, require_radii=True)
    bg = S.background(seed).to(dev)
    dcol = S.upstream_grads(5, H, W, seed, False)[0].to(dev)

    def run(fn):
        pc, cam = FakeModel(dev, P, seed), FakeCamera(dev, seed, W, H, learn_last)
        out = fn(cam, pc, pipe, bg, mod)
        (out["render"] * dcol).sum().backward()
        return out, pc, cam
    ref, pc_r, cam_r = run(reference_render)
    got, pc_f, cam_f = run(lambda c, p, pi, b, m: F.render_fused(c, p, pi, b, m))

    assert got["render"].shape == (5, H, W)
    assert (got["radii"] != ref["radii"]).float().mean().item() <= 1e-4
    err = (got["render"] - ref["render"]).abs()
    assert (err > 1e-4).float().mean().item() <= 1e-4, [float(err[c].max()) for c in range(5)]
    assert got["visibility_filter"].shape[1] == 1
    for name, a, b in zip(("xyz", "features_dc", "opacity", "scaling", "rotation"), pc_f.params(), pc_r.params()):
        assert a.grad is not None and a.grad.shape == b.grad.shape, name
        assert rel(a.grad, b.grad) < 1e-3, (name, rel(a.grad, b.grad))
    assert rel(got["viewspace_points"].grad, ref["viewspace_points"].grad) < 1e-3
    for name, a, b in zip(("world_view_transform", "affine", "last_row"), cam_f.leaves(), cam_r.leaves()):
        if b.grad is None:
            assert a.grad is None or float(a.grad.abs().max()) == 0.0, name
        else:
            assert rel(a.grad, b.grad) < 1e-3, (name, rel(a.grad, b.grad))


def test_fused_render_defers_to_the_unfused_path_for_override_color(cuda_dev):
    dev = cuda_dev
    P, W, H = 5000, 96, 80
    pipe = SimpleNamespace(debug=False, antialiasing=False, compute_cov3D_python=False, require_radii=False)
    pc, cam = FakeModel(dev, P, 9), FakeCamera(dev, 9, W, H)
    bg = S.background(9).to(dev)
    override = torch.rand(P, 5, device=dev)
    out = F.render_fused(cam, pc, pipe, bg, override_color=override)
    rs = GaussianRasterizationSettings(H, W, math.tan(0.25), math.tan(0.25), bg, 1.0, cam.world_view_transform,
                                       cam.full_proj_transform, 0, cam.camera_center, False, False, False)
    img, _, _ = GaussianRasterizer(rs)(means3D=pc.get_xyz, means2D=torch.zeros_like(pc.get_xyz), opacities=pc.get_opacity,
                                       colors_precomp=override, scales=pc.get_scaling, rotations=pc.get_rotation)
    assert torch.equal(out["render"], img) and "radii" not in out
