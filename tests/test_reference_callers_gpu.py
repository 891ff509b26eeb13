"""GPU: SURVEY.md section 8 rows a19 / N1 pinned by the REFERENCE's own Python callers.

tests/golden/renderer_ref.npz was produced in the build container by importing, from /root/reference,
gaussian_renderer/renderer.py `render`, renderer_cc_shadow.py `render_resample_virtual_camera`,
scene/gaussian_model.py `GaussianModel` and scene/cameras/affine_cameras.py `AffineCamera` (the real code,
tests/golden/make_golden_renderer.py + ref_import.py; CPU stand-in rasterizer = oracle/eogs_oracle.c, forward and the reference's hand-written backward).  Here the
same raw GaussianModel parameters and cameras go through

  (a) the callers' torch sequence on the DROP-IN package `diff_gaussian_rasterization` (what a user of the
      reference gets after switching: renderer.py's own lines, restated in tests/test_fused_gpu.py /
      tests/iteration_ref.py because /root/reference does not exist on the GPU box), and
  (b) the fused kernels (eogs2_b200/fused.py, shadow.py),

and both must reproduce what the reference's code produced.  Bars: images max-abs 1e-4 except a <= 2e-3 fraction of
pixels (one alpha-threshold flip between CUDA expf and the CPU's exp moves a pixel by <= |colour| / 255), radii equal,
gradients w.r.t. every raw parameter within the per-Gaussian bar of tests/parity_util.py (1e-3 of the Gaussian's
gradient vector) for ALL Gaussians except the few that own a flipped pixel (a flip adds or removes one whole term of a
small Gaussian's gradient), and 1e-3 in L2 over the rest.
"""
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as TF

from eogs2_b200 import fused as F
from eogs2_b200 import shadow as SH
from parity_util import violations
from test_fused_gpu import reference_render

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "renderer_ref.npz"


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def rel(a, b):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


class Model:
    """The slice of GaussianModel that render() touches, on the fixture's raw parameters."""

    def __init__(self, dev, g):
        t = lambda k: torch.from_numpy(g["raw_" + k]).to(dev).requires_grad_(True)
        self._xyz, self._features_dc, self._opacity = t("xyz"), t("f_dc"), t("opacity")
        self._scaling, self._rotation = t("scaling"), t("rotation")
        self.active_sh_degree = 0

    get_xyz = property(lambda s: s._xyz)
    get_scaling = property(lambda s: torch.exp(s._scaling))
    get_rotation = property(lambda s: TF.normalize(s._rotation))
    get_opacity = property(lambda s: torch.sigmoid(s._opacity))

    def get_covariance(self, scaling_modifier=1):            # scene/gaussian_model.py:33-38,149-152
        s, q = scaling_modifier * self.get_scaling, self._rotation / self._rotation.norm(dim=1, keepdim=True)
        r, x, y, z = q.unbind(-1)
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                         2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                         2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).view(-1, 3, 3)
        L = R * s[:, None, :]
        cov = L @ L.transpose(1, 2)
        return torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], -1)

    def grads(self):
        return {"xyz": self._xyz.grad, "f_dc": self._features_dc.grad, "opacity": self._opacity.grad,
                "scaling": self._scaling.grad, "rotation": self._rotation.grad}


def affine_t(dev, coef, inter):
    """AffineCamera.affine (affine_cameras.py:146-153): [[A, b], [0, 1]] transposed."""
    m = torch.eye(4)
    m[:3, :3] = torch.from_numpy(coef)
    m[:3, 3] = torch.from_numpy(inter)
    return m.t().contiguous().to(dev)


class Camera:
    def __init__(self, dev, view, W, H, learn_last=False, last_row=None):
        self.world_view_transform = self.full_proj_transform = self.affine = view
        self.image_width, self.image_height = W, H
        self.FoVx = self.FoVy = 1
        self.camera_center = torch.zeros(3, device=dev)
        self.learn_wv_only_lastparam = learn_last
        self.last_row = None if last_row is None else torch.from_numpy(last_row).to(dev).requires_grad_(True)
        self.image_name = "synthetic"
        self.UV_grid = torch.meshgrid(torch.linspace(-1, 1, W, device=dev), torch.linspace(-1, 1, H, device=dev),
                                      indexing="xy")

    def ECEF_to_UVA(self, xyz):                               # affine_cameras.py:432-438
        At, bt = self.affine[:3, :3], self.affine[3, :3]
        return xyz[:, 0:1] * At[0] + xyz[:, 1:2] * At[1] + xyz[:, 2:3] * At[2] + bt


def check_image(got, want, what):
    """Returns the number of pixels that differ by more than the bar (accept-decision flips)."""
    err = (got.detach().cpu().numpy() - want)
    bad = np.abs(err) > 1e-4 * np.maximum(1.0, np.abs(want))
    assert bad.mean() <= 2e-3, (what, float(np.abs(err).max()), float(bad.mean()))
    return int(bad.reshape(bad.shape[0], -1).any(0).sum()) if bad.ndim == 3 else int(bad.sum())


def check_grad(got, want, what, flips, l2=1e-3):
    """Per-Gaussian bar for every Gaussian but those owning a flipped pixel; L2 bar over the others."""
    P = want.shape[0]
    n, worst, bad = violations(got.reshape(P, -1), want.reshape(P, -1), 1e-3, 1e-5)
    assert n <= 4 * flips, f"{what}: {n} Gaussians over the per-element bar with {flips} flipped pixels (worst {worst:.2e})"
    keep = np.ones(P, bool); keep[bad] = False
    a = got.detach().double().cpu().numpy().reshape(P, -1)[keep]
    b = np.asarray(want, np.float64).reshape(P, -1)[keep]
    r = np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)
    assert r < l2, (what, r)


CASES = {"main": dict(), "aa_lastrow": dict(aa=True, learn_last=True), "mod": dict(mod=0.8), "cov3d": dict(cov3d=True)}


@pytest.mark.parametrize("tag", list(CASES))
@pytest.mark.parametrize("path", ["drop_in", "fused"])
def test_render_matches_the_reference_renderer(cuda_dev, gold, tag, path):
    dev, g, kw = cuda_dev, gold, CASES[tag]
    W, H = int(g["W"]), int(g["H"])
    pc = Model(dev, g)
    cam = Camera(dev, affine_t(dev, g["cam_affine_coef"], g["cam_affine_inter"]), W, H, kw.get("learn_last", False),
                 g.get(f"{tag}_last_row"))
    pipe = SimpleNamespace(debug=False, antialiasing=kw.get("aa", False), compute_cov3D_python=kw.get("cov3d", False),
                           require_radii=True)
    bg, dcol = torch.from_numpy(g["bg"]).to(dev), torch.from_numpy(g["dcol"]).to(dev)
    mod = kw.get("mod", 1.0)
    if path == "fused":
        out = F.render_fused(cam, pc, pipe, bg, mod)
    elif kw.get("cov3d"):
        out = F.render_fused(cam, pc, pipe, bg, mod)          # defers to renderer.py:78-83's sequence on the drop-in
    else:
        out = reference_render(cam, pc, pipe, bg, mod)
    (out["render"] * dcol).sum().backward()

    flips = check_image(out["render"], g[f"{tag}_render"], f"{tag}/{path} render")
    assert (out["radii"].cpu().numpy() != g[f"{tag}_radii"]).mean() <= 2e-3
    assert abs(out["visibility_filter"].shape[0] - g[f"{tag}_visibility_filter"].shape[0]) <= 2
    for name, grad in pc.grads().items():
        check_grad(grad, g[f"{tag}_grad_{name}"], f"{tag}/{path}/{name}", flips)
    check_grad(out["viewspace_points"].grad, g[f"{tag}_grad_viewspace"], f"{tag}/{path}/viewspace", flips)
    if kw.get("learn_last"):
        assert rel(cam.last_row.grad, g[f"{tag}_grad_last_row"]) < 1e-3


@pytest.mark.parametrize("path", ["drop_in", "fused"])
def test_sun_shadow_composition_matches_the_reference_callers(cuda_dev, gold, path):
    """train_pan.py:278-329 for one camera: main render -> sun camera (get_sun_camera) -> resample -> shadow map ->
    render_pipeline's shaded image, and the gradients of a loss on (shaded, sun_rgb) w.r.t. the raw parameters."""
    dev, g = cuda_dev, gold
    W, H = int(g["W"]), int(g["H"])
    pc = Model(dev, g)
    cam = Camera(dev, affine_t(dev, g["cam_affine_coef"], g["cam_affine_inter"]), W, H)
    # AffineCamera.get_sun_camera (affine_cameras.py:350-370): sun_affine @ diag(1/2, 1/2, 1, 1), at 2W x 2H
    scal = torch.diag(torch.tensor([0.5, 0.5, 1.0, 1.0], device=dev))
    sun_view = affine_t(dev, g["cam_sun_affine_coef"], g["cam_sun_affine_inter"]) @ scal
    assert torch.allclose(sun_view.cpu(), torch.from_numpy(g["shadow_sun_view"]), atol=1e-6)
    sun = Camera(dev, sun_view, 2 * W, 2 * H)
    cam2virt = scal[:3, :3] @ torch.from_numpy(g["cam_camera_to_sun"]).to(dev)
    assert torch.allclose(cam2virt.cpu(), torch.from_numpy(g["shadow_cam2virt"]), atol=1e-7)
    pipe = SimpleNamespace(debug=False, antialiasing=False, compute_cov3D_python=False, require_radii=True)
    bg = torch.from_numpy(g["bg"]).to(dev)

    render_fn = F.render_fused if path == "fused" else reference_render
    pkg = render_fn(cam, pc, pipe, bg)
    raw_render, altitude_render = pkg["render"][:3], pkg["render"][3]
    rendered_uva = torch.stack(cam.UV_grid + (altitude_render,), dim=-1)
    if path == "fused":
        sun_rgb, sun_alt, sun_uv, sun_render = SH.render_resample_virtual_camera(sun, cam2virt, rendered_uva, pc, pipe, bg,
                                                                                 return_extra=True)
    else:                                                      # renderer_cc_shadow.py:28-46
        sun_render = reference_render(sun, pc, pipe, bg)["render"]
        sun_uv = torch.einsum("...ij,...j->...i", cam2virt, rendered_uva)[..., :2]
        smp = TF.grid_sample(sun_render.unsqueeze(0), sun_uv.unsqueeze(0), align_corners=True).squeeze(0)
        sun_rgb, sun_alt = smp[:3], smp[3]
        sun_alt[(sun_uv.abs() > 1).any(-1)] = -100
    diff = altitude_render - sun_alt
    shadow = torch.exp(0.4 * diff.clip(max=0.0))              # ShadowMap.forward, affine_cameras.py:33-40
    inshadow = torch.from_numpy(g["shadow_inshadow_cc"]).to(dev)
    shaded = shadow * raw_render + (1 - shadow) * inshadow * raw_render      # render_pipeline :336-341 (identity colour correction)
    loss = (shaded * torch.from_numpy(g["shadow_d_shaded"]).to(dev)).sum() + \
        (sun_rgb * torch.from_numpy(g["shadow_d_sun_rgb"]).to(dev)).sum()
    loss.backward()

    flips = check_image(sun_render, g["shadow_sun_render"], "sun render") + \
        check_image(pkg["render"], g["main_render"], "main render")
    # sun_uv = cam2virt @ (u, v, rendered altitude): a flipped pixel of the main render moves its altitude by <= range / 255
    assert (np.abs(sun_uv.detach().cpu().numpy() - g["shadow_sun_uv"]) > 1e-5).mean() <= 2e-3
    # resampled images: bilinear taps of a render that may hold a few threshold-flip pixels
    for got, want, what in ((sun_rgb, g["shadow_sun_rgb"], "sun rgb"), (sun_alt, g["shadow_sun_alt"], "sun altitude"),
                            (shadow, g["shadow_shadowmap"], "shadow map"), (shaded, g["shadow_shaded"], "shaded")):
        err = np.abs(got.detach().cpu().numpy() - want)
        assert (err > 2e-4 * np.maximum(1.0, np.abs(want))).mean() <= 5e-3, (what, float(err.max()))
    assert abs(float(loss) - float(g["shadow_loss"])) <= 1e-4 * max(1.0, abs(float(g["shadow_loss"])))
    for name, grad in pc.grads().items():
        check_grad(grad, g[f"shadow_grad_{name}"], f"shadow/{path}/{name}", flips, l2=2e-3)
