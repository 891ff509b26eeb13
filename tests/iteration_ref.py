"""Test / bench helper (not product code): the torch formulation of the iteration stages, as the reference
runs them, and a synthetic camera pair.  Imported by tests/test_iteration_gpu.py and tools/bench_configs.py."""
import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from eogs2_b200 import losses as L
from eogs2_b200 import scene as S

C0 = 0.28209479177387814
SUN_Q = (-0.0030, -0.0025)


def make_cameras(dev, seed, W, H):
    view = S.make_camera(seed).to(dev)
    u, v = torch.meshgrid(torch.linspace(-1, 1, W, device=dev), torch.linspace(-1, 1, H, device=dev), indexing="xy")
    base = dict(FoVx=0.5, FoVy=0.5, camera_center=torch.zeros(3, device=dev), learn_wv_only_lastparam=False,
                image_name="synthetic")
    cam = SimpleNamespace(world_view_transform=view, full_proj_transform=view, affine=view, image_width=W, image_height=H,
                          UV_grid=(u, v), **base)
    sun_view = S.sun_camera(view.cpu(), SUN_Q, 2).to(dev)
    # the sun camera's altitude colour uses ITS affine (third column is unchanged by the shear and the 1/f scaling)
    sun = SimpleNamespace(world_view_transform=sun_view, full_proj_transform=sun_view, affine=sun_view,
                          image_width=2 * W, image_height=2 * H, **base)
    c2s = torch.eye(3, device=dev)
    c2s[0, 2], c2s[1, 2] = SUN_Q
    cam2sun = torch.diag(torch.tensor([0.5, 0.5, 1.0], device=dev)) @ c2s            # get_sun_camera, affine_cameras.py:350-370
    for c in (cam, sun):
        c.ECEF_to_UVA = (lambda xyz, a=c.affine: xyz[:, 0:1] * a[0, :3] + xyz[:, 1:2] * a[1, :3] + xyz[:, 2:3] * a[2, :3] + a[3, :3])
    return cam, sun, cam2sun


def torch_render(cam, pc, pipe, bg, scaling_modifier=1.0):
    """gaussian_renderer/renderer.py:27-144 on the (unfused) drop-in rasterizer."""
    sp = torch.zeros_like(pc._xyz, requires_grad=True) + 0
    sp.retain_grad()
    rs = GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=math.tan(0.25), tanfovy=math.tan(0.25), bg=bg,
        scale_modifier=scaling_modifier, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
        sh_degree=0, campos=cam.camera_center, prefiltered=False, debug=False, antialiasing=False)
    rgb = (pc._features_dc * C0 + 0.5).squeeze(1)
    altitude = cam.ECEF_to_UVA(pc._xyz)[..., 2].unsqueeze(-1)
    colors = torch.cat([rgb, altitude, torch.ones_like(altitude)], dim=-1)
    img, radii, _ = GaussianRasterizer(rs)(
        means3D=pc._xyz, means2D=sp, colors_precomp=colors, opacities=torch.sigmoid(pc._opacity),
        scales=torch.exp(pc._scaling), rotations=F.normalize(pc._rotation))
    return {"render": img, "viewspace_points": sp, "radii": radii}


def torch_resample(virtual_camera, cam2virt, rendered_uva, gaussians, pipe, background):
    """gaussian_renderer/renderer_cc_shadow.py:6-54."""
    virtual_render = torch_render(virtual_camera, gaussians, pipe, background)["render"]
    uv = torch.einsum("...ij,...j->...i", cam2virt, rendered_uva)[..., :2]
    smp = F.grid_sample(virtual_render.unsqueeze(0), uv.unsqueeze(0), align_corners=True).squeeze(0)
    rgb, alt = smp[:3], smp[3]
    alt[(uv.abs() > 1).any(-1)] = -100
    return rgb, alt, uv


def torch_photometric(image, gt, lam):
    """loss/shadow.py:21-29 with utils/loss_utils.py:18-85."""
    ch = image.size(-3)
    w1 = L.gaussian_window().unsqueeze(1)
    window = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0).expand(ch, 1, 11, 11).contiguous().to(image.device)
    conv = lambda x: F.conv2d(x, window, padding=5, groups=ch)
    mu1, mu2 = conv(image), conv(gt)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1, s2, s12 = conv(image * image) - mu1_sq, conv(gt * gt) - mu2_sq, conv(image * gt) - mu12
    ssim_map = ((2 * mu12 + 0.01 ** 2) * (2 * s12 + 0.03 ** 2)) / ((mu1_sq + mu2_sq + 0.01 ** 2) * (s1 + s2 + 0.03 ** 2))
    return (1.0 - lam) * torch.abs(image - gt).mean() + lam * (1.0 - ssim_map.mean())


def raw_params(dev, P, seed):
    sc = S.make_scene(P, "trained", seed)
    return {"xyz": sc.means3D.to(dev), "f_dc": ((sc.rgb - 0.5) / C0).unsqueeze(1).to(dev),
            "opacity": torch.logit(sc.opacities).to(dev), "scaling": torch.log(sc.scales).to(dev),
            "rotation": (sc.rotations * 1.3).to(dev)}
