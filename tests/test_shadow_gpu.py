"""GPU: the fused virtual-camera resample (eogs2_b200/shadow.py, csrc/resample.cu) against the torch
sequence of render_resample_virtual_camera (gaussian_renderer/renderer_cc_shadow.py:32-46):
einsum + F.grid_sample(align_corners=True) + channel split + -100 overwrite, forward and autograd.
Bars: values 1e-5 max-abs, gradients 1e-4 relative (float-atomic order)."""
import numpy as np
import pytest
import torch

from eogs2_b200 import shadow as SH

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def torch_reference(virtual_render, cam2virt, rendered_uva):
    virtual_uv = torch.einsum("...ij,...j->...i", cam2virt, rendered_uva)[..., :2]
    sample = torch.nn.functional.grid_sample(virtual_render.unsqueeze(0), virtual_uv.unsqueeze(0),
                                             align_corners=True).squeeze(0)
    rgb = sample[:3]
    alt = sample[3]
    alt[(virtual_uv.abs() > 1).any(-1)] = -100
    return rgb, alt, virtual_uv


def make_inputs(dev, H, W, f, seed, spill):
    g = torch.Generator().manual_seed(seed)
    virt = torch.randn(5, f * H, f * W, generator=g)
    u, v = torch.meshgrid(torch.linspace(-1, 1, W), torch.linspace(-1, 1, H), indexing="xy")    # AffineCamera.UV_grid
    alt = torch.rand(H, W, generator=g) * 60 - 20
    uva = torch.stack([u, v, alt], -1)
    # camera_to_sun-like shear (u, v displaced by altitude) times diag(1/f, 1/f, 1); `spill` pushes part of the
    # footprint outside [-1, 1] so that zero padding and the -100 overwrite are exercised
    M = torch.tensor([[spill / f, 0.02 / f, -0.003 / f], [-0.01 / f, spill / f, -0.0025 / f], [0.0, 0.0, 1.0]])
    return [t.to(dev).requires_grad_(True) for t in (virt, M, uva)]


@pytest.mark.parametrize("H,W,f,seed,spill", [(96, 128, 2, 1, 1.0), (75, 53, 2, 2, 2.6), (64, 64, 1, 3, 0.7),
                                                 (5, 123, 3, 9609, 1.7), (117, 191, 3, 9282, 4.5)])
def test_resample_matches_torch_grid_sample(cuda_dev, H, W, f, seed, spill):
    a = make_inputs(cuda_dev, H, W, f, seed, spill)
    b = make_inputs(cuda_dev, H, W, f, seed, spill)
    rgb_r, alt_r, uv_r = torch_reference(*a)
    rgb, alt, uv = SH.resample_virtual(*b)
    assert rgb.shape == (3, H, W) and alt.shape == (H, W) and uv.shape == (H, W, 2)
    assert float((uv - uv_r).detach().abs().max()) <= 1e-5
    assert float((rgb - rgb_r).detach().abs().max()) <= 2e-5 and float((alt - alt_r).detach().abs().max()) <= 2e-4
    outside = (uv_r.abs() > 1).any(-1)
    assert torch.equal(alt[outside], torch.full_like(alt[outside], -100.0))
    if spill / f > 1.2:                                   # the footprint scale is spill / f
        assert 0.2 < outside.float().mean().item() < 0.98
    g = torch.Generator().manual_seed(seed + 10)
    w_rgb, w_alt, w_uv = (torch.randn(*s, generator=g).to(cuda_dev) for s in ((3, H, W), (H, W), (H, W, 2)))
    ((rgb_r * w_rgb).sum() + (alt_r * w_alt).sum() + (uv_r * w_uv).sum()).backward()
    ((rgb * w_rgb).sum() + (alt * w_alt).sum() + (uv * w_uv).sum()).backward()
    for name, x, y in zip(("virtual_render", "cam2virt", "rendered_uva"), b, a):
        assert rel(x.grad, y.grad) < 1e-4, (name, rel(x.grad, y.grad))


def test_resample_without_some_upstream_gradients(cuda_dev):
    a = make_inputs(cuda_dev, 40, 48, 2, 5, 1.0)
    b = make_inputs(cuda_dev, 40, 48, 2, 5, 1.0)
    torch_reference(*a)[1].sum().backward()             # altitude only (sun_altitude_diff path of train_pan.py:319)
    SH.resample_virtual(*b)[1].sum().backward()
    for x, y in zip(b, a):
        assert rel(x.grad, y.grad) < 1e-4
