"""One EOGS++ training iteration for one camera on the fused kernels.

The per-camera block of `train_pan.training` (train_pan.py:268-469), reduced to the parts that touch
the hot path and its neighbours:

    main render            render(cam, gaussians, pipe, bg)                       train_pan.py:278
    sun-view shadow pass   render_resample_virtual_camera(sun_camera, ...)        :305-316
    shading                exp(0.4 * min(altitude - sun_altitude, 0))             affine_cameras.py:33-40,336-341
    photometric loss       (1 - l) * L1 + l * (1 - SSIM)                          :423-465, loss/shadow.py:21-29
    backward + Adam        loss.backward(); optimizer.step()                      :469, :664-670

Every stage runs on this repository's kernels: `fused.render_fused`, `shadow.render_resample_virtual_camera`,
`losses.photometric_loss`, `optim.FlatGaussianAdam`.  Camera-specific colour correction, the random-camera
and regularisation losses and densification stay in the caller (they are torch modules of the scene).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import fused, losses, shadow


def model_view(params: dict) -> SimpleNamespace:
    """The attributes of GaussianModel that render() reads, backed by FlatGaussianAdam.params (or any dict of
    tensors named xyz, f_dc, opacity, scaling, rotation)."""
    return SimpleNamespace(_xyz=params["xyz"], _features_dc=params["f_dc"], _opacity=params["opacity"],
                           _scaling=params["scaling"], _rotation=params["rotation"], active_sh_degree=0)


def camera_iteration(cam, sun_cam, cam2sun: torch.Tensor, pc, pipe, bg: torch.Tensor, gt_image: torch.Tensor,
                     lambda_dssim: float = 0.2, inshadow: float = 0.3, render_fn=None, resample_fn=None, loss_fn=None):
    """Forward of one camera's iteration; returns (loss, dict of intermediates).  Call `.backward()` on the loss
    and step the optimiser.  `cam.UV_grid` is the (u, v) meshgrid of AffineCamera (affine_cameras.py:139-143).
    render_fn / resample_fn / loss_fn default to the fused kernels; pass the reference's functions to compare."""
    render_fn = render_fn or fused.render_fused
    resample_fn = resample_fn or shadow.render_resample_virtual_camera
    loss_fn = loss_fn or losses.photometric_loss
    pkg = render_fn(cam, pc, pipe, bg)
    raw_render, altitude_render = pkg["render"][:3], pkg["render"][3]
    rendered_uva = torch.stack(tuple(cam.UV_grid) + (altitude_render,), dim=-1)            # train_pan.py:282
    sun_rgb, sun_altitude, sun_uv = resample_fn(sun_cam, cam2sun, rendered_uva, pc, pipe, bg)
    sun_altitude_diff = altitude_render - sun_altitude                                     # train_pan.py:319
    shadow_map = torch.exp(0.4 * sun_altitude_diff.clip(max=0.0))                          # ShadowMap.forward
    shaded = shadow_map * raw_render + (1 - shadow_map) * inshadow * raw_render            # render_pipeline :336-341
    loss = loss_fn(shaded, gt_image, lambda_dssim)
    return loss, dict(render=pkg["render"], viewspace_points=pkg["viewspace_points"], shaded=shaded,
                      shadow=shadow_map, sun_uv=sun_uv, sun_rgb=sun_rgb)
