"""Fused render glue: EOGS++'s `render()` with the torch glue around the
rasterizer folded into the sm_100a geometry kernels.

The reference's `render()` (gaussian_renderer/renderer.py:27-144) runs, before every rasterizer
call, `exp(_scaling)`, `normalize(_rotation)`, `sigmoid(_opacity)` (scene/gaussian_model.py:41-53,
109-137), `SH2RGB(_features_dc)` (utils/sh_utils.py:125-126), the altitude colour
`ECEF_to_UVA(_xyz)[..., 2]` (scene/cameras/affine_cameras.py:432-438), `ones_like`, `cat`, and a
`zeros_like(...) + 0` for the screen-space gradient slot — ~10 kernels and 5 P-sized temporaries —
and autograd replays their backward after `_C.rasterize_gaussians_backward`.  Here the RAW
parameters of `GaussianModel` go straight into `eogs_forward_geometry_params_band` /
`eogs_backward_params_band` (include/eogs_raster.h): the activations, `colors_precomp` and their
chain rules are computed per Gaussian inside the preprocess kernels.

    render_fused(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, ...)

has `render()`'s signature and returns the same dict ("render", "viewspace_points",
"visibility_filter", "radii"); `rasterize_params(...)` is the tensor-level call.  With
`override_color`, `pipe.compute_cov3D_python` or `use_trained_exp` it defers to the ordinary
(unfused) path through `GaussianRasterizer`, like the reference would.  No CPU path.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _cabi
from .rasterizer import (ERR_ALTITUDE_ABOVE_200, ERR_TOO_MANY_INSTANCES, ForwardState, GaussianRasterizationSettings, GaussianRasterizer,
                         _debug_sync, _f32c, _info_host, _ptr, _quat16, _wait_info, assemble_grad_viewmatrix)

SH_C0 = 0.28209479177387814        # utils/sh_utils.py:25


def forward_params_raw(bg, xyz, features_dc, opacity_logits, log_scales, raw_rotations, alt_affine, scale_modifier,
                       viewmatrix, image_height, image_width, antialiasing=False, debug=False, band=None) -> ForwardState:
    """Forward through the fused-parameter geometry kernel + the ordinary render stage."""
    lib = _cabi.load()
    if xyz.dim() != 2 or xyz.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not xyz.is_cuda:
        raise _cabi.EogsRasterError("xyz must be a CUDA tensor: this rasterizer has no CPU path")
    dev = xyz.device
    P, H, W = int(xyz.size(0)), int(image_height), int(image_width)
    rb, re = (0, (H + 15) // 16) if band is None else (int(band[0]), int(band[1]))
    Hb = min(H, 16 * re) - 16 * rb
    band = None if band is None else (rb, re)
    with torch.cuda.device(dev):
        if P == 0:
            return ForwardState(0, W, H, 5, 0, None, None, None, torch.empty((0,), dtype=torch.int32, device=dev),
                                torch.zeros((5, Hb, W), device=dev), torch.zeros((1, Hb, W), device=dev), band)
        color = torch.empty((5, Hb, W), dtype=torch.float32, device=dev)
        invdepth = torch.empty((1, Hb, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        xyz = _f32c(xyz, "xyz", dev)
        features_dc = _f32c(features_dc.reshape(P, 3), "features_dc", dev)
        opacity_logits = _f32c(opacity_logits, "opacity", dev)
        log_scales = _f32c(log_scales, "scaling", dev)
        raw_rotations = _quat16(_f32c(raw_rotations, "rotation", dev))
        alt_affine = _f32c(alt_affine, "alt_affine", dev)
        viewmatrix = _f32c(viewmatrix, "viewmatrix", dev)
        bg = _f32c(bg, "bg", dev)
        if bg.numel() != 5 or alt_affine.numel() != 4 or opacity_logits.numel() != P:
            raise RuntimeError("fused path: bg must have 5 channels, alt_affine 4 entries, opacity one row per Gaussian")
        stream = torch.cuda.current_stream(dev).cuda_stream
        geom_bytes = lib.eogs_geom_bytes(P)
        geom = torch.empty(geom_bytes + 256, dtype=torch.uint8, device=dev)
        info_dev = geom.data_ptr() + geom_bytes
        info_host, info_np = _info_host(dev)
        _cabi.check(lib.eogs_forward_geometry_params_band(
            stream, P, W, H, rb, re, _ptr(xyz), _ptr(log_scales), _ptr(raw_rotations), _ptr(opacity_logits),
            _ptr(features_dc), _ptr(alt_affine), _ptr(viewmatrix), float(scale_modifier), int(bool(antialiasing)),
            radii.data_ptr(), geom.data_ptr(), info_dev, info_host.data_ptr()), "eogs_forward_geometry_params_band")
        num_rendered, err = _wait_info(info_np, dev)
        if err & ERR_TOO_MANY_INSTANCES:
            raise _cabi.EogsRasterError("more than 2^32 (Gaussian, tile) instances: render the view in tile bands "
                                        "(eogs2_b200.bands)")
        if err & ERR_ALTITUDE_ABOVE_200:
            raise RuntimeError("Point is too high: a Gaussian's altitude exceeds 200 (depth = 200 - altitude < 0)")
        _debug_sync(debug, "preprocess")
        image = torch.empty(lib.eogs_image_bytes_band(W, H, rb, re), dtype=torch.uint8, device=dev)
        point_list = binning = None
        if num_rendered > 0:
            point_list = torch.empty(lib.eogs_point_list_words(num_rendered), dtype=torch.int32, device=dev)
            binning = torch.empty(lib.eogs_binning_bytes(W, H, num_rendered), dtype=torch.uint8, device=dev)
        _cabi.check(lib.eogs_forward_render_band(
            stream, P, W, H, 5, rb, re, num_rendered, geom.data_ptr(), _ptr(point_list), _ptr(binning),
            image.data_ptr(), _ptr(bg), color.data_ptr(), invdepth.data_ptr()), "eogs_forward_render_band")
        _debug_sync(debug, "render")
    return ForwardState(P, W, H, 5, num_rendered, geom, point_list, image, radii, color, invdepth, band)


def backward_params_raw(state: ForwardState, bg, xyz, opacity_logits, log_scales, raw_rotations, alt_affine,
                        scale_modifier, viewmatrix, projmatrix, dL_dcolor, dL_dinvdepth, antialiasing=False,
                        debug=False):
    """Returns (dL_dmeans2D, dL_dfeatures_dc [P,3], dL_dopacity_logits [P,1], dL_dxyz, dL_dlog_scales,
    dL_draw_rotations, cam_sums[16], alt_sums[4])."""
    lib = _cabi.load()
    dev = xyz.device
    P, W, H = state.P, state.W, state.H
    rb, re = state.rows
    with torch.cuda.device(dev):
        opts = dict(dtype=torch.float32, device=dev)
        g2d = torch.empty((P, 3), **opts)
        gfdc = torch.empty((P, 3), **opts)
        gop = torch.empty((P, 1), **opts)
        gxyz = torch.empty((P, 3), **opts)
        gsc = torch.empty((P, 3), **opts)
        grot = torch.empty((P, 4), **opts)
        sums = torch.zeros(20, **opts)                      # cam_sums[16] | alt_sums[4]
        if P == 0:
            return g2d, gfdc, gop, gxyz, gsc, grot, sums[:16], sums[16:]
        xyz = _f32c(xyz, "xyz", dev)
        opacity_logits = _f32c(opacity_logits, "opacity", dev)
        log_scales = _f32c(log_scales, "scaling", dev)
        raw_rotations = _quat16(_f32c(raw_rotations, "rotation", dev))
        alt_affine = _f32c(alt_affine, "alt_affine", dev)
        viewmatrix = _f32c(viewmatrix, "viewmatrix", dev)
        projmatrix = _f32c(projmatrix, "projmatrix", dev)
        bg = _f32c(bg, "bg", dev)
        dL_dcolor = _f32c(dL_dcolor, "grad_out_color", dev)
        if dL_dcolor.numel() != 5 * state.band_height * W:
            raise RuntimeError("grad_out_color does not match the rendered image")
        if dL_dinvdepth is not None:
            dL_dinvdepth = _f32c(dL_dinvdepth, "grad_out_depth", dev)
        grad_scratch = torch.empty(lib.eogs_grad_scratch_floats(P), **opts)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _cabi.check(lib.eogs_backward_params_band(
            stream, P, W, H, rb, re, state.num_rendered, _ptr(xyz), _ptr(log_scales), _ptr(raw_rotations),
            _ptr(opacity_logits), _ptr(alt_affine), _ptr(viewmatrix), _ptr(projmatrix), float(scale_modifier),
            int(bool(antialiasing)), _ptr(bg), state.radii.data_ptr(), state.geom.data_ptr(), _ptr(state.point_list),
            state.image.data_ptr(), _ptr(dL_dcolor), _ptr(dL_dinvdepth), grad_scratch.data_ptr(),
            g2d.data_ptr(), gfdc.data_ptr(), gop.data_ptr(), gxyz.data_ptr(), gsc.data_ptr(), grot.data_ptr(),
            sums.data_ptr(), sums[16:].data_ptr()), "eogs_backward_params_band")
        _debug_sync(debug, "backward")
    return g2d, gfdc, gop, gxyz, gsc, grot, sums[:16], sums[16:]


class _RasterizeFromParams(torch.autograd.Function):
    """(xyz, means2D, features_dc, opacity_logits, log_scales, raw_rotations, viewmatrix, alt_affine) ->
    (color[5,H,W], radii[P], invdepth[1,H,W]); the counterpart of activations + colors_precomp +
    _RasterizeGaussians (DGR/diff_gaussian_rasterization/__init__.py:53-216) in one autograd node."""

    @staticmethod
    def forward(ctx, xyz, means2D, features_dc, opacity_logits, log_scales, raw_rotations, viewmat, alt_affine,
                raster_settings):
        rs = raster_settings
        state = forward_params_raw(rs.bg, xyz, features_dc, opacity_logits, log_scales, raw_rotations, alt_affine,
                                   rs.scale_modifier, viewmat, rs.image_height, rs.image_width, rs.antialiasing, rs.debug)
        ctx.raster_settings = rs
        ctx.fdc_shape = features_dc.shape
        ctx.logit_shape = opacity_logits.shape
        ctx.state = ForwardState(state.P, state.W, state.H, 5, state.num_rendered, state.geom, state.point_list,
                                 state.image, state.radii, None, None, state.band)
        ctx.save_for_backward(xyz, opacity_logits, log_scales, raw_rotations, alt_affine)
        ctx.mark_non_differentiable(state.radii)
        ctx.set_materialize_grads(False)
        return state.color, state.radii, state.invdepth

    @staticmethod
    def backward(ctx, grad_color, _, grad_invdepth):
        rs = ctx.raster_settings
        xyz, opacity_logits, log_scales, raw_rotations, alt_affine = ctx.saved_tensors
        st = ctx.state
        if grad_color is None:
            grad_color = torch.zeros((5, st.H, st.W), dtype=torch.float32, device=xyz.device)
        g2d, gfdc, gop, gxyz, gsc, grot, cam_sums, alt_sums = backward_params_raw(
            st, rs.bg, xyz, opacity_logits, log_scales, raw_rotations, alt_affine, rs.scale_modifier,
            rs.viewmatrix, rs.projmatrix, grad_color, grad_invdepth, rs.antialiasing, rs.debug)
        grad_view = None
        if ctx.needs_input_grad[6]:
            with torch.no_grad():
                grad_view = assemble_grad_viewmatrix(cam_sums, rs.viewmatrix, st.W, st.H)
        grad_alt = alt_sums.to(alt_affine.dtype) if ctx.needs_input_grad[7] else None
        return (gxyz, g2d, gfdc.view(ctx.fdc_shape), gop.view(ctx.logit_shape), gsc, grot, grad_view, grad_alt, None)


def rasterize_params(xyz, means2D, features_dc, opacity_logits, log_scales, raw_rotations, alt_affine,
                     raster_settings: GaussianRasterizationSettings):
    """Tensor-level fused call.  `raster_settings` is the reference's NamedTuple; `alt_affine[4]` = (a, b) of
    altitude = a . xyz + b."""
    return _RasterizeFromParams.apply(xyz, means2D, features_dc, opacity_logits, log_scales, raw_rotations,
                                      raster_settings.viewmatrix, alt_affine, raster_settings)


def altitude_row(affine_t: torch.Tensor) -> torch.Tensor:
    """(a, b) of the altitude component from a camera's TRANSPOSED 4x4 affine (affine_cameras.py:432-438:
    uva = xyz @ affine[:3,:3] + affine[3,:3])."""
    return torch.cat([affine_t[:3, 2], affine_t[3:4, 2]])


def render_fused(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier: float = 1.0,
                 override_color: Optional[torch.Tensor] = None, use_trained_exp: bool = False) -> dict:
    """Drop-in for gaussian_renderer.renderer.render (renderer.py:27-144): same arguments, same dict."""
    xyz = pc._xyz
    screenspace_points = torch.zeros_like(xyz, requires_grad=True)     # grad slot of the 2D means (renderer.py:31-40)
    viewmatrix = viewpoint_camera.world_view_transform
    projmatrix = viewpoint_camera.full_proj_transform
    if getattr(viewpoint_camera, "learn_wv_only_lastparam", False):    # renderer.py:59-65
        viewmatrix = viewmatrix.clone()
        projmatrix = projmatrix.clone()
        viewmatrix[-1, :] = viewmatrix[-1, :] + viewpoint_camera.last_row
        projmatrix[-1, :] = projmatrix[-1, :] + viewpoint_camera.last_row
    settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(getattr(viewpoint_camera, "FoVx", 0.0) * 0.5),
        tanfovy=math.tan(getattr(viewpoint_camera, "FoVy", 0.0) * 0.5),
        bg=bg_color, scale_modifier=scaling_modifier, viewmatrix=viewmatrix, projmatrix=projmatrix,
        sh_degree=getattr(pc, "active_sh_degree", 0), campos=getattr(viewpoint_camera, "camera_center", None),
        prefiltered=False, debug=bool(getattr(pipe, "debug", False)),
        antialiasing=bool(getattr(pipe, "antialiasing", False)))

    fused_ok = override_color is None and not getattr(pipe, "compute_cov3D_python", False) and not use_trained_exp
    if fused_ok:
        affine_t = getattr(viewpoint_camera, "affine", viewpoint_camera.world_view_transform)
        rendered_image, radii, _ = rasterize_params(xyz, screenspace_points, pc._features_dc, pc._opacity,
                                                    pc._scaling, pc._rotation, altitude_row(affine_t), settings)
    else:
        # the reference's own sequence (renderer.py:84-122), through the unfused rasterizer
        if override_color is None:
            rgb = (pc._features_dc * SH_C0 + 0.5).squeeze(1)
            altitude = viewpoint_camera.ECEF_to_UVA(xyz)[..., 2].unsqueeze(-1)
            colors = torch.cat([rgb, altitude, torch.ones_like(altitude)], dim=-1)
        else:
            colors = override_color
        kw = dict(cov3D_precomp=pc.get_covariance(scaling_modifier)) if getattr(pipe, "compute_cov3D_python", False) \
            else dict(scales=pc.get_scaling, rotations=pc.get_rotation)
        rendered_image, radii, _ = GaussianRasterizer(settings)(
            means3D=xyz, means2D=screenspace_points, opacities=pc.get_opacity, colors_precomp=colors, **kw)
        if use_trained_exp:                                            # renderer.py:124-132
            exposure = pc.get_exposure_from_name(viewpoint_camera.image_name)
            rendered_image = torch.matmul(rendered_image.permute(1, 2, 0), exposure[:3, :3]).permute(2, 0, 1) \
                + exposure[:3, 3, None, None]
    out = {"render": rendered_image, "viewspace_points": screenspace_points}
    if getattr(pipe, "require_radii", False):
        out["visibility_filter"] = (radii > 0).nonzero()
        out["radii"] = radii
    return out
