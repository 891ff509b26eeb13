"""Sun-view shadow pass (north-star subsystem (d)): the virtual-camera render and
its reprojection into the actual camera's frame.

Mirror of `render_resample_virtual_camera` (gaussian_renderer/renderer_cc_shadow.py:6-54): same
arguments, same returns.  Step 1 (the 2W x 2H render from the sun camera,
scene/cameras/affine_cameras.py:350-370) goes through the fused render glue
(`eogs2_b200.fused.render_fused`) or any `render`-shaped callable; steps 2-3 — the per-pixel
`cam2virt @ uva` einsum, the 5-channel bilinear `grid_sample`, the channel split and the `-100`
out-of-footprint overwrite, plus their autograd — are ONE forward and ONE backward sm_100a kernel
(`csrc/resample.cu`, `eogs_resample_forward/backward`).  No CPU path.
"""
from __future__ import annotations

import torch

from . import _cabi
from .rasterizer import _f32c, _ptr


class _ResampleVirtual(torch.autograd.Function):
    """(virtual_render [Cv,Hv,Wv], cam2virt [3,3], rendered_uva [H,W,3]) ->
    (rgb [3,H,W], altitude [H,W], virtual_uv [H,W,2])."""

    @staticmethod
    def forward(ctx, virtual_render, cam2virt, rendered_uva):
        lib = _cabi.load()
        if not virtual_render.is_cuda:
            raise _cabi.EogsRasterError("virtual_render must be a CUDA tensor: the resample has no CPU path")
        dev = virtual_render.device
        Cv, Hv, Wv = (int(x) for x in virtual_render.shape)
        H, W = int(rendered_uva.shape[0]), int(rendered_uva.shape[1])
        virt = _f32c(virtual_render, "virtual_render", dev)
        M = _f32c(cam2virt, "cam2virt", dev)
        uva = _f32c(rendered_uva, "rendered_uva", dev)
        if M.numel() != 9 or uva.shape[-1] != 3:
            raise RuntimeError("cam2virt must be 3x3 and rendered_uva [H,W,3]")
        with torch.cuda.device(dev):
            rgb = torch.empty((3, H, W), dtype=torch.float32, device=dev)
            alt = torch.empty((H, W), dtype=torch.float32, device=dev)
            uv = torch.empty((H, W, 2), dtype=torch.float32, device=dev)
            _cabi.check(lib.eogs_resample_forward(
                torch.cuda.current_stream(dev).cuda_stream, Cv, Hv, Wv, H, W, _ptr(virt), _ptr(M), _ptr(uva),
                rgb.data_ptr(), alt.data_ptr(), uv.data_ptr()), "eogs_resample_forward")
        ctx.save_for_backward(virt, M, uva)
        ctx.set_materialize_grads(False)
        return rgb, alt, uv

    @staticmethod
    def backward(ctx, d_rgb, d_alt, d_uv):
        lib = _cabi.load()
        virt, M, uva = ctx.saved_tensors
        dev = virt.device
        Cv, Hv, Wv = (int(x) for x in virt.shape)
        H, W = int(uva.shape[0]), int(uva.shape[1])
        with torch.cuda.device(dev):
            g = [None if t is None else _f32c(t, "grad", dev) for t in (d_rgb, d_alt, d_uv)]
            d_virt = torch.empty_like(virt)
            d_uva = torch.empty_like(uva)
            d_M = torch.empty(9, dtype=torch.float32, device=dev)
            _cabi.check(lib.eogs_resample_backward(
                torch.cuda.current_stream(dev).cuda_stream, Cv, Hv, Wv, H, W, _ptr(virt), _ptr(M), _ptr(uva),
                _ptr(g[0]), _ptr(g[1]), _ptr(g[2]), d_virt.data_ptr(), d_uva.data_ptr(), d_M.data_ptr()),
                "eogs_resample_backward")
        return d_virt, d_M.view(3, 3), d_uva


def resample_virtual(virtual_render: torch.Tensor, cam2virt: torch.Tensor, rendered_uva: torch.Tensor):
    """Steps 2-3 of render_resample_virtual_camera as one differentiable op."""
    return _ResampleVirtual.apply(virtual_render, cam2virt, rendered_uva)


def render_resample_virtual_camera(virtual_camera, cam2virt, rendered_uva, gaussians, pipe, background,
                                   return_extra: bool = False, render_fn=None):
    """renderer_cc_shadow.py:6-54.  render_fn defaults to the fused glue (eogs2_b200.fused.render_fused);
    pass the reference's `render` to keep its glue."""
    if render_fn is None:
        from .fused import render_fused as render_fn
    virtual_render = render_fn(virtual_camera, gaussians, pipe, background)["render"]
    rgb, altitude, uv = resample_virtual(virtual_render, cam2virt, rendered_uva)
    if return_extra:
        return rgb, altitude, uv, virtual_render
    return rgb, altitude, uv
