"""Host-side mirror of the reference rasterizer's Python API, on top of libeogs_raster.so.

Same names, argument order, return values and error behaviour as
DGR/diff_gaussian_rasterization/__init__.py (DGR = src/gaussiansplatting/submodules/
diff-gaussian-rasterization of gardiens/EOGS2):

  GaussianRasterizationSettings   __init__.py:219-232   (13 fields, same order)
  GaussianRasterizer              __init__.py:235-300   (forward, markVisible)
  rasterize_gaussians             __init__.py:26-50
  _RasterizeGaussians             __init__.py:53-216    (autograd.Function; backward returns
                                                         the same 10-tuple, incl. grad_viewmatrix)

PyTorch is plumbing here (device memory, streams, autograd graph); all arithmetic runs in
the hand-written sm_100a kernels behind the C ABI.  There is no CPU path: tensors must live
on a CUDA device and the shared library must be present.
"""
from __future__ import annotations

import ctypes as C
import threading
import time
from dataclasses import dataclass
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _cabi

NUM_CHANNELS = 5          # DGR/cuda_rasterizer/config.h:14
ERR_ALTITUDE_ABOVE_200 = 1
ERR_TOO_MANY_INSTANCES = 2


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    antialiasing: bool


# ---------------------------------------------------------------------------------------------
# raw (non-autograd) layer: one function per C-ABI call group
# ---------------------------------------------------------------------------------------------
@dataclass
class ForwardState:
    """Everything the backward (and the parity tests) need from one forward call."""
    P: int
    W: int
    H: int
    channels: int
    num_rendered: int
    geom: Optional[torch.Tensor]
    point_list: Optional[torch.Tensor]
    image: Optional[torch.Tensor]
    radii: torch.Tensor
    color: Optional[torch.Tensor]
    invdepth: Optional[torch.Tensor]
    band: Optional[tuple] = None      # (row_begin, row_end) tile rows this state covers; None = whole image

    @property
    def rows(self) -> tuple:
        return self.band if self.band is not None else (0, (self.H + 15) // 16)

    @property
    def band_height(self) -> int:
        rb, re = self.rows
        return min(self.H, 16 * re) - 16 * rb


_pinned_info: dict = {}


def _info_key(device: torch.device):
    # one pinned slot per (device, stream, host thread): two Python threads rendering on the same stream must not
    # share the words a device -> host copy is about to fill
    return (device.index, torch.cuda.current_stream(device).cuda_stream, threading.get_ident())


def _info_host(device: torch.device) -> torch.Tensor:
    key = _info_key(device)
    t = _pinned_info.get(key)
    if t is None:
        buf = torch.zeros(4, dtype=torch.int32).pin_memory()        # eogs_forward_info: I, error, ready, reserved
        t = [buf, buf.numpy(), False]      # the numpy view reads the pinned words without a torch dispatch
        _pinned_info[key] = t
    if t[2]:
        # the previous call on this stream never collected its words (it raised in between): its copy may still be
        # in flight and would raise `ready` for THIS call — let it land first
        torch.cuda.current_stream(device).synchronize()
    t[2] = True
    t[1][2] = 0                            # `ready` is raised by the device -> host copy of this call
    return t[0], t[1]


_POLL_SECONDS = 2e-3


def _wait_info(info_np, device: torch.device):
    """Wait for the geometry stage's (I, error) words.  Their copies into pinned memory are enqueued right after the
    projection kernel, before the depth sort — payload first, then the `ready` word (cabi.cu) — so polling `ready`
    returns while the sort still runs and the render stage can be enqueued behind it without a GPU bubble (the
    reference blocks on a cudaMemcpy after its scan, rasterizer_impl.cu:284).  The payload is read only AFTER `ready`
    was seen.  Time-bounded: after 2 ms of polling it falls back to a stream synchronisation, which also surfaces
    CUDA errors."""
    deadline = time.perf_counter() + _POLL_SECONDS
    while info_np[2] == 0:
        if time.perf_counter() > deadline:
            torch.cuda.current_stream(device).synchronize()
            if info_np[2] == 0:
                raise _cabi.EogsRasterError("geometry stage finished without publishing its instance count")
            break
    entry = _pinned_info.get(_info_key(device))
    if entry is not None:
        entry[2] = False                   # collected
    return int(info_np[0]) & 0xFFFFFFFF, int(info_np[1])


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t: torch.Tensor, name: str, device: torch.device) -> torch.Tensor:
    if t.numel() == 0:
        return t
    if t.device != device:
        raise _cabi.EogsRasterError(f"{name} is on {t.device}, expected {device} (no CPU path)")
    if t.dtype != torch.float32:
        raise _cabi.EogsRasterError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _quat16(t: torch.Tensor) -> torch.Tensor:
    """Quaternions are read as float4: a view whose storage offset leaves it only 4-byte aligned (legal input for the
    reference, which indexes floats) is copied once into an aligned buffer."""
    return t if t.numel() == 0 or t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


def _debug_sync(debug: bool, what: str) -> None:
    # CHECK_CUDA(..., debug): synchronise and raise after each stage (DGR auxiliary.h:178-185)
    if debug:
        try:
            torch.cuda.synchronize()
        except RuntimeError as e:
            raise RuntimeError(f"[CUDA ERROR] after {what}: {e}") from e


def rasterize_forward_raw(bg, means3D, colors, opacities, scales, rotations, scale_modifier,
                          cov3D_precomp, viewmatrix, image_height, image_width,
                          antialiasing=False, debug=False, band=None) -> ForwardState:
    """Counterpart of _C.rasterize_gaussians (DGR/rasterize_points.cu:35-124).

    band = (row_begin, row_end): render only those tile rows (multi-GPU tile sharding of one view,
    eogs2_b200/bands.py); color / invdepth then hold pixel rows [16*row_begin, min(H, 16*row_end))."""
    lib = _cabi.load()
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")       # rasterize_points.cu:58-60
    if not means3D.is_cuda:
        raise _cabi.EogsRasterError("means3D must be a CUDA tensor: this rasterizer has no CPU path")
    dev = means3D.device
    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)
    rb, re = (0, (H + 15) // 16) if band is None else (int(band[0]), int(band[1]))
    if not (0 <= rb < re <= (H + 15) // 16):
        raise _cabi.EogsRasterError(f"bad band: tile rows [{rb}, {re}) of {(H + 15) // 16}")
    Hb = min(H, 16 * re) - 16 * rb
    band = None if band is None else (rb, re)

    if colors is None or colors.numel() == 0:
        if P != 0:
            raise RuntimeError("For non-RGB, provide precomputed Gaussian colors!")   # rasterizer_impl.cu:244-247
        channels = NUM_CHANNELS
    else:
        channels = int(colors.size(1))

    with torch.cuda.device(dev):
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        if P == 0:
            # rasterize_points.cu:88 — nothing is launched; the image stays zero (not bg)
            color = torch.zeros((channels, Hb, W), dtype=torch.float32, device=dev)
            invdepth = torch.zeros((1, Hb, W), dtype=torch.float32, device=dev)
            return ForwardState(0, W, H, channels, 0, None, None, None, radii, color, invdepth, band)

        means3D = _f32c(means3D, "means3D", dev)
        colors = _f32c(colors, "colors_precomp", dev)
        opacities = _f32c(opacities, "opacities", dev)
        scales = _f32c(scales, "scales", dev)
        rotations = _quat16(_f32c(rotations, "rotations", dev))
        cov3D_precomp = _f32c(cov3D_precomp, "cov3D_precomp", dev)
        viewmatrix = _f32c(viewmatrix, "viewmatrix", dev)
        bg = _f32c(bg, "bg", dev)
        if bg.numel() != channels:
            raise RuntimeError(f"bg has {bg.numel()} channels, colors_precomp has {channels}")
        if opacities.numel() != P or colors.size(0) != P:
            raise RuntimeError("opacities / colors_precomp must have one row per Gaussian")

        stream = torch.cuda.current_stream(dev).cuda_stream
        geom_bytes = lib.eogs_geom_bytes(P)
        geom = torch.empty(geom_bytes + 256, dtype=torch.uint8, device=dev)
        info_dev = geom.data_ptr() + geom_bytes
        info_host, info_np = _info_host(dev)

        _cabi.check(lib.eogs_forward_geometry_band(
            stream, P, W, H, channels, rb, re, _ptr(means3D), _ptr(scales), _ptr(rotations), _ptr(cov3D_precomp),
            _ptr(opacities), _ptr(colors), _ptr(viewmatrix), float(scale_modifier), int(bool(antialiasing)),
            radii.data_ptr(), geom.data_ptr(), info_dev, info_host.data_ptr()), "eogs_forward_geometry_band")
        # everything the render stage needs is allocated while the projection kernel already runs
        color = torch.empty((channels, Hb, W), dtype=torch.float32, device=dev)
        invdepth = torch.empty((1, Hb, W), dtype=torch.float32, device=dev)
        image = torch.empty(lib.eogs_image_bytes_band(W, H, rb, re), dtype=torch.uint8, device=dev)
        # The instance count sizes the binning buffers (reference: blocking cudaMemcpy,
        # rasterizer_impl.cu:284).
        num_rendered, err = _wait_info(info_np, dev)
        if err & ERR_TOO_MANY_INSTANCES:
            raise _cabi.EogsRasterError("more than 2^32 (Gaussian, tile) instances: render the view in tile bands "
                                        "(eogs2_b200.bands)")
        if err & ERR_ALTITUDE_ABOVE_200:
            # reference: device printf("Point is too high") + __trap() (forward.cu:267-272)
            raise RuntimeError("Point is too high: a Gaussian's altitude exceeds 200 (depth = 200 - altitude < 0)")
        _debug_sync(debug, "preprocess")

        point_list = None
        binning = None
        if num_rendered > 0:
            point_list = torch.empty(lib.eogs_point_list_words(num_rendered), dtype=torch.int32, device=dev)
            binning = torch.empty(lib.eogs_binning_bytes(W, H, num_rendered), dtype=torch.uint8, device=dev)
        _cabi.check(lib.eogs_forward_render_band(
            stream, P, W, H, channels, rb, re, num_rendered, geom.data_ptr(), _ptr(point_list), _ptr(binning),
            image.data_ptr(), _ptr(bg), color.data_ptr(), invdepth.data_ptr()), "eogs_forward_render_band")
        _debug_sync(debug, "render")
        del binning   # scratch; the caching allocator keeps it stream-ordered
    return ForwardState(P, W, H, channels, num_rendered, geom, point_list, image, radii, color, invdepth, band)


def rasterize_backward_raw(state: ForwardState, bg, means3D, colors, opacities, scales, rotations,
                           scale_modifier, cov3D_precomp, viewmatrix, projmatrix, dL_dcolor,
                           dL_dinvdepth, antialiasing=False, debug=False, out=None):
    """Counterpart of _C.rasterize_gaussians_backward (DGR/rasterize_points.cu:126-224) plus the
    reductions of __init__.py:174-202.  Returns (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D,
    dL_dcov3D | None, dL_dscales | None, dL_drotations | None, cam_sums[16]).

    out: optional dict of preallocated contiguous fp32 CUDA tensors keyed "means2D", "colors",
    "opacity", "means3D", "cov3D", "scales", "rotations", "cam_sums" — the kernels then write the
    gradients straight into them (e.g. views of the data-parallel all-reduce bucket, dp.py)."""
    lib = _cabi.load()
    dev = means3D.device
    P, W, H, ch = state.P, state.W, state.H, state.channels
    rb, re = state.rows
    with torch.cuda.device(dev):
        opts = dict(dtype=torch.float32, device=dev)
        out = out or {}

        def buf(key, shape):
            t = out.get(key)
            if t is None:
                return torch.empty(shape, **opts)
            if (t.device != dev or t.dtype != torch.float32 or not t.is_contiguous()
                    or t.numel() != int(torch.Size(shape).numel())):
                raise _cabi.EogsRasterError(f"out[{key!r}] must be a contiguous float32 tensor of shape {tuple(shape)} on {dev}")
            if key == "rotations" and t.numel() and t.data_ptr() % 16 != 0:
                raise _cabi.EogsRasterError("out['rotations'] must be 16-byte aligned (quaternion gradients are stored as float4): "
                                            "start the segment on a multiple of 4 floats (dp.segment_layout)")
            return t.view(shape)

        dL_dmeans2D = buf("means2D", (P, 3))
        dL_dcolors = buf("colors", (P, ch))
        dL_dopacity = buf("opacity", (P, 1))
        dL_dmeans3D = buf("means3D", (P, 3))
        cam_sums = buf("cam_sums", (16,))
        has_cov = cov3D_precomp is not None and cov3D_precomp.numel() != 0
        dL_dcov3D = buf("cov3D", (P, 6)) if has_cov else None
        dL_dscales = None if has_cov else buf("scales", (P, 3))
        dL_drotations = None if has_cov else buf("rotations", (P, 4))
        if P == 0:
            cam_sums.zero_()
            return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dscales, dL_drotations, cam_sums

        means3D = _f32c(means3D, "means3D", dev)
        colors = _f32c(colors, "colors_precomp", dev)
        opacities = _f32c(opacities, "opacities", dev)
        scales = _f32c(scales, "scales", dev)
        rotations = _quat16(_f32c(rotations, "rotations", dev))
        cov3D_precomp = _f32c(cov3D_precomp, "cov3D_precomp", dev)
        viewmatrix = _f32c(viewmatrix, "viewmatrix", dev)
        projmatrix = _f32c(projmatrix, "projmatrix", dev)
        bg = _f32c(bg, "bg", dev)
        dL_dcolor = _f32c(dL_dcolor, "grad_out_color", dev)
        if dL_dcolor.numel() != ch * state.band_height * W:
            raise RuntimeError(f"grad_out_color has {dL_dcolor.numel()} elements, expected {ch}x{state.band_height}x{W}")
        if dL_dinvdepth is not None:
            dL_dinvdepth = _f32c(dL_dinvdepth, "grad_out_depth", dev)
            if dL_dinvdepth.numel() != state.band_height * W:
                raise RuntimeError("grad_out_depth does not match the rendered band")
        grad_scratch = torch.empty(lib.eogs_grad_scratch_floats(P), **opts)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _cabi.check(lib.eogs_backward_band(
            stream, P, W, H, ch, rb, re, state.num_rendered,
            _ptr(means3D), _ptr(scales), _ptr(rotations), _ptr(cov3D_precomp), _ptr(opacities), _ptr(colors),
            _ptr(viewmatrix), _ptr(projmatrix), float(scale_modifier), int(bool(antialiasing)), _ptr(bg),
            state.radii.data_ptr(), state.geom.data_ptr(), _ptr(state.point_list), state.image.data_ptr(),
            _ptr(dL_dcolor), _ptr(dL_dinvdepth), grad_scratch.data_ptr(),
            dL_dmeans2D.data_ptr(), dL_dcolors.data_ptr(), dL_dopacity.data_ptr(), dL_dmeans3D.data_ptr(),
            _ptr(dL_dcov3D), _ptr(dL_dscales), _ptr(dL_drotations), cam_sums.data_ptr()), "eogs_backward_band")
        _debug_sync(debug, "backward")
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dscales, dL_drotations, cam_sums


_assembly_cache: dict = {}


def _assembly_matrix(device: torch.device, W: int, H: int) -> torch.Tensor:
    """16x16 matrix M with grad_viewmatrix.flatten() = M @ cam_sums (built once per (device, W, H))."""
    key = (str(device), int(W), int(H))
    M = _assembly_cache.get(key)
    if M is None:
        M = torch.zeros(16, 16, dtype=torch.float32)
        scale = (W / 2.0, H / 2.0, 1.0)
        for r in range(3):
            for c in range(2):
                M[4 * r + c, 3 * c + r] += scale[r]          # diag(W/2, H/2, 1) @ (sum_b dL_dT[b]).view(2, 3)^T
                M[4 * r + c, 6 + 2 * r + c] += 1.0           # means3D^T @ grad_means2D
        for c in range(2):
            M[12 + c, 12 + c] += 1.0                         # grad_means2D.sum(0)
        M = M.to(device)
        _assembly_cache[key] = M
    return M


def assemble_grad_viewmatrix(cam_sums: torch.Tensor, like: torch.Tensor, W: int, H: int) -> torch.Tensor:
    """grad_viewmatrix from the 14 kernel-side sums, term by term as __init__.py:172-202:
         [:3, :2] += diag(W/2, H/2, 1) @ (sum_b dL_dT[b]).view(2, 3)^T
         [:3, :3] += means3D^T @ grad_means2D          (third column is zero: grad_means2D.z == 0)
         [-1, :3] += grad_means2D.sum(0)
    One matrix-vector product with a cached constant matrix: no host<->device traffic, so the
    backward never synchronises (the reference issues ~10 small torch kernels here)."""
    M = _assembly_matrix(cam_sums.device, W, H)
    return torch.mv(M, cam_sums.to(torch.float32)).view(4, 4).to(like.dtype)


def export_state(state: ForwardState) -> dict:
    """Internal state in the reference's layouts (geomState / binningState / imgState fields),
    for bit-level parity tests."""
    lib = _cabi.load()
    dev = state.radii.device
    P, W, H, I = state.P, state.W, state.H, state.num_rendered
    rb, re = state.rows
    tiles = ((W + 15) // 16) * (re - rb)
    npix = state.band_height * W
    with torch.cuda.device(dev):
        out = {
            "means2D": torch.zeros((P, 2), dtype=torch.float32, device=dev),
            "depths": torch.zeros((P,), dtype=torch.float32, device=dev),
            "conic_opacity": torch.zeros((P, 4), dtype=torch.float32, device=dev),
            "tiles_touched": torch.zeros((P,), dtype=torch.int32, device=dev),
            "keys_sorted": torch.zeros((I,), dtype=torch.int64, device=dev),
            "ranges": torch.zeros((tiles, 2), dtype=torch.int32, device=dev),
            "final_T": torch.zeros((npix,), dtype=torch.float32, device=dev),
            "n_contrib": torch.zeros((npix,), dtype=torch.int32, device=dev),
        }
        if P > 0:
            stream = torch.cuda.current_stream(dev).cuda_stream
            _cabi.check(lib.eogs_export_state_band(
                stream, P, W, H, rb, re, I, state.geom.data_ptr(), _ptr(state.point_list), state.image.data_ptr(),
                out["means2D"].data_ptr(), out["depths"].data_ptr(), out["conic_opacity"].data_ptr(),
                out["tiles_touched"].data_ptr(), _ptr(out["keys_sorted"]), out["ranges"].data_ptr(),
                out["final_T"].data_ptr(), out["n_contrib"].data_ptr()), "eogs_export_state_band")
        out["point_list"] = state.point_list[:I] if state.point_list is not None else \
            torch.zeros((0,), dtype=torch.int32, device=dev)        # (the I culling bytes behind the ids are internal)
        out["radii"] = state.radii
    return out


# ---------------------------------------------------------------------------------------------
# reference-shaped API
# ---------------------------------------------------------------------------------------------
def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(
        means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
        raster_settings.viewmatrix, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, viewmat, raster_settings):
        rs = raster_settings
        state = rasterize_forward_raw(
            rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
            cov3Ds_precomp, viewmat, rs.image_height, rs.image_width, rs.antialiasing, rs.debug)
        ctx.raster_settings = rs
        ctx.num_rendered = state.num_rendered
        # Keep only what the backward reads.  The output images must NOT be reachable from ctx:
        # color.grad_fn is this node, so ctx -> state -> color would be a reference cycle that only
        # Python's cyclic GC frees, and every step would then cudaMalloc fresh buffers.
        ctx.state = ForwardState(state.P, state.W, state.H, state.channels, state.num_rendered,
                                 state.geom, state.point_list, state.image, state.radii, None, None, state.band)
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, opacities)
        ctx.mark_non_differentiable(state.radii)
        # An unused output (EOGS never reads invdepths) then arrives as None in backward instead of a
        # materialised zero image (rasterize_points.cu:177-183 always gets one); same gradients.
        ctx.set_materialize_grads(False)
        return state.color, state.radii, state.invdepth

    @staticmethod
    def backward(ctx, grad_out_color, _, grad_out_depth):
        rs = ctx.raster_settings
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, opacities = ctx.saved_tensors
        state = ctx.state
        if grad_out_color is None:
            grad_out_color = torch.zeros((state.channels, state.H, state.W), dtype=torch.float32,
                                         device=means3D.device)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp,
         grad_scales, grad_rotations, cam_sums) = rasterize_backward_raw(
            state, rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
            cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, grad_out_color, grad_out_depth,
            rs.antialiasing, rs.debug)
        grad_viewmatrix = None
        if ctx.needs_input_grad[8]:
            with torch.no_grad():
                grad_viewmatrix = assemble_grad_viewmatrix(cam_sums, rs.viewmatrix, state.W, state.H)
        return (grad_means3D, grad_means2D, None, grad_colors_precomp, grad_opacities, grad_scales,
                grad_rotations, grad_cov3Ds_precomp, grad_viewmatrix, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # The reference's frustum test culls nothing for affine cameras (its body is commented
        # out, DGR/cuda_rasterizer/auxiliary.h:151-176): every Gaussian is visible.
        with torch.no_grad():
            lib = _cabi.load()
            if not positions.is_cuda:
                raise _cabi.EogsRasterError("positions must be a CUDA tensor")
            P = int(positions.size(0))
            present = torch.empty((P,), dtype=torch.bool, device=positions.device)
            rs = self.raster_settings
            with torch.cuda.device(positions.device):
                _cabi.check(lib.eogs_mark_visible(
                    torch.cuda.current_stream(positions.device).cuda_stream, P, _ptr(positions.contiguous()),
                    _ptr(rs.viewmatrix), _ptr(rs.projmatrix), _ptr(present)), "eogs_mark_visible")
        return present

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")

        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception(
                "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")

        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])

        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, raster_settings)
