"""Tile-band sharding of ONE view across GPUs (BASELINE configs[4]: 5 M Gaussians,
8192x8192 altitude/DSM render on 8 B200).

The reference renders a view on one GPU (utils/general_utils.py:155 pins cuda:0; no multi-device
code).  After binning, tiles are independent (renderCUDA is one block per tile,
DGR/cuda_rasterizer/forward.cu:288-411), so a huge view shards by horizontal bands of 16-pixel
tile rows: Gaussians are replicated, every rank projects all of them (the per-Gaussian stage is
~1 % of a render), emits / sorts / blends only the instances of its own tile rows through the
`*_band` entry points of the C ABI (include/eogs_raster.h), and the image bands are exchanged with
ONE all-gather.  For training, each rank backpropagates its band and the per-Gaussian gradients
(plus the 14 camera sums) are summed with ONE all-reduce — the same bucket as the data-parallel
path (eogs2_b200/dp.py).  A band's sorted list and tile ranges are the whole-image ones restricted
to its tiles (tests/test_bands_gpu.py checks that bit for bit), so the gathered image is
bit-identical to a single-GPU render.

One process per GPU (torchrun); NCCL over NVLink on a B200 box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

TILE = 16


def split_rows(grid_y: int, world: int, weights: Optional[Sequence[float]] = None) -> List[Tuple[int, int]]:
    """Partition tile rows [0, grid_y) into `world` contiguous bands [row_begin, row_end).

    Without weights the rows are split evenly (the first grid_y % world bands get one more row).
    With per-row weights (e.g. instances per tile row of the previous frame, read from its tile
    ranges) band boundaries follow the weight prefix sum, so every rank gets about the same blend
    work; every band keeps at least one row.  Ranks beyond grid_y get an empty band (b, b)."""
    if grid_y <= 0 or world <= 0:
        raise ValueError("grid_y and world must be positive")
    n = min(world, grid_y)
    if weights is None:
        base, extra = divmod(grid_y, n)
        cuts = [0]
        for r in range(n):
            cuts.append(cuts[-1] + base + (1 if r < extra else 0))
    else:
        if len(weights) != grid_y:
            raise ValueError("weights must have one entry per tile row")
        w = [max(float(x), 0.0) for x in weights]
        prefix = [0.0]
        for x in w:
            prefix.append(prefix[-1] + x)
        total = prefix[-1]
        cuts = [0]
        for r in range(1, n):
            y = cuts[-1] + 1
            while y < grid_y and prefix[y] < total * r / n:
                y += 1
            cuts.append(min(max(y, cuts[-1] + 1), grid_y - (n - r)))      # >= 1 row for every band
        cuts.append(grid_y)
    bands = [(cuts[i], cuts[i + 1]) for i in range(n)]
    bands += [(grid_y, grid_y)] * (world - n)
    return bands


def band_height(band: Tuple[int, int], H: int) -> int:
    rb, re = band
    return max(0, min(H, TILE * re) - TILE * rb)


def gather_bands(local: torch.Tensor, bands: Sequence[Tuple[int, int]], H: int,
                 group=None) -> torch.Tensor:
    """All-gather image bands [C, band_h, W] (band_h may differ per rank) into the full [C, H, W]
    image on every rank: one collective on a buffer padded to the tallest band."""
    world = len(bands)
    C, _, W = local.shape
    heights = [band_height(b, H) for b in bands]
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        assert heights[0] == local.shape[1]
        return local
    hmax = max(heights)
    padded = local
    if local.shape[1] != hmax:
        padded = local.new_zeros((C, hmax, W))
        padded[:, :local.shape[1]] = local
    out = local.new_empty((world, C, hmax, W))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group) if local.is_cuda else \
        dist.all_gather(list(out.unbind(0)), padded.contiguous(), group=group)
    full = local.new_empty((C, H, W))
    for r, (rb, _) in enumerate(bands):
        if heights[r]:
            full[:, TILE * rb:TILE * rb + heights[r]] = out[r, :, :heights[r]]
    return full


def render_sharded(render_band: Callable[[Tuple[int, int]], Tuple[torch.Tensor, torch.Tensor]],
                   H: int, rank: int, world: int, weights: Optional[Sequence[float]] = None,
                   gather: bool = True, group=None):
    """Render this rank's band with `render_band((row_begin, row_end)) -> (color[C,h,W],
    invdepth[1,h,W])` and (optionally) all-gather the full images.  Returns
    (color, invdepth, band): full images when gather=True, else this rank's bands."""
    grid_y = (H + TILE - 1) // TILE
    if grid_y < world:
        # raised on EVERY rank (a rank that bailed out alone would leave the others in the all-gather)
        raise RuntimeError(f"{grid_y} tile rows cannot be sharded over {world} ranks: use at most {grid_y}")
    bands = split_rows(grid_y, world, weights)
    band = bands[rank]
    color, invdepth = render_band(band)
    if gather and world > 1:
        both = gather_bands(torch.cat([color, invdepth], 0), bands, H, group)
        color, invdepth = both[:-1], both[-1:]
    return color, invdepth, band


def forward_band(bg, means3D, colors, opacities, scales, rotations, scale_modifier, cov3D_precomp,
                 viewmatrix, H: int, W: int, rank: int, world: int, antialiasing: bool = False,
                 weights: Optional[Sequence[float]] = None, gather: bool = True, group=None):
    """Band-sharded forward of one view through the sm_100a kernels.  Returns
    (color, invdepth, state): `state` is this rank's ForwardState (for backward_band)."""
    from .rasterizer import rasterize_forward_raw
    holder = {}

    def render_band(band):
        st = rasterize_forward_raw(bg, means3D, colors, opacities, scales, rotations, scale_modifier,
                                   cov3D_precomp, viewmatrix, H, W, antialiasing, False, band=band)
        holder["st"] = st
        return st.color, st.invdepth

    color, invdepth, _ = render_sharded(render_band, H, rank, world, weights, gather, group)
    return color, invdepth, holder["st"]


def backward_band(state, bg, means3D, colors, opacities, scales, rotations, scale_modifier, cov3D_precomp,
                  viewmatrix, projmatrix, dL_dcolor_full: torch.Tensor, dL_dinvdepth_full: Optional[torch.Tensor],
                  antialiasing: bool = False, reduce: bool = True, group=None):
    """Backward of this rank's band given the upstream gradient of the FULL image, then one
    all-reduce (SUM) of a flat bucket holding every per-Gaussian gradient and the 16 camera sums.
    Returns the tuple of rasterize_backward_raw, identical on every rank when reduce=True."""
    from .rasterizer import rasterize_backward_raw
    rb, _ = state.rows
    y0, h = TILE * rb, state.band_height
    dcol = dL_dcolor_full[:, y0:y0 + h].contiguous()
    dinv = None if dL_dinvdepth_full is None else dL_dinvdepth_full.reshape(-1, dL_dcolor_full.shape[-1])[y0:y0 + h].contiguous()
    grads = rasterize_backward_raw(state, bg, means3D, colors, opacities, scales, rotations, scale_modifier,
                                   cov3D_precomp, viewmatrix, projmatrix, dcol, dinv, antialiasing)
    if reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        live = [g for g in grads if g is not None]
        flat = torch.cat([g.reshape(-1) for g in live])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for g in live:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
    return grads


def row_weights_from_state(state) -> List[float]:
    """Instances per tile row of a rendered whole-image state (from its tile ranges): the weights
    that balance the NEXT frame's bands."""
    from .rasterizer import export_state
    if state.band is not None:
        raise ValueError("row weights need a whole-image state")
    grid_x = (state.W + TILE - 1) // TILE
    r = export_state(state)["ranges"].to(torch.int64)
    per_tile = (r[:, 1] - r[:, 0]).view(-1, grid_x)
    return per_tile.sum(1).cpu().tolist()


def forward_strips(bg, means3D, colors, opacities, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                   H: int, W: int, antialiasing: bool = False, max_tiles: int = 65536):
    """Forward-only render of ONE huge view on ONE GPU as consecutive bands of at most `max_tiles` tiles (offline
    products: the full-resolution nadir DSM render, train_pan.py:738-787).  Above 65 536 tiles the tile sort needs
    32-bit keys and a third radix pass; a band below that limit sorts 16-bit keys in two passes, and a band's list
    is the whole-image list restricted to its tiles (tests/test_bands_gpu.py), so the stitched image is bit-identical
    to the single call — 5 M Gaussians at 8192²: 40.5 ms whole image, 30.0 ms in 8 bands (profiles/r03z).
    Returns (color [C,H,W], invdepth [1,H,W], radii [P])."""
    from .rasterizer import rasterize_forward_raw
    grid_x, grid_y = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    rows = max(1, min(grid_y, max_tiles // max(grid_x, 1)))
    color = invdepth = radii = None
    for rb in range(0, grid_y, rows):
        st = rasterize_forward_raw(bg, means3D, colors, opacities, scales, rotations, scale_modifier, cov3D_precomp,
                                   viewmatrix, H, W, antialiasing, False, band=(rb, min(grid_y, rb + rows)))
        if color is None:
            color = st.color.new_empty((st.color.shape[0], H, W))
            invdepth = st.invdepth.new_empty((1, H, W))
            radii = st.radii
        y0 = TILE * rb
        color[:, y0:y0 + st.band_height] = st.color
        invdepth[:, y0:y0 + st.band_height] = st.invdepth.reshape(1, st.band_height, W)
    return color, invdepth, radii
