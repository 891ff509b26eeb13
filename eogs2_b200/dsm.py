"""DSM splat on the GPU: the `plyflatten` step of `compute_dsm_from_view`
(utils/dsm_utils.py:7-51) without the device->host copy of the 3·H·W cloud.

`plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius, sigma)` has the signature and the
`[ysize, xsize, 1]` float32 result of the third-party function the reference imports (utils/dsm_utils.py:1,28-37);
`dsm_grid` is the extent arithmetic of utils/dsm_utils.py:20-25, and `compute_dsm` chains the two on a CUDA
cloud and returns the rasterio profile fields the reference fills (:41-50) except the CRS, which needs the
dataset's UTM zone and the `plyflatten.utils` helpers (on-disk formats stay with the reference).
No CPU path: the kernels live in libeogs_raster.so.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _cabi


def plyflatten(cloud: torch.Tensor, xoff: float, yoff: float, resolution: float, xsize: int, ysize: int,
               radius: int, sigma: float) -> torch.Tensor:
    if not isinstance(cloud, torch.Tensor) or not cloud.is_cuda:
        raise _cabi.EogsRasterError("plyflatten needs a CUDA tensor: there is no CPU path")
    if cloud.dim() != 2 or cloud.shape[1] != 3:
        raise _cabi.EogsRasterError("cloud must have dimensions (num_points, 3): x, y and one value column")
    lib = _cabi.load()
    pts = cloud.detach().to(torch.float64).contiguous()        # plyflatten converts its input to double
    with torch.cuda.device(pts.device):
        accum = torch.empty(2 * int(xsize) * int(ysize), dtype=torch.float64, device=pts.device)
        raster = torch.empty((int(ysize), int(xsize), 1), dtype=torch.float32, device=pts.device)
        stream = torch.cuda.current_stream(pts.device).cuda_stream
        _cabi.check(lib.eogs_dsm_splat(C.c_void_p(stream), C.c_longlong(pts.shape[0]), C.c_void_p(pts.data_ptr()),
                                       C.c_double(xoff), C.c_double(yoff), C.c_double(resolution), C.c_int(int(xsize)),
                                       C.c_int(int(ysize)), C.c_int(int(radius)), C.c_float(float(sigma)),
                                       C.c_void_p(accum.data_ptr()), C.c_void_p(raster.data_ptr())), "eogs_dsm_splat")
    return raster


def dsm_grid(xmin: float, xmax: float, ymin: float, ymax: float, resolution: float):
    """utils/dsm_utils.py:20-25."""
    xoff = math.floor(xmin / resolution) * resolution
    xsize = int(1 + math.floor((xmax - xoff) / resolution))
    yoff = math.ceil(ymax / resolution) * resolution
    ysize = int(1 - math.floor((ymin - yoff) / resolution))
    return xoff, yoff, xsize, ysize


def compute_dsm(cloud: torch.Tensor, resolution: float, radius: int = 1, sigma: float = float("inf")):
    """The numeric body of compute_dsm_from_view (utils/dsm_utils.py:18-50) for a CUDA cloud [N,3] in UTM metres:
    one 32-byte device->host read for the extent (the reference copies the whole cloud), then the splat."""
    pts = cloud.detach().to(torch.float64)
    lo, hi = torch.aminmax(pts[:, :2], dim=0)
    xmin, ymin, xmax, ymax = torch.cat([lo, hi]).tolist()
    xoff, yoff, xsize, ysize = dsm_grid(xmin, xmax, ymin, ymax, resolution)
    dsm = plyflatten(pts, xoff, yoff, resolution, xsize, ysize, radius=radius, sigma=sigma)
    profile = {"dtype": "float32", "height": ysize, "width": xsize, "count": 1, "driver": "GTiff",
               "nodata": float("nan"), "transform": (resolution, 0.0, xoff, 0.0, -resolution, yoff)}
    return profile, dsm
