"""distCUDA2 — mean squared distance to the 3 nearest neighbours.

Mirror of `simple_knn._C.distCUDA2` (submodules/simple-knn/ext.cpp:15-17, spatial.cu:15-26), which
`GaussianModel.create_from_pcd` uses to initialise the scales (scene/gaussian_model.py:179-186:
`clamp_min(distCUDA2(points), 1e-7)` -> `log(sqrt(.))`).  Same signature: a float CUDA tensor `[P, 3]` in,
a float32 tensor `[P]` out, bit-identical to the reference's values (csrc/knn.cu explains why a different
search structure can be).  Runs on torch's current stream without a host synchronisation.  No CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if not isinstance(points, torch.Tensor) or not points.is_cuda:
        raise _cabi.EogsRasterError("distCUDA2 needs a CUDA tensor: there is no CPU path")
    if points.dim() != 2 or points.shape[1] != 3:
        raise _cabi.EogsRasterError("points must have dimensions (num_points, 3)")
    lib = _cabi.load()
    # spatial.cu:23 reinterprets points.contiguous().data<float>(): a non-float tensor is an error there too
    if points.dtype != torch.float32:
        raise _cabi.EogsRasterError("expected scalar type Float")
    pts = points.detach().contiguous()
    P = int(pts.shape[0])
    with torch.cuda.device(pts.device):
        means = torch.full((P,), 0.0, dtype=torch.float32, device=pts.device)
        if P == 0:
            return means
        nbytes = int(lib.eogs_knn_bytes(P))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=pts.device)
        stream = torch.cuda.current_stream(pts.device).cuda_stream
        _cabi.check(lib.eogs_knn_dist2(C.c_void_p(stream), P, C.c_void_p(pts.data_ptr()),
                                       C.c_void_p(scratch.data_ptr()), C.c_size_t(nbytes),
                                       C.c_void_p(means.data_ptr())), "eogs_knn_dist2")
    return means
