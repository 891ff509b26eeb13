"""ctypes driver of oracle/_ref/libeogs_ref.so — the reference CUDA rasterizer compiled for
sm_100a from /root/reference by oracle/ref_build/Makefile.

TEST INFRASTRUCTURE.  Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and
`bench.py --impl reference` may import this; nothing under eogs2_b200/ does.

It mirrors what DGR/rasterize_points.cu does around CudaRasterizer::Rasterizer (allocate
outputs and the three growable byte buffers, marshal pointers) and what
DGR/diff_gaussian_rasterization/__init__.py:113-216 does after the backward call (the
grad_viewmatrix assembly), so that a test can call the reference exactly as EOGS++ does.
The reference launches on the legacy default stream: call it with torch's default stream current.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

LIB_PATH = Path(__file__).resolve().parent / "_ref" / "libeogs_ref.so"
# The same sources with the one-line `dL_dT[6*idx+k]` stride fix of backward.cu:320-325 (oracle/ref_build/Makefile):
# the race-free reference value of the camera-covariance term of grad_viewmatrix.
STRIDEFIX_PATH = LIB_PATH.with_name("libeogs_ref_stridefix.so")
NUM_CHANNELS = 5
_ALLOC_T = C.CFUNCTYPE(C.c_void_p, C.c_int, C.c_size_t)
_libs: dict = {}


def available(stridefix: bool = False) -> bool:
    return (STRIDEFIX_PATH if stridefix else LIB_PATH).exists()


def load(stridefix: bool = False) -> C.CDLL:
    path = STRIDEFIX_PATH if stridefix else LIB_PATH
    lib = _libs.get(path)
    if lib is None:
        if not path.exists():
            raise RuntimeError(f"{path} missing: run `make -C oracle/ref_build` where /root/reference exists")
        lib = C.CDLL(os.fspath(path))
        lib.eogs_ref_forward.restype = C.c_int
        lib.eogs_ref_backward.restype = C.c_int
        _libs[path] = lib
    return lib


def _p(t):
    return C.c_void_p(None) if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())


class RefState:
    pass


def forward(bg, means3D, colors, opacities, scales, rotations, scale_modifier, cov3D_precomp,
            viewmatrix, projmatrix, tanfovx, tanfovy, H, W, campos, prefiltered=False,
            antialiasing=False, debug=False) -> RefState:
    """_C.rasterize_gaussians (DGR/rasterize_points.cu:35-124)."""
    lib = load()
    dev = means3D.device
    P = means3D.shape[0]
    st = RefState()
    st.P, st.H, st.W = P, H, W
    st.color = torch.zeros((NUM_CHANNELS, H, W), dtype=torch.float32, device=dev)
    st.invdepth = torch.zeros((1, H, W), dtype=torch.float32, device=dev)
    st.radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    st.bufs = {}

    def alloc(which, n):
        t = torch.empty(max(int(n), 1), dtype=torch.uint8, device=dev)
        st.bufs[which] = t
        return t.data_ptr()

    cb = _ALLOC_T(alloc)
    err = C.create_string_buffer(512)
    st.num_rendered = 0
    if P != 0:
        args = [means3D, colors, opacities, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, campos, bg]
        means3D, colors, opacities, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, campos, bg = [
            a.contiguous() if a is not None else None for a in args]
        rc = lib.eogs_ref_forward(
            C.c_int(P), _p(bg), C.c_int(W), C.c_int(H), _p(means3D), _p(colors), _p(opacities), _p(scales),
            C.c_float(scale_modifier), _p(rotations), _p(cov3D_precomp), _p(viewmatrix), _p(projmatrix),
            _p(campos), C.c_float(tanfovx), C.c_float(tanfovy), C.c_int(int(prefiltered)), _p(st.color),
            _p(st.invdepth), C.c_int(int(antialiasing)), _p(st.radii), C.c_int(int(debug)), cb, err, C.c_int(512))
        if rc < 0:
            raise RuntimeError(err.value.decode())
        st.num_rendered = rc
    return st


def backward(st: RefState, bg, means3D, colors, opacities, scales, rotations, scale_modifier,
             cov3D_precomp, viewmatrix, projmatrix, tanfovx, tanfovy, dL_dcolor, dL_dinvdepth, campos,
             antialiasing=False, debug=False, stridefix=False) -> dict:
    """_C.rasterize_gaussians_backward (DGR/rasterize_points.cu:126-224).  stridefix=True runs the backward of
    the stride-fixed build on the same forward state (identical layouts; only the dL_dT store differs)."""
    lib = load(stridefix)
    dev = means3D.device
    P, H, W = st.P, st.H, st.W
    z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
    out = dict(dL_dmeans3D=z(P, 3), dL_dmeans2D=z(P, 3), dL_dcolors=z(P, NUM_CHANNELS), dL_dconic=z(P, 2, 2),
               dL_dopacity=z(P, 1), dL_dcov3D=z(P, 6), dL_dsh=z(P, 0, 3), dL_dscales=z(P, 3),
               dL_drotations=z(P, 4), dL_dinvdepths=z(P, 1), dL_dT=z(P, 6))
    if P != 0:
        err = C.create_string_buffer(512)
        args = [means3D, colors, opacities, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, campos, bg,
                dL_dcolor, dL_dinvdepth]
        (means3D, colors, opacities, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, campos, bg,
         dL_dcolor, dL_dinvdepth) = [a.contiguous() if a is not None else None for a in args]
        rc = lib.eogs_ref_backward(
            C.c_int(P), C.c_int(st.num_rendered), _p(bg), C.c_int(W), C.c_int(H), _p(means3D), _p(colors),
            _p(opacities), _p(scales), C.c_float(scale_modifier), _p(rotations), _p(cov3D_precomp),
            _p(viewmatrix), _p(projmatrix), _p(campos), C.c_float(tanfovx), C.c_float(tanfovy), _p(st.radii),
            _p(st.bufs[0]), _p(st.bufs[1]), _p(st.bufs[2]), _p(dL_dcolor), _p(dL_dinvdepth),
            _p(out["dL_dmeans2D"]), _p(out["dL_dconic"]), _p(out["dL_dopacity"]), _p(out["dL_dcolors"]),
            _p(out["dL_dinvdepths"]), _p(out["dL_dmeans3D"]), _p(out["dL_dcov3D"]), _p(out["dL_dsh"]),
            _p(out["dL_dscales"]), _p(out["dL_drotations"]), C.c_int(int(antialiasing)), C.c_int(int(debug)),
            _p(out["dL_dT"]), err, C.c_int(512))
        if rc < 0:
            raise RuntimeError(err.value.decode())
    return out


def grad_viewmatrix_terms(out: dict, means3D, viewmatrix, H, W) -> dict:
    """The three terms of grad_viewmatrix, DGR/diff_gaussian_rasterization/__init__.py:172-202,
    returned separately because the first one inherits the reference's dL_dT stride race
    (backward.cu:320-325) and cannot be compared bit for bit."""
    grad_T, grad_means2D = out["dL_dT"], out["dL_dmeans2D"]
    B = grad_T.shape[0]
    T_example = grad_T.view(B, 2, 3)
    N = torch.eye(3, device=grad_T.device)
    N[0, 0] = W / 2
    N[1, 1] = H / 2
    dL_dA = (N @ T_example.transpose(1, 2)).sum(axis=0)
    t1 = torch.zeros_like(viewmatrix); t1[:3, :2] += dL_dA
    t2 = torch.zeros_like(viewmatrix); t2[:3, :3] += means3D.T @ grad_means2D
    t3 = torch.zeros_like(viewmatrix); t3[-1, :3] += grad_means2D.sum(axis=0)
    return dict(cov_term=t1, mean_term=t2, bias_term=t3, total=t1 + t2 + t3)


def export_state(st: RefState) -> dict:
    """Fields of geomState / binningState / imgState (rasterizer_impl.h:29-62) as tensors."""
    lib = load()
    P, R, N = st.P, st.num_rendered, st.W * st.H
    tiles = ((st.W + 15) // 16) * ((st.H + 15) // 16)
    g, b, im = st.bufs[0], st.bufs.get(1), st.bufs[2]

    def view(buf, off, count, dtype):
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        return buf[off:off + nbytes].view(dtype).clone()

    gl = (C.c_uint64 * 7)(); lib.eogs_ref_geom_layout(C.c_void_p(g.data_ptr()), C.c_size_t(P), gl)
    il = (C.c_uint64 * 3)(); lib.eogs_ref_image_layout(C.c_void_p(im.data_ptr()), C.c_size_t(N), il)
    out = {
        "depths": view(g, gl[0], P, torch.float32),
        "means2D": view(g, gl[2], 2 * P, torch.float32).view(P, 2),
        "cov3D": view(g, gl[3], 6 * P, torch.float32).view(P, 6),
        "conic_opacity": view(g, gl[4], 4 * P, torch.float32).view(P, 4),
        "tiles_touched": view(g, gl[5], P, torch.int32),
        "point_offsets": view(g, gl[6], P, torch.int32),
        "final_T": view(im, il[0], N, torch.float32),
        "n_contrib": view(im, il[1], N, torch.int32),
        "ranges": view(im, il[2], 2 * tiles, torch.int32).view(tiles, 2),
        "radii": st.radii,
    }
    if R > 0 and b is not None:
        bl = (C.c_uint64 * 4)(); lib.eogs_ref_binning_layout(C.c_void_p(b.data_ptr()), C.c_size_t(R), bl)
        out["point_list"] = view(b, bl[0], R, torch.int32)
        out["keys_sorted"] = view(b, bl[2], R, torch.int64)
    else:
        out["point_list"] = torch.zeros(0, dtype=torch.int32, device=g.device)
        out["keys_sorted"] = torch.zeros(0, dtype=torch.int64, device=g.device)
    return out
