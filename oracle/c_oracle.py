"""numpy/ctypes front end of oracle/eogs_oracle.c (the scalar CPU restatement).

TEST INFRASTRUCTURE — importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  See the header of eogs_oracle.c for what is restated
(with reference file:line) and how it is pinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "eogs_oracle.c"
LIB = HERE / "libeogs_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        # -ffp-contract=off: every FMA in the oracle is an explicit fmaf(); gcc must not add others.
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", str(LIB), str(SRC), "-lm"]
        subprocess.run(cmd, check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(os.fspath(LIB))
        lib.oracle_preprocess.restype = C.c_longlong
        lib.oracle_bin.restype = C.c_int
        lib.oracle_higher_msb.restype = C.c_uint32
        _lib = lib
    return _lib


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return C.c_void_p(None) if a is None or a.size == 0 else a.ctypes.data_as(C.c_void_p)


def forward(means3D, scales, rotations, opacities, colors, view, bg, W, H, scale_modifier=1.0,
            antialiasing=False, cov3D_precomp=None) -> dict:
    """Whole forward; returns every intermediate in the reference's layouts."""
    lib = load()
    means3D, scales, rotations, colors, bg = map(_f32, (means3D, scales, rotations, colors, bg))
    opacities = _f32(opacities).reshape(-1)
    view = _f32(view).reshape(-1)
    cov3D_precomp = _f32(cov3D_precomp)
    P, Cc = means3D.shape[0], colors.shape[1]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    o = dict(P=P, W=W, H=H, C=Cc)
    o["radii"] = np.zeros(P, np.int32)
    o["means2D"] = np.zeros((P, 2), np.float32)
    o["depths"] = np.zeros(P, np.float32)
    o["conic_opacity"] = np.zeros((P, 4), np.float32)
    o["cov3D"] = np.zeros((P, 6), np.float32)
    o["tiles_touched"] = np.zeros(P, np.uint32)
    n = lib.oracle_preprocess(
        C.c_int(P), C.c_int(W), C.c_int(H), _p(means3D), _p(scales), _p(rotations), _p(cov3D_precomp),
        _p(opacities), _p(view), C.c_float(scale_modifier), C.c_int(int(antialiasing)), _p(o["radii"]),
        _p(o["means2D"]), _p(o["depths"]), _p(o["conic_opacity"]), _p(o["cov3D"]), _p(o["tiles_touched"]))
    if n < 0:
        raise RuntimeError("Point is too high: altitude above 200 (reference traps, forward.cu:267-272)")
    o["num_rendered"] = int(n)
    o["keys_sorted"] = np.zeros(n, np.uint64)
    o["point_list"] = np.zeros(n, np.uint32)
    o["ranges"] = np.zeros((gx * gy, 2), np.uint32)
    rc = lib.oracle_bin(C.c_int(P), C.c_int(W), C.c_int(H), C.c_longlong(n), _p(o["radii"]), _p(o["means2D"]),
                        _p(o["depths"]), _p(o["keys_sorted"]), _p(o["point_list"]), _p(o["ranges"]))
    if rc != 0:
        raise RuntimeError(f"oracle_bin failed rc={rc}")
    o["color"] = np.zeros((Cc, H, W), np.float32)
    o["invdepth"] = np.zeros((1, H, W), np.float32)
    o["final_T"] = np.zeros(H * W, np.float32)
    o["n_contrib"] = np.zeros(H * W, np.uint32)
    lib.oracle_blend_fwd(C.c_int(W), C.c_int(H), C.c_int(Cc), _p(o["ranges"]), _p(o["point_list"]), _p(o["means2D"]),
                         _p(o["conic_opacity"]), _p(colors), _p(o["depths"]), _p(bg), _p(o["color"]),
                         _p(o["invdepth"]), _p(o["final_T"]), _p(o["n_contrib"]))
    o["_inputs"] = dict(means3D=means3D, scales=scales, rotations=rotations, opacities=opacities, colors=colors,
                        view=view, bg=bg, scale_modifier=scale_modifier, antialiasing=antialiasing,
                        cov3D_precomp=cov3D_precomp)
    return o


def backward(o: dict, dL_dcolor, dL_dinvdepth=None, proj=None) -> dict:
    lib = load()
    i = o["_inputs"]
    P, W, H, Cc = o["P"], o["W"], o["H"], o["C"]
    dL_dcolor = _f32(dL_dcolor)
    dL_dinvdepth = _f32(dL_dinvdepth)
    proj = i["view"] if proj is None else _f32(proj).reshape(-1)
    d_mean2D = np.zeros((P, 2), np.float64)
    d_conic = np.zeros((P, 4), np.float64)
    d_op = np.zeros(P, np.float64)
    d_col = np.zeros((P, Cc), np.float64)
    d_invd = np.zeros(P, np.float64)
    lib.oracle_blend_bwd(C.c_int(P), C.c_int(W), C.c_int(H), C.c_int(Cc), _p(o["ranges"]), _p(o["point_list"]),
                         _p(o["means2D"]), _p(o["conic_opacity"]), _p(i["colors"]), _p(o["depths"]), _p(i["bg"]),
                         _p(o["final_T"]), _p(o["n_contrib"]), _p(dL_dcolor), _p(dL_dinvdepth),
                         _p(d_mean2D), _p(d_conic), _p(d_op), _p(d_col), _p(d_invd))
    g = dict(dL_dmeans2D=np.concatenate([d_mean2D, np.zeros((P, 1))], 1).astype(np.float32),
             dL_dconic=d_conic.astype(np.float32), dL_dcolors=d_col.astype(np.float32),
             dL_dinvdepths=d_invd.astype(np.float32))
    d_op32 = d_op.astype(np.float32)
    m2 = np.ascontiguousarray(d_mean2D.astype(np.float32))
    g["dL_dmeans3D"] = np.zeros((P, 3), np.float32)
    g["dL_dcov3D"] = np.zeros((P, 6), np.float32)
    g["dL_dscales"] = np.zeros((P, 3), np.float32)
    g["dL_drotations"] = np.zeros((P, 4), np.float32)
    g["dL_dT"] = np.zeros((P, 6), np.float32)
    lib.oracle_pre_bwd(C.c_int(P), C.c_int(W), C.c_int(H), _p(i["means3D"]), _p(i["scales"]), _p(i["rotations"]),
                       _p(i["cov3D_precomp"]), _p(i["opacities"]), _p(i["view"]), _p(proj),
                       C.c_float(i["scale_modifier"]), C.c_int(int(i["antialiasing"])), _p(o["radii"]), _p(m2),
                       _p(g["dL_dconic"]), _p(d_op32), _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]),
                       _p(g["dL_dscales"]), _p(g["dL_drotations"]), _p(g["dL_dT"]))
    g["dL_dopacity"] = d_op32.reshape(P, 1)
    g["grad_viewmatrix"] = grad_viewmatrix(g["dL_dT"], g["dL_dmeans2D"], i["means3D"], W, H)
    return g


def grad_viewmatrix(dL_dT, dL_dmeans2D, means3D, W, H) -> np.ndarray:
    """DGR/diff_gaussian_rasterization/__init__.py:172-202, in float64, with dL_dT [P,6] at the
    intended per-Gaussian stride."""
    T = dL_dT.astype(np.float64).reshape(-1, 2, 3)
    N = np.eye(3); N[0, 0] = W / 2; N[1, 1] = H / 2
    dL_dA = (N @ T.transpose(0, 2, 1)).sum(0)
    g = np.zeros((4, 4))
    g[:3, :2] += dL_dA
    g[:3, :3] += means3D.astype(np.float64).T @ dL_dmeans2D.astype(np.float64)
    g[-1, :3] += dL_dmeans2D.astype(np.float64).sum(0)
    return g


def dist2(points) -> np.ndarray:
    """distCUDA2 (simple-knn spatial.cu:15-26) by brute force: O(P^2), keep P in the tens of thousands."""
    lib = load()
    pts = _f32(points).reshape(-1, 3)
    out = np.zeros(pts.shape[0], np.float32)
    lib.oracle_dist2(C.c_int(pts.shape[0]), _p(pts), _p(out))
    return out


def plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius, sigma) -> np.ndarray:
    """plyflatten(...) as called by utils/dsm_utils.py:28-37 -> [ysize, xsize, 1] float32.  PARITY UNPINNED
    (the package is absent; see the header of eogs_oracle.c)."""
    lib = load()
    c = np.ascontiguousarray(np.asarray(cloud, dtype=np.float64)).reshape(-1, 3)
    out = np.zeros((ysize, xsize, 1), np.float32)
    lib.oracle_plyflatten(C.c_longlong(c.shape[0]), _p(c), C.c_double(xoff), C.c_double(yoff), C.c_double(resolution),
                          C.c_int(xsize), C.c_int(ysize), C.c_int(radius), C.c_float(sigma), _p(out))
    return out
