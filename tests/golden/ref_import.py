"""Import the REFERENCE's own Python callers of the rasterizer (gaussian_renderer/renderer.py,
renderer_cc_shadow.py, scene/gaussian_model.py, scene/cameras/affine_cameras.py) on the CPU of the build
container, to generate golden vectors ("a Python reference can be imported in THIS
container ... commit the vectors as small fixtures together with the script that made them").

TEST INFRASTRUCTURE.  Nothing here is product code and nothing is copied from the reference: the modules
are imported from where they lie under /root/reference.  Three things stand between them and a CPU-only
container, all handled with stand-ins registered in sys.modules / module globals for the duration of the
import:

  * `diff_gaussian_rasterization` is a CUDA extension.  The stand-in exposes the same
    GaussianRasterizationSettings / GaussianRasterizer / autograd.Function surface (DGR __init__.py:53-300)
    on top of oracle/eogs_oracle.c — the scalar C restatement of the reference kernels, forward and
    hand-written backward, that tests/test_oracle_golden.py pins to the compiled reference.  The screen-space
    gradient comes back through `means2D` and the camera gradient through `viewmatrix`, as in
    __init__.py:172-214 (with the intended 6*idx stride of dL_dT).
  * `simple_knn._C`, `plyfile`, `arguments` (hydra / omegaconf) are absent: minimal stubs (never called on
    the paths exercised here, except GroupParams as a type name).
  * the callers hard-code device="cuda" (renderer.py:33, affine_cameras.py:356,376, general_utils.py:109): the `torch` name seen by
    those modules is a proxy that maps that device string to "cpu" for tensor factories.

`scene/__init__.py` pulls the dataset readers (rasterio, rpcm, ...): a namespace stand-in for the `scene`
package lets `scene.gaussian_model` / `scene.cameras.affine_cameras` import without executing it.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path
from typing import NamedTuple

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
REF_SRC = Path("/root/reference/src/gaussiansplatting")


def available() -> bool:
    return (REF_SRC / "gaussian_renderer" / "renderer.py").exists()


# ---------------------------------------------------------------------------------------------
# CPU stand-in for the CUDA extension
class GaussianRasterizationSettings(NamedTuple):      # DGR __init__.py:219-232, same field order
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


class _RasterizeGaussians(torch.autograd.Function):
    """The stand-in's counterpart of DGR __init__.py:53-216: forward AND backward are the reference kernels' own
    arithmetic as restated in oracle/eogs_oracle.c (pinned bit for bit / 1e-4 / 1e-3 to the compiled reference by
    tests/test_oracle_golden.py) — not autograd through the forward: the reference's hand-written backward differs
    from the true derivative in documented places (scale gradient not multiplied by scale_modifier,
    backward.cu:385-393; antialiasing compensation evaluated at the dilated covariance, :222-231), and the callers'
    gradients inherit exactly that."""

    @staticmethod
    def forward(ctx, means3D, means2D, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, viewmatrix, rs):
        if str(ROOT) not in sys.path:
            sys.path.insert(0, str(ROOT))
        from oracle import c_oracle as O
        np_ = lambda t: None if t is None else t.detach().cpu().numpy()
        o = O.forward(np_(means3D), np_(scales), np_(rotations), np_(opacities), np_(colors_precomp), np_(viewmatrix),
                      np_(rs.bg), int(rs.image_width), int(rs.image_height), float(rs.scale_modifier),
                      bool(rs.antialiasing), np_(cov3Ds_precomp))
        ctx.o, ctx.O, ctx.rs = o, O, rs
        ctx.has = (scales is not None, cov3Ds_precomp is not None)
        ctx.shapes = (opacities.shape, )
        radii = torch.from_numpy(o["radii"].copy())
        ctx.mark_non_differentiable(radii)
        return torch.from_numpy(o["color"].copy()), radii, torch.from_numpy(o["invdepth"].copy())

    @staticmethod
    def backward(ctx, grad_color, _, grad_invdepth):
        o, O = ctx.o, ctx.O
        z = lambda t, like: torch.zeros_like(like) if t is None else t
        g = O.backward(o, z(grad_color, torch.from_numpy(o["color"])).numpy(),
                       z(grad_invdepth, torch.from_numpy(o["invdepth"])).numpy())
        t = lambda k: torch.from_numpy(np.ascontiguousarray(g[k], dtype=np.float32))
        has_sr, has_cov = ctx.has
        return (t("dL_dmeans3D"), t("dL_dmeans2D"), t("dL_dcolors"), t("dL_dopacity").reshape(ctx.shapes[0]),
                t("dL_dscales") if has_sr else None, t("dL_drotations") if has_sr else None,
                t("dL_dcov3D") if has_cov else None,
                torch.from_numpy(g["grad_viewmatrix"].astype(np.float32)), None)


class GaussianRasterizer(torch.nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        return _RasterizeGaussians.apply(means3D, means2D, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                         rs.viewmatrix, rs)


def _rasterizer_module() -> types.ModuleType:
    m = types.ModuleType("diff_gaussian_rasterization")
    m.GaussianRasterizationSettings = GaussianRasterizationSettings
    m.GaussianRasterizer = GaussianRasterizer
    m.__doc__ = "CPU stand-in (oracle/eogs_oracle.c) for golden-vector generation"
    return m


class _TorchProxy:
    """`torch` as seen by the reference modules: tensor factories asked for device="cuda" get the CPU."""
    _FACTORIES = ("zeros_like", "ones_like", "zeros", "ones", "eye", "tensor", "empty", "full", "randn", "rand",
                  "linspace", "arange")

    def __init__(self):
        self._t = torch

    def __getattr__(self, name):
        attr = getattr(self._t, name)
        if name in self._FACTORIES:
            def wrapped(*a, **k):
                if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
                    k["device"] = "cpu"
                return attr(*a, **k)
            return wrapped
        return attr


_loaded = None


def load():
    """Returns a namespace with the reference's render, render_resample_virtual_camera, GaussianModel,
    AffineCamera, SunCamera (the real objects, imported from /root/reference)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"{REF_SRC} is not present (golden vectors are generated in the build container only)")
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in ("diff_gaussian_rasterization", "simple_knn", "simple_knn._C", "plyfile",
                                                  "arguments", "scene", "scene.cameras", "utils", "gaussian_renderer")}
    try:
        sys.path.insert(0, str(REF_SRC))
        for k in ("utils", "gaussian_renderer", "scene", "scene.cameras"):
            sys.modules.pop(k, None)
        sys.modules["diff_gaussian_rasterization"] = _rasterizer_module()
        knn = types.ModuleType("simple_knn"); knn_c = types.ModuleType("simple_knn._C")
        knn_c.distCUDA2 = lambda pts: (_ for _ in ()).throw(RuntimeError("distCUDA2 stub"))
        knn._C = knn_c
        sys.modules["simple_knn"], sys.modules["simple_knn._C"] = knn, knn_c
        ply = types.ModuleType("plyfile"); ply.PlyData = ply.PlyElement = object
        sys.modules["plyfile"] = ply
        args = types.ModuleType("arguments"); args.GroupParams = type("GroupParams", (), {})
        sys.modules["arguments"] = args
        for pkg, sub in (("scene", "scene"), ("scene.cameras", "scene/cameras")):
            m = types.ModuleType(pkg)
            m.__path__ = [str(REF_SRC / sub)]
            sys.modules[pkg] = m
        gm = importlib.import_module("scene.gaussian_model")
        ac = importlib.import_module("scene.cameras.affine_cameras")
        renderer = importlib.import_module("gaussian_renderer.renderer")
        shadow = importlib.import_module("gaussian_renderer.renderer_cc_shadow")
        proxy = _TorchProxy()
        for mod in (renderer, ac, gm, importlib.import_module("utils.general_utils")):
            mod.torch = proxy
        _loaded = types.SimpleNamespace(render=renderer.render,
                                        render_resample_virtual_camera=shadow.render_resample_virtual_camera,
                                        GaussianModel=gm.GaussianModel, AffineCamera=ac.AffineCamera,
                                        SunCamera=ac.SunCamera, gaussian_model=gm, affine_cameras=ac,
                                        renderer=renderer)
        return _loaded
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            if k in ("scene", "scene.cameras", "utils", "gaussian_renderer"):
                continue                       # the reference's own packages stay importable for its lazy imports
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
