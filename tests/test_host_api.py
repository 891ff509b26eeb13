"""CPU: host-side logic of the API mirror — argument validation and error behaviour identical to the
reference (DGR/diff_gaussian_rasterization/__init__.py:263-275), loud failure without CUDA, and the
grad_viewmatrix assembly against the reference's torch formula (__init__.py:172-202)."""
import pytest
import torch

import eogs2_b200 as E
from eogs2_b200 import _cabi


def settings(view=None, H=32, W=32):
    view = torch.eye(4) if view is None else view
    return E.GaussianRasterizationSettings(H, W, 1.0, 1.0, torch.zeros(5), 1.0, view, view, 0, torch.zeros(3),
                                           False, False, False)


def test_exactly_one_of_shs_or_colors():
    r = E.GaussianRasterizer(settings())
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(m, m, torch.ones(4, 1), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(m, m, torch.ones(4, 1), shs=torch.ones(4, 1, 3), colors_precomp=torch.ones(4, 5),
          scales=torch.ones(4, 3), rotations=torch.ones(4, 4))


def test_exactly_one_of_scale_rotation_or_cov():
    r = E.GaussianRasterizer(settings())
    m = torch.zeros(4, 3)
    msg = "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!"
    with pytest.raises(Exception, match=msg):
        r(m, m, torch.ones(4, 1), colors_precomp=torch.ones(4, 5))
    with pytest.raises(Exception, match=msg):
        r(m, m, torch.ones(4, 1), colors_precomp=torch.ones(4, 5), scales=torch.ones(4, 3))
    with pytest.raises(Exception, match=msg):
        r(m, m, torch.ones(4, 1), colors_precomp=torch.ones(4, 5), scales=torch.ones(4, 3),
          rotations=torch.ones(4, 4), cov3D_precomp=torch.ones(4, 6))


def test_cpu_tensors_fail_loudly_no_fallback():
    r = E.GaussianRasterizer(settings())
    m = torch.zeros(4, 3)
    with pytest.raises(_cabi.EogsRasterError, match="no CPU path"):
        r(m, m, torch.ones(4, 1), colors_precomp=torch.ones(4, 5), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))


def test_means3D_shape_check():
    with pytest.raises(RuntimeError, match=r"means3D must have dimensions \(num_points, 3\)"):
        E.rasterize_forward_raw(torch.zeros(5), torch.zeros(4, 2), torch.ones(4, 5), torch.ones(4, 1),
                                torch.ones(4, 3), torch.ones(4, 4), 1.0, torch.empty(0), torch.eye(4), 32, 32)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", tmp_path / "libeogs_raster.so")
    with pytest.raises(_cabi.EogsRasterError, match="no CPU or PyTorch fallback"):
        _cabi.load()


def test_grad_viewmatrix_assembly_matches_reference_formula():
    torch.manual_seed(0)
    P, W, H = 257, 640, 480
    grad_T = torch.randn(P, 6)
    grad_means2D = torch.cat([torch.randn(P, 2), torch.zeros(P, 1)], 1)
    means3D = torch.randn(P, 3)
    view = torch.randn(4, 4)
    # reference: __init__.py:172-202
    ref = torch.zeros_like(view)
    N = torch.eye(3); N[0, 0] = W / 2; N[1, 1] = H / 2
    ref[:3, :2] += (N @ grad_T.view(P, 2, 3).transpose(1, 2)).sum(axis=0)
    ref[:3, :3] += means3D.T @ grad_means2D
    ref[-1, :3] += grad_means2D.sum(axis=0)
    cam_sums = torch.zeros(16)
    cam_sums[0:6] = grad_T.sum(0)
    cam_sums[6:12] = (means3D.T @ grad_means2D[:, :2]).reshape(-1)
    cam_sums[12:14] = grad_means2D[:, :2].sum(0)
    mine = E.assemble_grad_viewmatrix(cam_sums, view, W, H)
    assert torch.allclose(mine, ref, rtol=1e-5, atol=1e-4)


def test_product_packages_never_import_the_oracle():
    """oracle/ is test infrastructure: nothing under eogs2_b200/ or the two shim packages may import it."""
    import ast
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    for pkg in ("eogs2_b200", "diff_gaussian_rasterization", "simple_knn"):
        for py in (root / pkg).rglob("*.py"):
            tree = ast.parse(py.read_text())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                assert not any(n == "oracle" or n.startswith("oracle.") for n in names), (py, names)
