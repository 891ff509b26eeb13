"""CPU: the distCUDA2 oracle (oracle/eogs_oracle.c: oracle_dist2) reproduces the reference's outputs
(tests/golden/knn_ref.npz, made by tests/golden/make_golden_knn.py with the compiled reference on a B200)
bit for bit; the host mirror has the reference's import path and refuses CPU tensors."""
from pathlib import Path

import numpy as np
import pytest
import torch

from knn_cases import GOLDEN_CASES, points
from oracle import c_oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden" / "knn_ref.npz"


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_matches_reference_golden(name):
    g = np.load(GOLDEN)
    kind, P, seed = GOLDEN_CASES[name]
    p = g[f"{name}_points"]
    assert np.array_equal(p, points(kind, P, seed)), "fixture inputs no longer match the seeded generator"
    mine = O.dist2(p)
    assert np.array_equal(mine.view(np.uint32), g[f"{name}_dist2"].view(np.uint32))


def test_oracle_agrees_with_float64_kdtree():
    from scipy.spatial import cKDTree
    p = points("uniform", 4000, 3)
    d = O.dist2(p)
    dd, _ = cKDTree(p.astype(np.float64)).query(p.astype(np.float64), k=4)
    assert np.allclose(d, (dd[:, 1:] ** 2).mean(1), rtol=1e-5, atol=0)


def test_fewer_than_four_points_keep_the_reference_sentinel():
    # simple_knn.cu:26,157: best[] starts at 1E+37 and missing neighbours stay there
    assert O.dist2(points("uniform", 1, 1))[0] == np.float32(np.float32(np.float32(1e37) * 2) + np.float32(1e37)) / np.float32(3)
    d = O.dist2(points("uniform", 3, 3))
    assert (d > 3e36).all()


def test_host_mirror_has_the_reference_import_path_and_no_cpu_path():
    from simple_knn._C import distCUDA2          # scene/gaussian_model.py:21
    from eogs2_b200._cabi import EogsRasterError
    with pytest.raises(EogsRasterError):
        distCUDA2(torch.zeros(10, 3))
