"""CPU: the plyflatten restatement (oracle/eogs_oracle.c:oracle_plyflatten, PARITY UNPINNED — the package is
absent) against an independent numpy formulation (weighted mean per cell), and the grid arithmetic of
utils/dsm_utils.py:20-25."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as O


def numpy_flatten(cloud, xoff, yoff, res, w, h, radius, sigma):
    sw, swv = np.zeros((h, w)), np.zeros((h, w))
    i = np.trunc(w * (cloud[:, 0] - xoff) / (w * res)).astype(int)
    j = np.trunc(h * (-cloud[:, 1] + yoff) / (h * res)).astype(int)
    ok = (i >= 0) & (i < w) & (j >= 0) & (j < h)
    for k1 in range(-radius, radius + 1):
        for k2 in range(-radius, radius + 1):
            ii, jj = i + k1, j + k2
            m = ok & (ii >= 0) & (ii < w) & (jj >= 0) & (jj < h)
            dx = cloud[m, 0] - (xoff + res * (0.5 + ii[m])); dy = cloud[m, 1] - (yoff - res * (0.5 + jj[m]))
            wgt = np.exp(-(dx * dx + dy * dy) / (2 * sigma * sigma)) if np.isfinite(sigma) else np.ones(m.sum())
            np.add.at(sw, (jj[m], ii[m]), wgt); np.add.at(swv, (jj[m], ii[m]), wgt * cloud[m, 2])
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(sw > 0, swv / sw, np.nan)


def utm_cloud(n, seed):
    rng = np.random.default_rng(seed)
    xy = np.array([4.35e5, 3.355e6]) + rng.random((n, 2)) * np.array([120.0, 90.0])
    z = 10 + 5 * np.sin(xy[:, :1] / 7) + rng.normal(0, 0.1, (n, 1))
    return np.concatenate([xy, z], 1)


@pytest.mark.parametrize("sigma,radius", [(float("inf"), 1), (0.4, 2), (float("inf"), 0)])
def test_oracle_matches_numpy(sigma, radius):
    from eogs2_b200.dsm import dsm_grid
    c = utm_cloud(20000, 1)
    xoff, yoff, w, h = dsm_grid(c[:, 0].min(), c[:, 0].max(), c[:, 1].min(), c[:, 1].max(), 0.5)
    a = O.plyflatten(c, xoff, yoff, 0.5, w, h, radius, sigma)[:, :, 0]
    b = numpy_flatten(c, xoff, yoff, 0.5, w, h, radius, sigma)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.nanmax(np.abs(a - b)) < 2e-5 * 20


def test_grid_matches_reference_arithmetic():
    from eogs2_b200.dsm import dsm_grid
    res = 0.3
    xmin, xmax, ymin, ymax = 1000.07, 1033.4, 500.2, 512.9
    xoff, yoff, w, h = dsm_grid(xmin, xmax, ymin, ymax, res)
    assert xoff == np.floor(xmin / res) * res and yoff == np.ceil(ymax / res) * res      # utils/dsm_utils.py:22,24
    assert w == int(1 + np.floor((xmax - xoff) / res)) and h == int(1 - np.floor((ymin - yoff) / res))


def test_no_cpu_path():
    from eogs2_b200.dsm import plyflatten
    from eogs2_b200._cabi import EogsRasterError
    with pytest.raises(EogsRasterError):
        plyflatten(torch.zeros(4, 3, dtype=torch.float64), 0.0, 0.0, 0.5, 4, 4, 1, float("inf"))
