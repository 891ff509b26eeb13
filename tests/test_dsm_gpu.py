"""GPU: the DSM splat (eogs_dsm_splat) against the plyflatten restatement of the oracle (sequential fp32 running
mean; PARITY UNPINNED, the package is absent) — same NaN mask, values within the fp32 rounding of that mean."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as O
from test_dsm_cpu import utm_cloud

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,res,sigma,radius", [(50000, 0.5, float("inf"), 1), (30000, 0.3, 0.4, 2),
                                                (1000, 0.5, float("inf"), 0), (1, 0.5, float("inf"), 1)])
def test_splat_matches_oracle(n, res, sigma, radius):
    from eogs2_b200.dsm import compute_dsm, dsm_grid
    c = utm_cloud(n, 3)
    profile, dsm = compute_dsm(torch.from_numpy(c).cuda(), res, radius=radius, sigma=sigma)
    xoff, yoff, w, h = dsm_grid(c[:, 0].min(), c[:, 0].max(), c[:, 1].min(), c[:, 1].max(), res)
    assert (profile["width"], profile["height"]) == (w, h) and profile["transform"][2] == xoff
    ref = O.plyflatten(c, xoff, yoff, res, w, h, radius, sigma)
    mine = dsm.cpu().numpy()
    assert mine.shape == ref.shape == (h, w, 1) and mine.dtype == np.float32
    assert np.array_equal(np.isnan(mine), np.isnan(ref))
    assert np.nanmax(np.abs(mine - ref)) < 1e-4          # metres; the reference's running mean is fp32


def test_full_size_nadir_render_properties():
    """2048^2 points (one per pixel of a nadir render): a constant-height cloud flattens to that constant
    wherever a point landed, and the number of reached cells matches the occupancy of the 3x3 footprints."""
    from eogs2_b200.dsm import compute_dsm
    n = 2048
    g = torch.Generator(device="cuda").manual_seed(0)
    u = (torch.arange(n, device="cuda", dtype=torch.float64) + 0.5) * 0.5
    xy = torch.stack(torch.meshgrid(4.3e5 + u, 3.36e6 - u, indexing="xy"), -1).reshape(-1, 2)
    xy = xy + (torch.rand(xy.shape, device="cuda", dtype=torch.float64, generator=g) - 0.5) * 0.2
    cloud = torch.cat([xy, torch.full((xy.shape[0], 1), 42.5, device="cuda", dtype=torch.float64)], 1)
    profile, dsm = compute_dsm(cloud, 0.5)
    ok = ~torch.isnan(dsm)
    assert float(ok.float().mean()) > 0.99
    assert torch.all(dsm[ok] == 42.5)
