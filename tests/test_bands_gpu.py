"""GPU: tile-band sharding of one view (eogs2_b200/bands.py, the *_band C-ABI entry points).

Rendering the bands one after the other on one GPU must reproduce the whole-image render bit for
bit: images, final_T, n_contrib, and — per band — the sorted (tile | depth) keys, the Gaussian
list and the tile ranges (shifted by the band's list base).  Band gradients must sum to the
whole-image gradients (1e-3 relative: float-atomic order differs)."""
import numpy as np
import pytest
import torch

import eogs2_b200 as E
from eogs2_b200 import bands as B
from eogs2_b200 import scene as S

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu().numpy(), b.double().cpu().numpy()
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def case(dev, P, W, H, seed):
    sc = S.make_scene(P, "trained", seed)
    view = S.make_camera(seed)
    dcol, dinv = S.upstream_grads(5, H, W, seed, True)
    t = dict(means3D=sc.means3D, scales=sc.scales, rotations=sc.rotations, opacities=sc.opacities,
             colors=S.colors_precomp(sc, view), view=view, bg=S.background(seed), dcol=dcol, dinv=dinv)
    return {k: v.to(dev) for k, v in t.items()}


def fwd(c, W, H, band=None):
    empty = torch.empty(0, device=c["means3D"].device)
    return E.rasterize_forward_raw(c["bg"], c["means3D"], c["colors"], c["opacities"], c["scales"], c["rotations"],
                                   1.0, empty, c["view"], H, W, False, False, band=band)


def bwd(c, st, dcol, dinv):
    empty = torch.empty(0, device=c["means3D"].device)
    return E.rasterize_backward_raw(st, c["bg"], c["means3D"], c["colors"], c["opacities"], c["scales"],
                                    c["rotations"], 1.0, empty, c["view"], c["view"], dcol, dinv)


@pytest.mark.parametrize("P,W,H,seed,world,weighted", [
    (30_000, 400, 300, 11, 3, False),        # ragged: H not a multiple of 16, uneven band heights
    (50_000, 512, 512, 1337, 8, False),      # BASELINE configs[0] geometry split 8 ways
    (30_000, 333, 517, 12, 4, True),         # weight-balanced bands
])
def test_bands_reproduce_the_whole_image(cuda_dev, P, W, H, seed, world, weighted):
    c = case(cuda_dev, P, W, H, seed)
    full = fwd(c, W, H)
    exf = E.export_state(full)
    gfull = bwd(c, full, c["dcol"], c["dinv"])
    grid_x, grid_y = (W + 15) // 16, (H + 15) // 16
    weights = B.row_weights_from_state(full) if weighted else None
    bands = B.split_rows(grid_y, world, weights)
    assert bands[0][0] == 0 and bands[-1][1] == grid_y and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))

    ranges_full = exf["ranges"].cpu().numpy().astype(np.int64)
    keys_full = exf["keys_sorted"].cpu().numpy()
    list_full = exf["point_list"].cpu().numpy()
    total_instances = 0
    gsum = None
    for rb, re in bands:
        st = fwd(c, W, H, band=(rb, re))
        ex = E.export_state(st)
        y0, h = 16 * rb, st.band_height
        assert st.color.shape == (5, h, W)
        assert torch.equal(st.radii, full.radii)                       # whole-image meaning on every rank
        assert torch.equal(st.color, full.color[:, y0:y0 + h])
        assert torch.equal(st.invdepth, full.invdepth[:, y0:y0 + h])
        assert torch.equal(ex["final_T"], exf["final_T"].view(H, W)[y0:y0 + h].reshape(-1))
        assert torch.equal(ex["n_contrib"], exf["n_contrib"].view(H, W)[y0:y0 + h].reshape(-1))
        # the band's list = the whole-image list restricted to the band's tiles
        t0, t1 = rb * grid_x, re * grid_x
        r = ranges_full[t0:t1]
        nonempty = r[r[:, 1] > r[:, 0]]
        base = int(nonempty[:, 0].min()) if len(nonempty) else 0
        end = int(nonempty[:, 1].max()) if len(nonempty) else 0
        assert st.num_rendered == end - base
        assert np.array_equal(ex["keys_sorted"].cpu().numpy(), keys_full[base:end])
        assert np.array_equal(ex["point_list"].cpu().numpy(), list_full[base:end])
        rbnd = ex["ranges"].cpu().numpy().astype(np.int64)
        live = rbnd[:, 1] > rbnd[:, 0]
        assert np.array_equal(live, r[:, 1] > r[:, 0])
        assert np.array_equal(rbnd[live] + base, r[live])
        total_instances += st.num_rendered
        g = bwd(c, st, c["dcol"][:, y0:y0 + h].contiguous(), c["dinv"].reshape(H, W)[y0:y0 + h].contiguous())
        gsum = [x.clone() for x in g if x is not None] if gsum is None else \
            [a + b for a, b in zip(gsum, [x for x in g if x is not None])]
    assert total_instances == full.num_rendered
    for a, b in zip(gsum, [x for x in gfull if x is not None]):
        assert rel(a, b) < 1e-3


def test_single_rank_helpers_match_the_plain_call(cuda_dev):
    W, H = 320, 240
    c = case(cuda_dev, 20_000, W, H, 21)
    empty = torch.empty(0, device=cuda_dev)
    color, invd, st = B.forward_band(c["bg"], c["means3D"], c["colors"], c["opacities"], c["scales"], c["rotations"],
                                     1.0, empty, c["view"], H, W, rank=0, world=1)
    full = fwd(c, W, H)
    assert torch.equal(color, full.color) and torch.equal(invd, full.invdepth)
    g = B.backward_band(st, c["bg"], c["means3D"], c["colors"], c["opacities"], c["scales"], c["rotations"], 1.0,
                        empty, c["view"], c["view"], c["dcol"], c["dinv"])
    gf = bwd(c, full, c["dcol"], c["dinv"])
    for a, b in zip(g, gf):
        if a is not None:
            assert rel(a, b) < 1e-3


def test_bad_band_is_rejected(cuda_dev):
    c = case(cuda_dev, 1000, 64, 64, 3)
    with pytest.raises(Exception):
        fwd(c, 64, 64, band=(2, 9))
    with pytest.raises(Exception):
        fwd(c, 64, 64, band=(3, 3))


def test_forward_strips_equals_the_single_call_above_65536_tiles(cuda_dev):
    """4112 x 4112 = 257 x 257 tiles (> 65 536: the single call sorts 32-bit keys, the strips 16-bit ones)."""
    W = H = 4112
    c = case(cuda_dev, 20_000, W, H, 31)
    full = fwd(c, W, H)
    empty = torch.empty(0, device=cuda_dev)
    color, invd, radii = B.forward_strips(c["bg"], c["means3D"], c["colors"], c["opacities"], c["scales"],
                                          c["rotations"], 1.0, empty, c["view"], H, W)
    assert torch.equal(color, full.color) and torch.equal(invd, full.invdepth) and torch.equal(radii, full.radii)
    color2, _, _ = B.forward_strips(c["bg"], c["means3D"], c["colors"], c["opacities"], c["scales"], c["rotations"],
                                    1.0, empty, c["view"], H, W, max_tiles=257 * 40)
    assert torch.equal(color2, full.color)
