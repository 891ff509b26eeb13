"""CPU property tests (hypothesis) of host-side arithmetic that shards or sizes work: tile-band splits, the DSM
grid of utils/dsm_utils.py:20-25, the view sharding of the data-parallel harness, and the scratch-size queries."""
import math

from hypothesis import given, settings, strategies as st

from eogs2_b200 import bands as B
from eogs2_b200 import dp
from eogs2_b200.dsm import dsm_grid


@settings(max_examples=200, deadline=None)
@given(grid_y=st.integers(1, 600), world=st.integers(1, 16), seed=st.integers(0, 10_000), weighted=st.booleans())
def test_split_rows_is_a_contiguous_cover(grid_y, world, seed, weighted):
    import random
    rnd = random.Random(seed)
    weights = [rnd.choice([0.0, 1.0, 5.0, 1000.0]) for _ in range(grid_y)] if weighted else None
    bands = B.split_rows(grid_y, world, weights)
    assert len(bands) == world and bands[0][0] == 0
    assert all(a[1] == b[0] for a, b in zip(bands, bands[1:])) and max(b[1] for b in bands) == grid_y
    live = [b for b in bands if b[1] > b[0]]
    assert len(live) == min(world, grid_y)                      # every rank gets >= 1 row while rows last
    assert sum(b[1] - b[0] for b in bands) == grid_y


@settings(max_examples=200, deadline=None)
@given(H=st.integers(1, 9000), world=st.integers(1, 8))
def test_band_heights_add_up_to_the_image(H, world):
    grid_y = (H + 15) // 16
    bands = B.split_rows(grid_y, world)
    assert sum(B.band_height(b, H) for b in bands) == H


@settings(max_examples=300, deadline=None)
@given(x0=st.floats(-1e6, 1e6), dx=st.floats(0.0, 5e3), y0=st.floats(-1e7, 1e7), dy=st.floats(0.0, 5e3),
       res=st.sampled_from([0.3, 0.5, 1.0, 2.0]))
def test_dsm_grid_contains_every_point(x0, dx, y0, dy, res):
    xoff, yoff, w, h = dsm_grid(x0, x0 + dx, y0, y0 + dy, res)
    assert w >= 1 and h >= 1
    # plyflatten's cell of the extreme points lies inside the raster (up to one ulp of the division at the far edge)
    for x in (x0, x0 + dx):
        assert -1 <= math.floor((x - xoff) / res) <= w
    for y in (y0, y0 + dy):
        assert -1 <= math.floor((yoff - y) / res) <= h


@given(n=st.integers(0, 64), world=st.integers(1, 16))
def test_shard_views_is_a_partition(n, world):
    parts = [dp.shard_views(n, r, world) for r in range(world)]
    assert sorted(sum(parts, [])) == list(range(n))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


@settings(max_examples=50, deadline=None)
@given(P=st.integers(0, 3_000_000))
def test_scratch_sizes_grow_with_the_problem(P):
    from eogs2_b200 import _cabi
    lib = _cabi.load()
    assert lib.eogs_geom_bytes(P + 1000) >= lib.eogs_geom_bytes(P) >= (P * (48 + 4 * 8) if P else 1)
    assert lib.eogs_knn_bytes(P + 1000) >= lib.eogs_knn_bytes(P) > 0
    assert lib.eogs_prune_temp_bytes(max(P, 1)) > 0
