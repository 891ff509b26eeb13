"""CPU, world_size 2 and 3, gloo: host logic of the tile-band sharding (eogs2_b200/bands.py) —
row partition, uneven band heights, the padded all-gather and the gradient all-reduce — with a
stand-in band renderer (the CUDA kernels need a GPU; tests/test_bands_gpu.py covers them)."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from eogs2_b200 import bands as B  # noqa: E402

W, H, C = 40, 150, 5                    # 10 tile rows, last one ragged (150 = 9*16 + 6)


def full_image():
    g = torch.Generator().manual_seed(5)
    return torch.randn(C + 1, H, W, generator=g)


def fake_render_band(band):
    """What the band kernels return: rows [16*rb, min(H, 16*re)) of the whole-image render."""
    img = full_image()
    y0, y1 = 16 * band[0], min(H, 16 * band[1])
    return img[:C, y0:y1].contiguous(), img[C:, y0:y1].contiguous()


def worker(rank, world, port, out, weights):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    color, invd, band = B.render_sharded(fake_render_band, H, rank, world, weights)
    # gradient exchange: each rank contributes (rank+1) * ones; the sum must arrive everywhere
    flat = torch.full((7,), float(rank + 1))
    dist.all_reduce(flat)
    torch.save({"color": color, "invd": invd, "band": band, "flat": flat}, out + f".{rank}")
    dist.destroy_process_group()


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("world,weights", [(2, None), (3, None), (3, [1, 1, 1, 1, 1, 1, 1, 1, 30, 30])])
def test_gathered_bands_equal_the_whole_image(tmp_path, world, weights):
    out = str(tmp_path / "res")
    mp.spawn(worker, args=(world, free_port(), out, weights), nprocs=world, join=True)
    img = full_image()
    seen = []
    for r in range(world):
        d = torch.load(out + f".{r}")
        assert torch.equal(d["color"], img[:C]) and torch.equal(d["invd"], img[C:])
        assert torch.equal(d["flat"], torch.full((7,), float(sum(range(1, world + 1)))))
        seen.append(d["band"])
    assert seen == B.split_rows((H + 15) // 16, world, weights)


def test_split_rows_properties():
    for grid_y in (1, 7, 10, 512):
        for world in (1, 2, 3, 8):
            bands = B.split_rows(grid_y, world)
            live = [b for b in bands if b[1] > b[0]]
            assert len(bands) == world and live[0][0] == 0 and live[-1][1] == grid_y
            assert all(a[1] == b[0] for a, b in zip(live, live[1:]))
            sizes = [b[1] - b[0] for b in live]
            assert max(sizes) - min(sizes) <= 1
    # weighted: heavy rows are isolated, every band keeps >= 1 row
    b = B.split_rows(8, 4, [1, 1, 1, 1, 5, 5, 1, 1])
    assert b == [(0, 4), (4, 5), (5, 6), (6, 8)]
    b = B.split_rows(10, 4, [0] * 9 + [100])
    assert all(x[1] > x[0] for x in b) and b[-1][1] == 10
    assert B.band_height((9, 10), 150) == 6 and B.band_height((0, 10), 150) == 150
    with pytest.raises(ValueError):
        B.split_rows(4, 2, [1, 2, 3])
