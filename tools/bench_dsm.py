"""DSM splat timing (eogs_dsm_splat) on a nadir-render-sized cloud, next to the CPU restatement of plyflatten
(oracle_plyflatten, scalar C) on the same points.   python tools/bench_dsm.py > gpurun_out/<tag>/dsm.json"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from eogs2_b200.dsm import compute_dsm, dsm_grid      # noqa: E402
from oracle import c_oracle as O                      # noqa: E402

rows = []
for n in (1024, 2048):
    g = torch.Generator(device="cuda").manual_seed(0)
    u = (torch.arange(n, device="cuda", dtype=torch.float64) + 0.5) * 0.5
    xy = torch.stack(torch.meshgrid(4.3e5 + u, 3.36e6 - u, indexing="xy"), -1).reshape(-1, 2)
    xy = xy + (torch.rand(xy.shape, device="cuda", dtype=torch.float64, generator=g) - 0.5) * 0.4
    z = 20 + 10 * torch.sin(xy[:, :1] / 15.0)
    cloud = torch.cat([xy, z], 1).contiguous()
    for _ in range(3):
        profile, dsm = compute_dsm(cloud, 0.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        profile, dsm = compute_dsm(cloud, 0.5)
    e1.record(); torch.cuda.synchronize()
    ours_ms = e0.elapsed_time(e1) / 10
    t0 = time.perf_counter()
    c = cloud.cpu().numpy()                                        # what the reference does first (dsm_utils.py:8)
    t_copy = time.perf_counter() - t0
    xoff, yoff, w, h = dsm_grid(c[:, 0].min(), c[:, 0].max(), c[:, 1].min(), c[:, 1].max(), 0.5)
    t0 = time.perf_counter()
    ref = O.plyflatten(c, xoff, yoff, 0.5, w, h, 1, float("inf"))
    t_cpu = time.perf_counter() - t0
    mine = dsm.cpu().numpy()
    ok = bool(np.array_equal(np.isnan(mine), np.isnan(ref)) and np.nanmax(np.abs(mine - ref)) < 1e-4)
    rows.append({"points": n * n, "raster": [h, w], "ours_ms_incl_extent_readback": round(ours_ms, 3),
                 "cpu_port_ms": round(t_cpu * 1e3, 1), "d2h_copy_of_cloud_ms": round(t_copy * 1e3, 1), "matches_oracle": ok})
    print(rows[-1], file=sys.stderr)
print(json.dumps(rows, indent=1))
