"""Developer probe (GPU): where does the end-to-end (host-input) step spend its time?"""
import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench as B

dev = torch.device("cuda:0")
wl = B.make_workload(dev, 0)
step, h2d, d2h = B.ours_e2e_factory(wl, dev)
for _ in range(5):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("e2e wall ms:", [round(t * 1e3, 2) for t in ts])

# H2D alone
h = wl["host"]
names = ["means3D", "scales", "rotations", "opacities", "colors"]
for _ in range(3):
    t0 = time.perf_counter()
    t = {k: h[k].to(dev, non_blocking=True) for k in names}
    torch.cuda.synchronize()
    print("h2d ms", round((time.perf_counter() - t0) * 1e3, 3), "GB/s", round(h2d / (time.perf_counter() - t0) / 1e9, 1))
one = torch.empty(64_000_000, dtype=torch.uint8).pin_memory()
for _ in range(3):
    t0 = time.perf_counter(); one.to(dev, non_blocking=True); torch.cuda.synchronize()
    print("h2d single 64MB ms", round((time.perf_counter() - t0) * 1e3, 3))

raw = B.ours_step_factory(wl, dev)
for _ in range(3): raw()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); raw(); torch.cuda.synchronize(); print("raw fwd+bwd wall ms", round((time.perf_counter() - t0) * 1e3, 3))

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=25, max_name_column_width=60))
