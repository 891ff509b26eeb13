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

# ---- device timeline of the last profiled step: busy time, idle gaps (and what sits on either side of each)
import json, os
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/e2e_trace.json")
ev = [e for e in json.load(open("gpurun_out/e2e_trace.json"))["traceEvents"]
      if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
# steps are separated by the D2H result copy (68 bytes): split on Memcpy DtoH
cuts = [i for i, e in enumerate(ev) if "DtoH" in e["name"] and e.get("args", {}).get("bytes", 0) == 68]
if len(cuts) >= 2:
    a, b = cuts[-2] + 1, cuts[-1] + 1
    step_ev = ev[a:b]
    t0, t1 = ev[cuts[-2]]["ts"] + ev[cuts[-2]]["dur"], step_ev[-1]["ts"] + step_ev[-1]["dur"]
    # H2D copies run on the side stream: leave them out of the critical path
    main = [e for e in step_ev if "HtoD" not in e["name"]]
    busy = sum(e["dur"] for e in main)
    print(f"last step: span {t1 - t0:.1f} us, main-stream busy {busy:.1f} us, {len(main)} device ops")
    prev_end, prev_name = t0, "<previous step's D2H>"
    for e in main:
        gap = e["ts"] - prev_end
        if gap > 4:
            print(f"  idle {gap:7.1f} us  after {prev_name[:50]:50s} before {e['name'][:50]}")
        prev_end, prev_name = max(prev_end, e["ts"] + e["dur"]), e["name"]
    agg = {}
    for e in main:
        agg[e["name"][:70]] = agg.get(e["name"][:70], 0) + e["dur"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:22]:
        print(f"  {v:8.1f} us  {k}")
