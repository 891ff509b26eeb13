"""Developer probe (GPU): where the time of a SMALL scene goes (BASELINE.json configs[0]: 50 k Gaussians, 512^2,
fwd+bwd).  Prints JSON: device time of one isolated call, of a back-to-back train of calls, the host time to enqueue
one call, and the library's per-stage event intervals.  Run the same script under
`ncu --metrics gpu__time_duration.sum` for the pure kernel durations (tools/gpu_launches.sh)."""
import ctypes
import json
import statistics
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import eogs2_b200 as E                      # noqa: E402
from eogs2_b200 import _cabi                # noqa: E402
from eogs2_b200 import scene as S           # noqa: E402

STAGES = ["", "preprocess", "depth_sort", "bin_rows", "bin_count", "bin_scatter", "blend_fwd", "bwd_zero", "blend_bwd",
          "preprocess_bwd"]


def main():
    P, W, H = (int(a) for a in (sys.argv[1:4] + ["50000", "512", "512"][len(sys.argv[1:4]):]))
    kind = sys.argv[4] if len(sys.argv) > 4 else "trained"
    dev = torch.device("cuda:0")
    sc = S.make_scene(P, kind, 1337)
    d = {k: getattr(sc, k).to(dev) for k in ("means3D", "scales", "rotations", "opacities")}
    view = S.make_camera(1337)
    colors = S.colors_precomp(sc, view).to(dev)
    view = view.to(dev)
    bg = S.background(1337).to(dev)
    dcol, dinv = (t.to(dev) for t in S.upstream_grads(5, H, W, 1337, False))
    empty = torch.empty(0, device=dev)

    def call():
        st = E.rasterize_forward_raw(bg, d["means3D"], colors, d["opacities"], d["scales"], d["rotations"], 1.0, empty,
                                     view, H, W, False, False)
        E.rasterize_backward_raw(st, bg, d["means3D"], colors, d["opacities"], d["scales"], d["rotations"], 1.0, empty,
                                 view, view, dcol, dinv)
        return st

    for _ in range(5):
        st = call()
    torch.cuda.synchronize()
    iso, host = [], []
    for _ in range(20):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.record(); call(); e.record()
        host.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
        iso.append(s.elapsed_time(e))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    N = 50
    s.record()
    for _ in range(N):
        call()
    e.record()
    torch.cuda.synchronize()
    train = s.elapsed_time(e) / N
    lib = _cabi.load()
    lib.eogs_profile_enable(1)
    buf = (ctypes.c_float * 16)()
    acc = [0.0] * len(STAGES)
    for _ in range(10):
        call()
        torch.cuda.synchronize()
        lib.eogs_profile_read(buf, 16)
        for i in range(len(STAGES)):
            acc[i] += buf[i] / 10
    lib.eogs_profile_enable(0)
    print(json.dumps(dict(P=P, W=W, H=H, kind=kind, instances=st.num_rendered, isolated_ms=round(statistics.median(iso), 4),
                          host_enqueue_ms=round(statistics.median(host), 4), back_to_back_ms=round(train, 4),
                          stage_interval_ms={STAGES[i]: round(acc[i], 4) for i in range(1, len(STAGES))},
                          stage_sum_ms=round(sum(acc), 4))))


if __name__ == "__main__":
    main()
