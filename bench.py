#!/usr/bin/env python
"""bench.py — fwd+bwd affine renders/sec of the EOGS++ rasterizer path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json `metric`): one "step" = one forward + backward rasterisation of a
synthetic 1 M-Gaussian scene ("trained-like") through a 2048x2048 affine camera,
5 channels (RGB, altitude, opacity) + inverse depth, dense upstream gradient, gradients to all
Gaussian parameters and to the affine camera matrix.

  value     renders/s over all ranks, inputs resident in HBM, device-timed (CUDA events per step,
            max over ranks), L2 flushed between steps (outside the timed spans).
  e2e       the same through the public API (diff_gaussian_rasterization.GaussianRasterizer +
            autograd) with HOST inputs: every step copies the Gaussian tensors (64 MB) from pinned host
            memory (double-buffered on a side stream: step s+1's copy overlaps step s's kernels, one
            copy per step inside the timed region, both arms), and reads the loss and the camera
            gradient back.
  roofline  dominant kernel (blend backward), algorithmic bytes / measured duration vs measured HBM peak.
  cpu_baseline  oracle/cpu_splat.py (PyTorch on the host cores) on a bounded tile sample.

--impl reference times the reference rasterizer's own CUDA kernels (oracle/_ref/libeogs_ref.so,
compiled for sm_100a from /root/reference by oracle/ref_build/Makefile) on the same workload;
when that library is absent it falls back to the CPU port.  N > 1: data parallel over views —
each rank renders its own camera and the 16*P-float gradient bucket (written in place by the backward kernels) is all-reduced with NCCL
inside the step (weak scaling).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

P_GAUSS = 1_000_000
IMG = 2048
SEED = 1337
STAGES = ["", "preprocess", "depth_sort", "bin_rows", "bin_count", "bin_scatter", "blend_fwd",
          "bwd_zero", "blend_bwd", "preprocess_bwd"]


# ----------------------------------------------------------------------------------------------
def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region, every 5 ms through NVML on a host thread (the
    timed region of a default run lasts ~60 ms: `nvidia-smi -lms` would not deliver a single sample in it, which is
    why the lines of round 1 carried `samples: 0` under torchrun).  One sampler per rank, on that rank's GPU."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, gpu_index: int):
        self.gpu, self.sm, self.reasons, self.mx, self.power = gpu_index, [], set(), None, []
        self._stop = threading.Event()
        self.t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:                                   # noqa: BLE001
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for name, bit in self.REASONS:
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:                               # noqa: BLE001
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.nv is not None:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def stop(self) -> dict:
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self._stop.set()
        if self.t is not None:
            self.t.join(timeout=1.0)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None, "source": "NVML, 5 ms period, during the timed regions"}


def make_workload(dev, rank: int):
    from eogs2_b200 import scene as S
    sc = S.make_scene(P_GAUSS, "trained", SEED)
    view = S.make_camera(SEED + rank)                 # one camera per rank (data parallel over views)
    host = dict(means3D=sc.means3D, scales=sc.scales, rotations=sc.rotations, opacities=sc.opacities,
                colors=S.colors_precomp(sc, view))
    host = {k: v.pin_memory() for k, v in host.items()}
    devt = {k: v.to(dev) for k, v in host.items()}
    dcol, dinv = S.upstream_grads(5, IMG, IMG, SEED + rank, False)
    return dict(host=host, dev=devt, view=view.to(dev), view_host=view, bg=S.background(SEED).to(dev),
                dcol=dcol.to(dev), dinv=dinv.to(dev), sc=sc)


class L2Flusher:
    def __init__(self, dev):
        self.buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def __call__(self):
        self.buf.fill_(1)


# ----------------------------------------------------------------------------------------------
# step functions
def ours_step_factory(wl, dev):
    import eogs2_b200 as E
    d = wl["dev"]
    empty = torch.empty(0, device=dev)

    def step():
        st = E.rasterize_forward_raw(wl["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"],
                                     d["rotations"], 1.0, empty, wl["view"], IMG, IMG, False, False)
        g = E.rasterize_backward_raw(st, wl["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"],
                                     d["rotations"], 1.0, empty, wl["view"], wl["view"], wl["dcol"], wl["dinv"])
        return st, g
    return step


def iteration_factory(wl, dev, rank, impl):
    """BASELINE metric, second half ("ms/iter"): the EOGS++ steady-state iteration pattern of ONE camera
    (train_pan.py:278,305-316,375-391) = main W x H + sun 2W x 2H + random-camera W x H
    renders, forward + backward each, same 1 M Gaussians; losses / optimiser excluded."""
    from eogs2_b200 import scene as S
    sc, d = wl["sc"], wl["dev"]
    empty = torch.empty(0, device=dev)
    campos = torch.zeros(3, device=dev)
    cam = wl["view_host"]
    specs = [(cam, IMG, IMG), (S.sun_camera(cam), 2 * IMG, 2 * IMG), (S.random_camera(cam, 0.01, SEED + rank), IMG, IMG)]
    views = []
    for k, (v, Wv, Hv) in enumerate(specs):
        dcol, dinv = S.upstream_grads(5, Hv, Wv, SEED + 10 * rank + k, False)
        views.append(dict(view=v.to(dev), colors=S.colors_precomp(sc, v).to(dev), W=Wv, H=Hv, dcol=dcol.to(dev),
                          dinv=dinv.to(dev)))
    if impl == "ours":
        import eogs2_b200 as E

        def it():
            for v in views:
                st = E.rasterize_forward_raw(wl["bg"], d["means3D"], v["colors"], d["opacities"], d["scales"],
                                             d["rotations"], 1.0, empty, v["view"], v["H"], v["W"], False, False)
                E.rasterize_backward_raw(st, wl["bg"], d["means3D"], v["colors"], d["opacities"], d["scales"],
                                         d["rotations"], 1.0, empty, v["view"], v["view"], v["dcol"], v["dinv"])
    else:
        from oracle import ref_rasterizer as R

        def it():
            for v in views:
                st = R.forward(wl["bg"], d["means3D"], v["colors"], d["opacities"], d["scales"], d["rotations"], 1.0,
                               empty, v["view"], v["view"], 1.0, 1.0, v["H"], v["W"], campos, False, False)
                R.backward(st, wl["bg"], d["means3D"], v["colors"], d["opacities"], d["scales"], d["rotations"], 1.0,
                           empty, v["view"], v["view"], 1.0, 1.0, v["dcol"], v["dinv"], campos, False)
    return it


def ours_e2e_factory(wl, dev, world=1):
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    h = wl["host"]
    names = ["means3D", "scales", "rotations", "opacities", "colors"]
    h2d = sum(h[k].numel() * 4 for k in names)
    out_host = torch.zeros(17, dtype=torch.float32).pin_memory()
    from eogs2_b200.dp import GradBucket, HostInputPipeline
    hsub = {k: h[k] for k in names}
    pipe = HostInputPipeline(dev)
    pipe.submit(hsub)                       # inputs of the first step; every step submits the next one's
    bucket = None
    if world > 1:
        # data parallel over views through the package's own exchange path: ONE flat bucket (16-byte aligned segments),
        # gradients packed into it, all-reduced by the NVLS kernel or ncclAllReduce (whichever calibrates faster)
        shapes = {k: torch.empty_like(h[k], device=dev) for k in names}
        bucket = GradBucket(shapes, {"viewmatrix": torch.empty(4, 4, device=dev)}, exchange="auto")

    def step():
        bufs = pipe.get()                   # this step's inputs (copied from pinned host memory on the side stream)
        t = {k: bufs[k].detach().requires_grad_(True) for k in names}
        view = wl["view"].clone().requires_grad_(True)
        settings = GaussianRasterizationSettings(
            image_height=IMG, image_width=IMG, tanfovx=1.0, tanfovy=1.0, bg=wl["bg"], scale_modifier=1.0,
            viewmatrix=view, projmatrix=view, sh_degree=0, campos=torch.zeros(3, device=dev), prefiltered=False,
            debug=False, antialiasing=False)
        means2D = torch.zeros_like(t["means3D"], requires_grad=True)
        color, radii, invd = GaussianRasterizer(settings)(
            means3D=t["means3D"], means2D=means2D, opacities=t["opacities"], colors_precomp=t["colors"],
            scales=t["scales"], rotations=t["rotations"])
        pipe.submit(hsub)                   # next step's H2D (one copy per step) overlaps this step's blend kernels
        # the upstream gradient dL/dcolor = dcol is handed to autograd directly (what `(color * dcol).sum().backward()`
        # produces, without autograd's broadcast-multiply kernel — the reference arm calls its backward with dcol too);
        # the loss itself is still computed for the read-back
        with torch.no_grad():
            loss = (color * wl["dcol"]).sum()
        torch.autograd.backward(color, wl["dcol"])
        if bucket is not None:
            # the step is not finished before the gradients are summed over ranks
            bucket.params, bucket.extras = t, {"viewmatrix": view}
            bucket.pack()
            flat = bucket.all_reduce()
            a, b = bucket.slices["extra:viewmatrix"]
            view_grad = flat[a:b]
        else:
            view_grad = view.grad.reshape(-1)
        res = torch.cat([loss.detach().reshape(1), view_grad])
        out_host.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(out_host[0])
    step.exchange_name = bucket.exchange_name if bucket is not None else None
    return step, h2d, out_host.numel() * 4


def ref_step_factory(wl, dev):
    from oracle import ref_rasterizer as R
    d = wl["dev"]
    empty = torch.empty(0, device=dev)
    campos = torch.zeros(3, device=dev)

    def step(dd=None):
        dd = dd or d
        st = R.forward(wl["bg"], dd["means3D"], dd["colors"], dd["opacities"], dd["scales"], dd["rotations"], 1.0,
                       empty, wl["view"], wl["view"], 1.0, 1.0, IMG, IMG, campos, False, False)
        g = R.backward(st, wl["bg"], dd["means3D"], dd["colors"], dd["opacities"], dd["scales"], dd["rotations"],
                       1.0, empty, wl["view"], wl["view"], 1.0, 1.0, wl["dcol"], wl["dinv"], campos, False)
        return st, g
    return step


def ref_e2e_factory(wl, dev):
    from oracle import ref_rasterizer as R
    h = wl["host"]
    names = ["means3D", "scales", "rotations", "opacities", "colors"]
    h2d = sum(h[k].numel() * 4 for k in names)
    out_host = torch.zeros(17, dtype=torch.float32).pin_memory()
    inner = ref_step_factory(wl, dev)
    from eogs2_b200.dp import HostInputPipeline     # same double-buffered staging as our arm (plain torch plumbing)
    hsub = {k: h[k] for k in names}
    pipe = HostInputPipeline(dev)
    pipe.submit(hsub)

    def step():
        t = pipe.get()
        st, g = inner(t)
        pipe.submit(hsub)
        # the Python half of the reference's backward (DGR __init__.py:172-202) and a loss read-back
        terms = R.grad_viewmatrix_terms(g, t["means3D"], wl["view"], IMG, IMG)
        loss = (st.color * wl["dcol"]).sum()
        out_host.copy_(torch.cat([loss.reshape(1), terms["total"].reshape(-1)]), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(out_host[0])
    return step, h2d, out_host.numel() * 4


# ----------------------------------------------------------------------------------------------
def timed_steps(step, steps, warmup, flush, world, post=None):
    """K steps, each bracketed by CUDA events on the current stream; L2 flushed between steps."""
    import torch.distributed as dist
    for _ in range(warmup):
        step()
        if post:
            post()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = []
    t0 = time.perf_counter()
    for _ in range(steps):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step()
        if post:
            post()
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    total_ms = sum(s.elapsed_time(e) for s, e in evs)
    timed_steps.local_ms = total_ms                      # this rank's own device time (per-rank report at N > 1)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    return total_ms, wall


def gather_to_rank0(obj, world):
    """Small Python objects of every rank, as a list on rank 0 (None elsewhere)."""
    if world == 1:
        return [obj]
    import torch.distributed as dist
    out = [None] * world if dist.get_rank() == 0 else None
    dist.gather_object(obj, out, dst=0)
    return out


def cpu_baseline(wl, seconds_hint=20.0):
    """PyTorch-on-CPU fwd+bwd (oracle/cpu_splat.py) on a bounded sample of the same workload: geometry
    and binning for all P Gaussians, blend + autograd for a central window of tiles; extrapolated to the
    full image by tile count."""
    from oracle import cpu_splat as CS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sc = wl["sc"]
    gx = IMG // 16
    tens = (sc.means3D, sc.scales, sc.rotations, sc.opacities, wl["host"]["colors"])
    bg_c, dcol_c = wl["bg"].cpu(), wl["dcol"].cpu()

    def window(win):
        tw = (gx // 2 - win // 2, gx // 2 - win // 2, gx // 2 + win // 2, gx // 2 + win // 2) if win else (0, 0, 0, 0)
        t0 = time.perf_counter()
        CS.render_fwd_bwd(tens, wl["view_host"], bg_c, IMG, IMG, dcol_c, None, False, tile_window=tw)
        return time.perf_counter() - t0
    t_geom = window(0)                                 # per-Gaussian work + sort + autograd of it, no tiles
    win = 6                                            # probe: 6x6 = 36 of 16384 tiles
    t_win = window(win)
    per_tile = max(t_win - t_geom, 1e-6) / (win * win)
    # the sample proper: a central window sized for about `seconds_hint` of CPU work (bounded: 64x64 tiles)
    big = min(64, int((0.6 * seconds_hint / per_tile) ** 0.5) // 2 * 2)
    if big > win + 2:
        win = big
        t_win = window(win)
        per_tile = max(t_win - t_geom, 1e-6) / (win * win)
    est = t_geom + per_tile * gx * gx
    return {"value": 1.0 / est, "unit": "renders/s", "cores": cores, "kind": "port",
            "sample": f"oracle/cpu_splat.py (PyTorch CPU, fp32, autograd): all 1M Gaussians preprocessed+sorted "
                      f"({t_geom:.1f} s) + {win*win} of {gx*gx} tiles blended fwd+bwd ({t_win - t_geom:.1f} s), "
                      f"extrapolated by tile count to {est:.0f} s per render"}


# ----------------------------------------------------------------------------------------------
def count_pairs():
    """Evaluated / blended (pixel, Gaussian) pairs of ONE fwd+bwd render of the bench workload (SURVEY.md section 8d:
    "always report pairs_eval, pairs_blend"), counted by the instrumented twin of the library
    (libeogs_raster_count.so, -DEOGS_COUNT_PAIRS=1) in a child process — never inside a timed region, never with
    the product library."""
    from eogs2_b200.build import COUNT_LIB
    if not COUNT_LIB.exists():
        return {"unavailable": f"{COUNT_LIB.name} not built"}
    env = dict(os.environ, EOGS_RASTER_LIB=str(COUNT_LIB))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--count-pairs-child"], env=env, capture_output=True,
                           text=True, timeout=300)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:                                  # noqa: BLE001
        return {"unavailable": f"counting run failed: {e}"}


def count_pairs_child():
    import eogs2_b200 as E
    from eogs2_b200 import _cabi
    lib = _cabi.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    wl = make_workload(dev, 0)
    buf = (ctypes.c_ulonglong * 16)()
    if lib.eogs_debug_counters(buf, 1) != 1:
        print(json.dumps({"unavailable": "library built without EOGS_COUNT_PAIRS"}))
        return 0
    st, _ = ours_step_factory(wl, dev)()
    torch.cuda.synchronize()
    lib.eogs_debug_counters(buf, 1)
    names = ["fwd_pairs_eval", "fwd_pairs_blend", "fwd_lane_slots", "fwd_entries", "bwd_pairs_eval", "bwd_pairs_blend",
             "bwd_lane_slots", "bwd_entries", "bwd_flushes"]
    print(json.dumps(dict({n: int(buf[i]) for i, n in enumerate(names)}, instances=st.num_rendered)))
    return 0


def ncu_record():
    """The committed ncu capture of the CURRENT kernels, if any: profiles/ncu_current.json is written by
    tools/ncu_to_json.py from a `ncu --set full` report of this same command, together with the hash of the CUDA
    sources it was taken from; a capture of other sources is not quoted."""
    from eogs2_b200.build import source_hash
    f = ROOT / "profiles" / "ncu_current.json"
    if not f.exists():
        return None, "no profiles/ncu_current.json"
    rec = json.loads(f.read_text())
    if rec.get("source_hash") != source_hash():
        return None, "profiles/ncu_current.json was captured from different kernel sources (hash mismatch): not quoted"
    return rec, f"profiles/ncu_current.json ({rec.get('report', '?')})"


def roofline_report(stage_ms, I, peaks, hbm_peak, peak_src, clocks):
    """Dominant kernel = blend backward.  The contract's HBM figure (algorithmic bytes / CUDA-event time / measured
    copy bandwidth) plus what actually bounds the kernel: instruction issue and FP32 rate (SURVEY.md section 8d:
    "report both; the binding one is the roofline fraction")."""
    bwd_ms = stage_ms[STAGES.index("blend_bwd")]
    fwd_ms = stage_ms[STAGES.index("blend_fwd")]
    # Algorithmic bytes per launch: per instance 4 B id + 48 B record + 4 B alpha_cut gathered + 44 B
    # (11 floats) reduced into the gradient record; per pixel 4*(C+1) B upstream gradient + 8 B final_T / n_contrib.
    bwd_bytes = I * (4 + 48 + 4 + 44) + IMG * IMG * (4 * 6 + 8)
    achieved = bwd_bytes / (bwd_ms * 1e-3) / 1e9 if bwd_ms > 0 else 0.0
    pairs = count_pairs()
    rec, rec_src = ncu_record()
    sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
    n_sm = 148
    fp32_peak = n_sm * 128 * 2 * sm_mhz * 1e6 / 1e12        # TFLOP/s at the SM clock sampled during the run
    out = {"kernel": "blend_bwd_kernel<5>", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
           "frac": achieved / hbm_peak, "peak_source": peak_src, "ms_per_launch": bwd_ms, "algorithmic_bytes": bwd_bytes,
           "traffic": None, "traffic_source": rec_src, "pairs": pairs,
           "binding": "instruction issue (see issue_frac): DRAM traffic is a fraction of the algorithmic bytes because the "
                      "48 MB record array is gathered out of L2 and entries behind a tile's last contributor are never "
                      "fetched; the HBM fraction is reported because the contract asks for it"}
    if "bwd_pairs_blend" in pairs:
        # FLOP model of SURVEY.md section 8d: forward 11 FLOP per evaluated pair + 16 per blended pair (C = 5 + inverse depth),
        # backward 2.2 x the forward's
        f_flop = pairs["fwd_pairs_eval"] * 11 + pairs["fwd_pairs_blend"] * 16
        b_flop = 2.2 * (pairs["bwd_pairs_eval"] * 11 + pairs["bwd_pairs_blend"] * 16)
        out["fp32_peak_tflops"] = fp32_peak
        out["fp32_frac"] = b_flop / (bwd_ms * 1e-3) / 1e12 / fp32_peak if bwd_ms > 0 else None
        out["fwd"] = {"kernel": "blend_fwd_kernel<5>", "ms_per_launch": fwd_ms,
                      "fp32_frac": f_flop / (fwd_ms * 1e-3) / 1e12 / fp32_peak if fwd_ms > 0 else None}
    if rec is not None:
        k = rec["kernels"].get("blend_bwd_kernel", {})
        out["traffic"] = k.get("dram_bytes")
        issue_slots = n_sm * 4 * sm_mhz * 1e6 * bwd_ms * 1e-3
        if k.get("inst_executed"):
            out["issue_frac"] = k["inst_executed"] / issue_slots          # warp instructions / issue slots of THIS run's launch time
            out["issue_active_pct_ncu"] = k.get("issue_active_pct")
            out["warps_active_per_sm_ncu"] = k.get("warps_active_per_sm")
            if pairs.get("bwd_pairs_blend"):
                out["thread_inst_per_blended_pair"] = 32.0 * k["inst_executed"] / pairs["bwd_pairs_blend"]
        kf = rec["kernels"].get("blend_fwd_kernel", {})
        if kf.get("inst_executed") and "fwd" in out:
            out["fwd"]["issue_frac"] = kf["inst_executed"] / (n_sm * 4 * sm_mhz * 1e6 * fwd_ms * 1e-3)
            out["fwd"]["traffic"] = kf.get("dram_bytes")
            if pairs.get("fwd_pairs_blend"):
                out["fwd"]["thread_inst_per_blended_pair"] = 32.0 * kf["inst_executed"] / pairs["fwd_pairs_blend"]
    return out


def config5_leg(dev, rank, world, flush, reps=3):
    """BASELINE configs[4]: 5 M Gaussians, one 8192 x 8192 nadir altitude/DSM render (forward), tile rows sharded
    over the N GPUs (eogs2_b200/bands.py: instance-balanced bands, one all-gather of the image); at N = 1 the image is
    rendered as consecutive bands on the one GPU.  Device-timed, max over ranks, all-gather included."""
    import eogs2_b200 as E
    from eogs2_b200 import bands as B
    from eogs2_b200 import scene as S
    import torch.distributed as dist
    P5, IMG5 = 5_000_000, 8192
    try:
        sc = S.make_scene(P5, "trained", SEED)
        A = torch.tensor([[1 / 0.72, 0, 0], [0, 1 / 0.72, 0], [0, 0, S.METRES_PER_UNIT]])      # nadir camera
        view = S.affine_to_viewmatrix(A, torch.zeros(3))
        colors = S.colors_precomp(sc, view).to(dev)
        view = view.to(dev)
        t = {k: getattr(sc, k).to(dev) for k in ("means3D", "scales", "rotations", "opacities")}
        bg = S.background(SEED).to(dev)
        empty = torch.empty(0, device=dev)
        grid_y = (IMG5 + 15) // 16
        lib = E._cabi.load()

        def render(weights=None):
            if world == 1:
                return B.forward_strips(bg, t["means3D"], colors, t["opacities"], t["scales"], t["rotations"], 1.0, empty,
                                        view, IMG5, IMG5, max_tiles=65536)
            return B.forward_band(bg, t["means3D"], colors, t["opacities"], t["scales"], t["rotations"], 1.0, empty, view,
                                  IMG5, IMG5, rank, world, weights=weights)
        weights, inst = None, None
        out = render()
        if world > 1:
            st = out[2]
            r = E.export_state(st)["ranges"].to(torch.int64)
            per_row = (r[:, 1] - r[:, 0]).view(-1, (IMG5 + 15) // 16).sum(1)
            all_rows = torch.zeros(grid_y, dtype=torch.int64, device=dev)
            all_rows[st.rows[0]:st.rows[1]] = per_row
            dist.all_reduce(all_rows)
            weights, inst = all_rows.cpu().tolist(), int(all_rows.sum().item())
        del out
        render(weights)
        ts, stages = [], [0.0] * len(STAGES)
        buf = (ctypes.c_float * 16)()
        for _ in range(reps):
            flush()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); out = render(weights); e.record()
            torch.cuda.synchronize()
            ms = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            ts.append(float(ms.item()))
            checksum = float(out[0][3].double().abs().sum().item())
            del out
        if world > 1:                                          # this rank's band, stage by stage (rank 0 reports its own)
            lib.eogs_profile_enable(1)
            flush(); render(weights)
            lib.eogs_profile_read(buf, 16)
            lib.eogs_profile_enable(0)
            stages = [buf[i] for i in range(len(STAGES))]
        res = {"ms": statistics.median(ts), "P": P5, "W": IMG5, "H": IMG5, "passes": "forward, 5 channels + inverse depth",
               "sharding": f"{world} instance-balanced tile bands + one all-gather" if world > 1 else
                           "one GPU: consecutive bands of <= 65 536 tiles",
               "instances": inst, "altitude_checksum": checksum, "max_over_ranks": world > 1}
        if world > 1:
            res["rank0_stage_ms"] = {STAGES[i]: round(stages[i], 3) for i in range(1, 7)}
        return res
    except Exception as ex:                                    # noqa: BLE001 — the headline line must survive this leg
        return {"ms": None, "error": str(ex)[:300]}
    finally:
        torch.cuda.empty_cache()


_REAL_STDOUT = None


def emit(text: str) -> None:
    """The result line goes to the process's original stdout (see the fd juggling around the NCCL init)."""
    if _REAL_STDOUT is None:
        print(text, flush=True)
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, (text + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the e2e leg (the line then carries e2e=null)")
    ap.add_argument("--no-config5", action="store_true", help="skip the 5 M / 8192^2 tile-band leg (config5_ms)")
    ap.add_argument("--count-pairs-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.count_pairs_child:
        return count_pairs_child()
    args.warmup = max(args.warmup, 3)
    rank, world, local = dist_env()

    if args.impl == "reference" and rank != 0:
        return 0                                           # rank 0 alone runs the reference arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the rasterizer has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ref_world = world
    if args.impl == "reference":
        world = 1
    if world > 1:
        import torch.distributed as dist
        # stdout carries ONE JSON line, but NCCL writes its debug output — including the "NCCL version ..." banner of
        # the VERSION / WARN levels this image exports — straight to file descriptor 1.  Point fd 1 at stderr for the
        # rest of the run and keep the real stdout for the result line (emit()).
        global _REAL_STDOUT
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    wl = make_workload(dev, rank)
    flush = L2Flusher(dev)
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    base = {"metric": "fwd+bwd affine renders/sec at 1M Gaussians 2048^2", "unit": "renders/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[2] main view: 1M trained-like Gaussians (seed 1337), one 2048x2048 "
                                   "affine camera per rank, 5 channels + inverse depth, fwd+bwd with dense upstream "
                                   "gradient, gradients to all Gaussian parameters and the camera matrix",
                       "P": P_GAUSS, "W": IMG, "H": IMG, "channels": 5,
                       "l2": "256 MiB buffer written between timed steps (outside the event spans)",
                       "e2e_inputs": "64 MB of pinned host tensors copied every step, double-buffered on a side stream",
                       "parallelism": f"dp{args.gpus} over views" if args.gpus > 1 else "single GPU"}}

    sampler = ClockSampler(local)

    if args.impl == "reference":
        from oracle import ref_rasterizer as R
        if not R.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libeogs_ref.so not built (needs /root/reference at build time)"}))
            return 0
        step = ref_step_factory(wl, dev)
        sampler.start()
        total_ms, wall = timed_steps(step, args.steps, args.warmup, flush, 1)
        e2e_step, h2d, d2h = ref_e2e_factory(wl, dev)
        e2e_ms, _ = timed_steps(e2e_step, args.steps, args.warmup, flush, 1)
        it_steps = max(3, min(args.steps, 5))
        iter_ms, _ = timed_steps(iteration_factory(wl, dev, 0, "reference"), it_steps, 3, flush, 1)
        clocks = sampler.stop()
        st, _ = step()
        val = args.steps / (total_ms / 1e3)
        line = dict(base, impl="reference", n_gpus=1, value=val, ms_per_step=total_ms / args.steps, clocks=clocks,
                    e2e={"value": args.steps / (e2e_ms / 1e3), "unit": "renders/s", "h2d_bytes_per_step": h2d,
                         "d2h_bytes_per_step": d2h},
                    gpu_launches=0,
                    cpu_baseline={"value": val, "unit": "renders/s", "cores": 0, "kind": "reference",
                                  "sample": "full workload; the reference has no CPU backend, so this arm runs its own CUDA "
                                            "kernels (DGR cuda_rasterizer, compiled for sm_100a with nvcc defaults) on the GPU"},
                    iter_ms=iter_ms / it_steps,
                    instances=st.num_rendered, requested_gpus=ref_world)
        emit(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ ours
    import eogs2_b200 as E
    from eogs2_b200 import _cabi
    lib = _cabi.load()
    step = ours_step_factory(wl, dev)
    post = None
    exchange = exchange_name = None
    d = wl["dev"]
    empty = torch.empty(0, device=dev)
    if world > 1:
        import torch.distributed as dist
        # ONE flat fp32 bucket holds every gradient the rasterizer returns for the replicated parameters —
        # means3D 3, colours 5 (rgb -> f_dc, altitude -> xyz through the altitude colour, 1), opacity 1,
        # scales 3, rotations 4 = 16 floats per Gaussian — plus the 16 camera sums.  The backward kernels
        # write straight into views of it (no packing copies); it is all-reduced over NVLink inside the step.
        P = P_GAUSS
        from eogs2_b200.nvls import make_grad_exchange
        # the exchange is the library's own NVLS kernel (multimem.ld_reduce + multimem.st on a symmetric-memory bucket),
        # its peer-to-peer kernel (2-4 GPUs) or ncclAllReduce, whichever a short calibration finds faster on this box; named in config["allreduce"]
        bucket, exchange, exchange_name = make_grad_exchange(16 * P + 16, dev)
        views = {"means3D": bucket[0:3 * P].view(P, 3), "colors": bucket[3 * P:8 * P].view(P, 5),
                 "opacity": bucket[8 * P:9 * P].view(P, 1), "scales": bucket[9 * P:12 * P].view(P, 3),
                 "rotations": bucket[12 * P:16 * P].view(P, 4), "cam_sums": bucket[16 * P:]}

        def make_step_dp(view, dcol, dinv, colors):
            def step_dp():
                st = E.rasterize_forward_raw(wl["bg"], d["means3D"], colors, d["opacities"], d["scales"],
                                             d["rotations"], 1.0, empty, view, IMG, IMG, False, False)
                g = E.rasterize_backward_raw(st, wl["bg"], d["means3D"], colors, d["opacities"], d["scales"],
                                             d["rotations"], 1.0, empty, view, view, dcol, dinv, out=views)
                return st, g
            return step_dp

        def post():
            exchange()
        run_step = make_step_dp(wl["view"], wl["dcol"], wl["dinv"], d["colors"])
        base["config"]["allreduce"] = exchange_name
    else:
        run_step = step

    sampler.start()
    total_ms, wall = timed_steps(run_step, args.steps, args.warmup, flush, world, post)
    local_ms = timed_steps.local_ms

    # per-stage device times (CUDA events recorded inside the library on the launch stream)
    lib.eogs_profile_enable(1)
    stage_ms = [0.0] * len(STAGES)
    reps = max(3, min(args.steps, 10))
    buf = (ctypes.c_float * 16)()
    for _ in range(reps):
        flush()
        step()
        lib.eogs_profile_read(buf, 16)
        for i in range(len(STAGES)):
            stage_ms[i] += buf[i] / reps
    lib.eogs_profile_enable(0)
    st, g = step()
    torch.cuda.synchronize()
    I = st.num_rendered

    # e2e through the public API with host inputs
    e2e_step, h2d, d2h = ours_e2e_factory(wl, dev, world)
    e2e_ms = None
    if not args.no_e2e:
        e2e_ms, _ = timed_steps(e2e_step, args.steps, args.warmup, flush, world)

    # "ms/iter": main + sun@2x + random renders fwd+bwd per rank (+ the gradient all-reduce at N > 1), max over ranks
    it_steps = max(3, min(args.steps, 10))
    iter_ms, _ = timed_steps(iteration_factory(wl, dev, rank, "ours"), it_steps, 3, flush, world, post)

    extra = {}
    if world > 1:
        import torch.distributed as dist
        from eogs2_b200 import scene as S
        # (1) pure weak scaling: EVERY rank renders rank 0's camera, so the step differs from the 1-GPU step only by
        #     the exchange and the barrier skew — no workload variance between cameras
        v0 = S.make_camera(SEED)
        dcol0, dinv0 = S.upstream_grads(5, IMG, IMG, SEED, False)
        same = make_step_dp(v0.to(dev), dcol0.to(dev), dinv0.to(dev), S.colors_precomp(wl["sc"], v0).to(dev))
        same_ms, _ = timed_steps(same, args.steps, args.warmup, flush, world, post)
        noex_ms, _ = timed_steps(same, args.steps, args.warmup, flush, world, None)
        # (2) the exchange alone, and a bit check of the own kernel against ncclAllReduce on the same data
        ex_ms, _ = timed_steps(lambda: None, args.steps, args.warmup, lambda: None, world, post)
        check = None
        if exchange_name.startswith("own "):
            gen = torch.Generator(device=dev).manual_seed(1234 + rank)
            pattern = torch.randn(bucket.numel(), device=dev, generator=gen)
            bucket.copy_(pattern); exchange(); mine = bucket.clone()
            ref = pattern.clone(); dist.all_reduce(ref)
            check = "bit-identical to ncclAllReduce" if torch.equal(mine, ref) else \
                f"max |own - nccl| = {float((mine - ref).abs().max()):.3e} (summation order differs)"
        extra["dp"] = {
            "e2e_exchange": getattr(e2e_step, "exchange_name", None),
            "per_rank": gather_to_rank0({"rank": rank, "camera_seed": SEED + rank, "instances": I,
                                         "ms_per_step_local": local_ms / args.steps}, world),
            "same_camera": {"ms_per_step": same_ms / args.steps, "value": world * args.steps / (same_ms / 1e3),
                            "ms_per_step_without_exchange": noex_ms / args.steps,
                            "what": "every rank renders camera 1337: weak scaling without workload variance"},
            "exchange_ms": ex_ms / args.steps, "exchange_bytes": bucket.numel() * 4, "exchange_check": check}
    clocks = sampler.stop()
    all_clocks = gather_to_rank0(clocks, world)
    if rank == 0 and world > 1:
        ok = [c for c in all_clocks if c and c.get("sm_mhz")]
        clocks = dict(clocks, per_rank_sm_mhz=[c.get("sm_mhz") for c in all_clocks],
                      sm_mhz=min(c["sm_mhz"] for c in ok) if ok else None,
                      reasons=sorted(set(r for c in all_clocks if c for r in c.get("reasons", []))),
                      samples=sum(c.get("samples", 0) for c in all_clocks if c))

    # config 5 (BASELINE configs[4]): 5 M Gaussians, 8192^2 forward, tile rows sharded over the N GPUs + all-gather
    c5 = None if args.no_config5 else config5_leg(dev, rank, world, flush)

    value = world * args.steps / (total_ms / 1e3)
    line = dict(base, value=value, ms_per_step=total_ms / args.steps, clocks=clocks,
                e2e=None if e2e_ms is None else {"value": world * args.steps / (e2e_ms / 1e3), "unit": "renders/s",
                     "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                stage_ms={STAGES[i]: round(stage_ms[i], 4) for i in range(1, len(STAGES))},
                iter_ms=iter_ms / it_steps,
                iter_pattern="one camera per rank: main 2048^2 + sun 4096^2 + random 2048^2, fwd+bwd each"
                             + (", then the gradient all-reduce" if world > 1 else ""),
                instances=I, wall_s=wall, **extra)
    if c5 is not None:
        line["config5_ms"] = c5["ms"]
        line["config5"] = c5
    if rank == 0:
        # preprocess, 4 binning passes (count / scan / offsets / scatter) x rows and columns, blend fwd, tile order,
        # blend bwd, preprocess bwd (+ the own all-reduce kernel when it is the chosen exchange); the CUB depth sort
        # and torch's fill / barrier kernels are not counted
        own = 1 + 8 + 2 + 2 + (1 if world > 1 and exchange_name.startswith("own ") else 0)
        line["gpu_launches"] = own * args.steps
        line["roofline"] = roofline_report(stage_ms, I, peaks, hbm_peak, peak_src, clocks)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
        emit(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
