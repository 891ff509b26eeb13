"""All five BASELINE.json configs on the GPU box, ours next to the compiled reference (oracle/_ref).

    python tools/bench_configs.py [--configs 1,2,3,5] [--reps 5] [--out gpurun_out/<tag>/configs.json]
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_configs.py --configs 5     (tile bands over N GPUs)

bench.py stays the one headline line (configs[2] main view); this script produces the table quoted in
DESIGN.md / profiles/: device time per config (CUDA events, median of --reps after 2 warm-ups, L2
flushed between reps), instances, and the same workload through the reference's own kernels.
  1  50 k Gaussians, 512^2, fwd+bwd                                     (configs[0])
  2  300 k Gaussians, 19 views 2048^2, one training iteration = 19 fwd+bwd   (configs[1])
  3  1 M Gaussians: main 2048^2 + sun 4096^2 + random 2048^2, fwd+bwd, camera gradients (configs[2])
  5  5 M Gaussians, 8192^2 forward (altitude/DSM), tile bands: 1 GPU = whole image; N GPUs = N bands
     + all-gather                                                        (configs[4])
  6  the render() glue: one fwd+bwd render of 1 M raw GaussianModel parameters at 2048^2
     through (a) torch activations + colors_precomp + our rasterizer (what an unchanged renderer.py does) and
     (b) eogs2_b200.fused.render_fused (activations and chain rules inside the geometry kernels)
  7  the sun-view resample (renderer_cc_shadow.py:32-46) alone, 4096^2 virtual render -> 2048^2 camera frame,
     fwd+bwd: torch einsum + grid_sample + mask vs eogs2_b200.shadow.resample_virtual
  8  the photometric loss (loss/shadow.py:21-29) on a 3 x 2048^2 image, fwd+bwd: torch l1 + 5 x conv2d SSIM vs
     eogs2_b200.losses.photometric_loss
  9  one full camera iteration at 1 M Gaussians, 2048^2 (main render + sun render at 4096^2 + resample + shading +
     photometric loss + backward + Adam): the reference's torch stages on our rasterizer vs everything fused
     (eogs2_b200/iteration.py)
(config 4 = config 3's cameras data-parallel over ranks is what `bench.py --gpus N` measures.)
"""
import argparse
import json
import os
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import eogs2_b200 as E                      # noqa: E402
from eogs2_b200 import bands as B           # noqa: E402
from eogs2_b200 import scene as S           # noqa: E402
from oracle import ref_rasterizer as R      # noqa: E402


def timed(fn, reps, flush, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return statistics.median(ts)


class Views:
    def __init__(self, dev, P, kind, seed):
        self.dev = dev
        self.sc = S.make_scene(P, kind, seed)
        self.d = {k: getattr(self.sc, k).to(dev) for k in ("means3D", "scales", "rotations", "opacities")}
        self.empty = torch.empty(0, device=dev)
        self.campos = torch.zeros(3, device=dev)
        self.bg = S.background(seed).to(dev)

    def view(self, view, W, H, seed):
        dcol, dinv = S.upstream_grads(5, H, W, seed, False)
        return dict(view=view.to(self.dev), colors=S.colors_precomp(self.sc, view).to(self.dev), W=W, H=H,
                    dcol=dcol.to(self.dev), dinv=dinv.to(self.dev))

    def ours(self, v, backward=True, band=None):
        d = self.d
        st = E.rasterize_forward_raw(self.bg, d["means3D"], v["colors"], d["opacities"], d["scales"], d["rotations"],
                                     1.0, self.empty, v["view"], v["H"], v["W"], False, False, band=band)
        if backward:
            E.rasterize_backward_raw(st, self.bg, d["means3D"], v["colors"], d["opacities"], d["scales"],
                                     d["rotations"], 1.0, self.empty, v["view"], v["view"], v["dcol"], v["dinv"])
        return st

    def ref(self, v, backward=True):
        d = self.d
        st = R.forward(self.bg, d["means3D"], v["colors"], d["opacities"], d["scales"], d["rotations"], 1.0,
                       self.empty, v["view"], v["view"], 1.0, 1.0, v["H"], v["W"], self.campos, False, False)
        if backward:
            R.backward(st, self.bg, d["means3D"], v["colors"], d["opacities"], d["scales"], d["rotations"], 1.0,
                       self.empty, v["view"], v["view"], 1.0, 1.0, v["dcol"], v["dinv"], self.campos, False)
        return st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,5")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--p5", type=int, default=5_000_000)
    ap.add_argument("--img5", type=int, default=8192)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    have_ref = R.available() and not args.no_ref
    rows = []
    want = [int(x) for x in args.configs.split(",")]

    def report(name, ms_ours, ms_ref, extra):
        row = dict(config=name, ours_ms=round(ms_ours, 4), ref_ms=None if ms_ref is None else round(ms_ref, 4),
                   speedup=None if ms_ref is None else round(ms_ref / ms_ours, 2), **extra)
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)

    if 1 in want and world == 1:
        for kind in ("trained", "init"):
            V = Views(dev, 50_000, kind, 1337)
            v = V.view(S.make_camera(1337), 512, 512, 1337)
            st = V.ours(v)
            report(f"1: 50k {kind}, 512^2 fwd+bwd", timed(lambda: V.ours(v), args.reps, flush),
                   timed(lambda: V.ref(v), args.reps, flush) if have_ref else None, dict(instances=st.num_rendered))

    if 2 in want and world == 1:
        V = Views(dev, 300_000, "trained", 1337)
        views = [V.view(S.make_camera(1337 + i), 2048, 2048, 1337 + i) for i in range(19)]
        inst = sum(V.ours(v, backward=False).num_rendered for v in views)
        report("2: 300k, 19 views 2048^2, one iteration (19 fwd+bwd)",
               timed(lambda: [V.ours(v) for v in views], args.reps, flush),
               timed(lambda: [V.ref(v) for v in views], args.reps, flush) if have_ref else None,
               dict(instances=inst))
        del views

    if 3 in want and world == 1:
        V = Views(dev, 1_000_000, "trained", 1337)
        cam = S.make_camera(1337)
        trio = [V.view(cam, 2048, 2048, 1), V.view(S.sun_camera(cam), 4096, 4096, 2),
                V.view(S.random_camera(cam, 0.01, 3), 2048, 2048, 3)]
        inst = [V.ours(v, backward=False).num_rendered for v in trio]
        for nm, v in zip(("main 2048^2", "sun 4096^2", "random 2048^2"), trio):
            report(f"3: 1M, {nm} fwd+bwd", timed(lambda: V.ours(v), args.reps, flush),
                   timed(lambda: V.ref(v), args.reps, flush) if have_ref else None,
                   dict(instances=V.ours(v, backward=False).num_rendered))
        report("3: 1M, iteration = main + sun@2x + random, fwd+bwd",
               timed(lambda: [V.ours(v) for v in trio], args.reps, flush),
               timed(lambda: [V.ref(v) for v in trio], args.reps, flush) if have_ref else None,
               dict(instances=sum(inst)))
        del trio

    if 5 in want:
        P5, IMG5 = args.p5, args.img5
        V = Views(dev, P5, "trained", 1337)
        A = torch.tensor([[1 / 0.72, 0, 0], [0, 1 / 0.72, 0], [0, 0, S.METRES_PER_UNIT]])      # nadir camera
        v = V.view(S.affine_to_viewmatrix(A, torch.zeros(3)), IMG5, IMG5, 5)
        grid_y = (IMG5 + 15) // 16
        if world == 1:
            st = V.ours(v, backward=False)
            inst = st.num_rendered
            del st
            ms_ref = None
            if have_ref:
                try:
                    ms_ref = timed(lambda: V.ref(v, backward=False), max(2, args.reps // 2), flush, warm=1)
                except Exception as ex:                       # reference may exceed its 32-bit / memory limits
                    print("reference failed on config 5:", str(ex)[:200], file=sys.stderr)
            report(f"5: {P5/1e6:g}M, {IMG5}^2 forward (altitude/DSM), 1 GPU whole image",
                   timed(lambda: V.ours(v, backward=False), args.reps, flush), ms_ref, dict(instances=inst))
            # the same image as 8 bands rendered one after the other on this GPU (sharding overhead)
            bands = B.split_rows(grid_y, 8)
            report(f"5: same, 8 bands sequentially on 1 GPU",
                   timed(lambda: [V.ours(v, backward=False, band=b) for b in bands], args.reps, flush), None,
                   dict(instances=inst))
        else:
            import torch.distributed as dist
            d = V.d

            def step(weights=None):
                return B.forward_band(V.bg, d["means3D"], v["colors"], d["opacities"], d["scales"], d["rotations"],
                                      1.0, V.empty, v["view"], IMG5, IMG5, rank, world, weights=weights)
            color, invd, st = step()
            # balanced bands from the per-row instance counts of this (static) view: every rank counts its rows
            ex = E.export_state(st)
            r = ex["ranges"].to(torch.int64)
            per_row = (r[:, 1] - r[:, 0]).view(-1, (IMG5 + 15) // 16).sum(1)
            all_rows = torch.zeros(grid_y, dtype=torch.int64, device=dev)
            all_rows[st.rows[0]:st.rows[1]] = per_row
            dist.all_reduce(all_rows)
            weights = all_rows.cpu().tolist()
            for nm, w in (("even rows", None), ("instance-balanced rows", weights)):
                for _ in range(2):
                    step(w)
                ts = []
                for _ in range(args.reps):
                    flush.fill_(1)
                    torch.cuda.synchronize(); dist.barrier()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record(); step(w); e.record()
                    torch.cuda.synchronize()
                    t = torch.tensor([s.elapsed_time(e)], device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ts.append(float(t.item()))
                report(f"5: {P5/1e6:g}M, {IMG5}^2 forward, {world} GPUs = {world} tile bands + all-gather, {nm}",
                       statistics.median(ts), None, dict(instances=int(all_rows.sum().item()), max_over_ranks=True))
            # checksum so runs at different N can be compared: the gathered image must not depend on N
            if rank == 0:
                print(json.dumps({"config5_checksum": float(color.double().sum().item()),
                                  "alt_checksum": float(color[3].double().abs().sum().item())}), flush=True)

    if 6 in want and world == 1:
        from types import SimpleNamespace
        from eogs2_b200 import fused as FU
        from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
        P6, IMG6 = 1_000_000, 2048
        sc = S.make_scene(P6, "trained", 1337)
        C0 = FU.SH_C0
        raw = dict(xyz=sc.means3D, fdc=((sc.rgb - 0.5) / C0).unsqueeze(1), op=torch.logit(sc.opacities),
                   sc=torch.log(sc.scales), rot=sc.rotations * 1.7)
        raw = {k: v.to(dev).requires_grad_(True) for k, v in raw.items()}
        view = S.make_camera(1337).to(dev).requires_grad_(True)
        bg = S.background(1337).to(dev)
        dcol = S.upstream_grads(5, IMG6, IMG6, 1337, False)[0].to(dev)
        pipe = SimpleNamespace(debug=False, antialiasing=False, compute_cov3D_python=False, require_radii=True)
        pc = SimpleNamespace(_xyz=raw["xyz"], _features_dc=raw["fdc"], _opacity=raw["op"], _scaling=raw["sc"],
                             _rotation=raw["rot"], active_sh_degree=0)
        cam = SimpleNamespace(world_view_transform=view, full_proj_transform=view, affine=view, image_width=IMG6,
                              image_height=IMG6, FoVx=0.5, FoVy=0.5, camera_center=torch.zeros(3, device=dev),
                              learn_wv_only_lastparam=False, image_name="synthetic")

        def zero():
            for t in list(raw.values()) + [view]:
                t.grad = None

        def glue_unfused():
            zero()
            sp = torch.zeros_like(raw["xyz"], requires_grad=True) + 0
            sp.retain_grad()
            rs = GaussianRasterizationSettings(IMG6, IMG6, 0.25, 0.25, bg, 1.0, view, view, 0, cam.camera_center,
                                               False, False, False)
            rgb = (raw["fdc"] * C0 + 0.5).squeeze(1)
            alt = (raw["xyz"] @ view[:3, :3] + view[3, :3])[..., 2].unsqueeze(-1)
            colors = torch.cat([rgb, alt, torch.ones_like(alt)], dim=-1)
            img, radii, _ = GaussianRasterizer(rs)(
                means3D=raw["xyz"], means2D=sp, opacities=torch.sigmoid(raw["op"]), colors_precomp=colors,
                scales=torch.exp(raw["sc"]), rotations=torch.nn.functional.normalize(raw["rot"]))
            (img * dcol).sum().backward()

        def glue_fused():
            zero()
            out = FU.render_fused(cam, pc, pipe, bg)
            (out["render"] * dcol).sum().backward()

        report("6: 1M, 2048^2 render() fwd+bwd from raw parameters: torch glue + rasterizer vs fused glue (ref_ms = unfused)",
               timed(glue_fused, args.reps, flush), timed(glue_unfused, args.reps, flush), dict(instances=None))

    if 7 in want and world == 1:
        from eogs2_b200 import shadow as SHD
        Hc = Wc = 2048
        g = torch.Generator().manual_seed(7)
        virt = torch.randn(5, 2 * Hc, 2 * Wc, generator=g).to(dev).requires_grad_(True)
        u, v = torch.meshgrid(torch.linspace(-1, 1, Wc), torch.linspace(-1, 1, Hc), indexing="xy")
        uva = torch.stack([u, v, torch.rand(Hc, Wc, generator=g) * 60 - 20], -1).to(dev).requires_grad_(True)
        M = torch.tensor([[0.5, 0.0, -0.0015], [0.0, 0.5, -0.00125], [0.0, 0.0, 1.0]], device=dev).requires_grad_(True)
        w_rgb = torch.randn(3, Hc, Wc, generator=g).to(dev)
        w_alt = torch.randn(Hc, Wc, generator=g).to(dev)

        def torch_path():
            for t in (virt, uva, M):
                t.grad = None
            uv = torch.einsum("...ij,...j->...i", M, uva)[..., :2]
            smp = torch.nn.functional.grid_sample(virt.unsqueeze(0), uv.unsqueeze(0), align_corners=True).squeeze(0)
            alt = smp[3]
            alt[(uv.abs() > 1).any(-1)] = -100
            ((smp[:3] * w_rgb).sum() + (alt * w_alt).sum()).backward()

        def fused_path():
            for t in (virt, uva, M):
                t.grad = None
            rgb, alt, uv = SHD.resample_virtual(virt, M, uva)
            ((rgb * w_rgb).sum() + (alt * w_alt).sum()).backward()

        report("7: sun-view resample 4096^2 -> 2048^2 fwd+bwd: torch einsum+grid_sample+mask vs fused kernel (ref_ms = torch)",
               timed(fused_path, args.reps, flush), timed(torch_path, args.reps, flush), dict(instances=None))

    if 8 in want and world == 1:
        import torch.nn.functional as TF
        from eogs2_b200 import losses as LS
        g = torch.Generator().manual_seed(8)
        gt = torch.rand(3, 2048, 2048, generator=g).to(dev)
        img = (gt + 0.1 * torch.randn(3, 2048, 2048, generator=g).to(dev)).clamp(0, 1).requires_grad_(True)
        w1 = LS.gaussian_window().unsqueeze(1)
        window = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0).expand(3, 1, 11, 11).contiguous().to(dev)

        def torch_loss():
            img.grad = None
            conv = lambda x: TF.conv2d(x, window, padding=5, groups=3)
            mu1, mu2 = conv(img), conv(gt)
            mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
            s1, s2, s12 = conv(img * img) - mu1_sq, conv(gt * gt) - mu2_sq, conv(img * gt) - mu12
            ssim_map = ((2 * mu12 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1_sq + mu2_sq + 1e-4) * (s1 + s2 + 9e-4))
            (0.8 * torch.abs(img - gt).mean() + 0.2 * (1.0 - ssim_map.mean())).backward()

        def fused_loss():
            img.grad = None
            LS.photometric_loss(img, gt, 0.2).backward()

        report("8: photometric loss 3x2048^2 fwd+bwd: torch L1 + conv2d SSIM vs fused kernels (ref_ms = torch)",
               timed(fused_loss, args.reps, flush), timed(torch_loss, args.reps, flush), dict(instances=None))

    if 9 in want and world == 1:
        from types import SimpleNamespace
        sys.path.insert(0, str(ROOT / "tests"))
        import iteration_ref as IR
        from eogs2_b200 import iteration as ITR
        from eogs2_b200 import optim as OPT
        P9, IMG9 = 1_000_000, 2048
        lrs = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "opacity": 5e-2, "scaling": 5e-3, "rotation": 1e-3}
        pipe = SimpleNamespace(debug=False, antialiasing=False, compute_cov3D_python=False, require_radii=False)
        bg = S.background(9).to(dev)
        cam, sun, cam2sun = IR.make_cameras(dev, 1337, IMG9, IMG9)
        init = IR.raw_params(dev, P9, 1337)
        gt = torch.rand(3, IMG9, IMG9, generator=torch.Generator().manual_seed(9)).to(dev)
        ref_p = {n: torch.nn.Parameter(p.clone()) for n, p in init.items()}
        ref_opt = torch.optim.Adam([{"params": [ref_p[n]], "lr": lrs[n], "name": n} for n in ref_p], lr=0.0, eps=1e-15)
        opt = OPT.FlatGaussianAdam(init, lrs)

        def torch_iter():
            ref_opt.zero_grad(set_to_none=True)
            loss, _ = ITR.camera_iteration(cam, sun, cam2sun, ITR.model_view(ref_p), pipe, bg, gt, render_fn=IR.torch_render,
                                           resample_fn=IR.torch_resample, loss_fn=IR.torch_photometric)
            loss.backward()
            ref_opt.step()

        def fused_iter():
            opt.zero_grad()
            loss, _ = ITR.camera_iteration(cam, sun, cam2sun, ITR.model_view(opt.params), pipe, bg, gt)
            loss.backward()
            opt.step()

        report("9: full camera iteration 1M, 2048^2 (main + sun@2x renders, resample, shading, L1+DSSIM, backward, Adam): "
               "torch stages on our rasterizer vs all fused (ref_ms = torch stages)",
               timed(fused_iter, args.reps, flush), timed(torch_iter, args.reps, flush), dict(instances=None))

    if args.out and rank == 0:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(rows, indent=1))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
