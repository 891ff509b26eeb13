"""GPU: one EOGS++ camera iteration composed from the fused kernels (eogs2_b200/iteration.py) against the same
iteration composed from the reference's torch stages (tests/iteration_ref.py) on the unfused rasterizer:
loss value, every parameter gradient, and a few optimiser steps (FlatGaussianAdam vs torch.optim.Adam)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import iteration_ref as R
from eogs2_b200 import iteration as IT
from eogs2_b200 import optim as O
from eogs2_b200 import scene as S

pytestmark = pytest.mark.gpu
LRS = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "opacity": 5e-2, "scaling": 5e-3, "rotation": 1e-3}


def rel(a, b):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def test_fused_iteration_matches_the_torch_iteration_and_trains(cuda_dev):
    dev, P, W, H = cuda_dev, 40_000, 256, 192
    pipe = SimpleNamespace(debug=False, antialiasing=False, compute_cov3D_python=False, require_radii=False)
    bg = S.background(3).to(dev)
    cam, sun, cam2sun = R.make_cameras(dev, 3, W, H)
    init = R.raw_params(dev, P, 3)
    g = torch.Generator().manual_seed(11)
    gt = torch.rand(3, H, W, generator=g).to(dev)

    # --- reference composition (torch stages, unfused rasterizer, torch Adam)
    ref_p = {n: torch.nn.Parameter(p.clone()) for n, p in init.items()}
    ref_opt = torch.optim.Adam([{"params": [ref_p[n]], "lr": LRS[n], "name": n} for n in ref_p], lr=0.0, eps=1e-15)
    # --- fused composition
    opt = O.FlatGaussianAdam(init, LRS)

    losses_f, losses_r = [], []
    for it in range(4):
        ref_opt.zero_grad(set_to_none=True)
        loss_r, _ = IT.camera_iteration(cam, sun, cam2sun, IT.model_view(ref_p), pipe, bg, gt, render_fn=R.torch_render,
                                        resample_fn=R.torch_resample, loss_fn=R.torch_photometric)
        loss_r.backward()
        opt.zero_grad()
        loss_f, aux = IT.camera_iteration(cam, sun, cam2sun, IT.model_view(opt.params), pipe, bg, gt)
        loss_f.backward()
        losses_f.append(float(loss_f.detach())); losses_r.append(float(loss_r.detach()))
        if it == 0:
            assert abs(losses_f[0] - losses_r[0]) <= 1e-5 * abs(losses_r[0]), (losses_f[0], losses_r[0])
            for n in ref_p:
                assert rel(opt.params[n].grad, ref_p[n].grad) < 2e-3, (n, rel(opt.params[n].grad, ref_p[n].grad))
            assert aux["viewspace_points"].grad is not None
        ref_opt.step()
        opt.step()
    # both compositions follow the same trajectory and the loss goes down
    assert all(abs(a - b) <= 1e-4 * abs(b) for a, b in zip(losses_f, losses_r)), (losses_f, losses_r)
    assert losses_f[-1] < losses_f[0]
    for n in ref_p:
        assert rel(opt.params[n], ref_p[n]) < 1e-4, n


def test_odd_gaussian_counts_after_prune_keep_every_view_aligned(cuda_dev):
    """After prune() the number of Gaussians is arbitrary.  The flat layout (dp.segment_layout) starts every segment
    on a 16-byte boundary, so the rotation view the geometry kernel loads as float4 stays aligned for odd P, and the
    composed iteration keeps running (it raised 'rotations must be 16-byte aligned' for every odd P before)."""
    dev, P, W, H = cuda_dev, 10_001, 128, 96
    pipe = SimpleNamespace(debug=False, antialiasing=False, compute_cov3D_python=False, require_radii=False)
    bg = S.background(4).to(dev)
    cam, sun, cam2sun = R.make_cameras(dev, 4, W, H)
    opt = O.FlatGaussianAdam(R.raw_params(dev, P, 4), LRS)
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(5)).to(dev)
    for keep_every in (None, 3):                       # P = 10 001 (odd), then 6 667 survivors (odd)
        if keep_every:
            keep = torch.ones(opt.P, dtype=torch.bool, device=dev)
            keep[::keep_every] = False
            opt.prune(keep)
        assert opt.P % 2 == 1
        for n in opt.names:
            assert opt.params[n].data_ptr() % 16 == 0, (n, opt.P)
        opt.zero_grad()
        loss, _ = IT.camera_iteration(cam, sun, cam2sun, IT.model_view(opt.params), pipe, bg, gt)
        loss.backward()
        assert all(torch.isfinite(opt.params[n].grad).all() for n in opt.names)
        opt.step()


def test_misaligned_quaternion_views_are_handled_not_faulted(cuda_dev):
    """A rotations INPUT view with an odd storage offset is legal for the reference (it indexes floats): it is copied
    into an aligned buffer.  A misaligned rotations OUTPUT buffer (out=) is refused with a Python error instead of a
    'misaligned address' fault that would poison the CUDA context."""
    import eogs2_b200 as E
    from eogs2_b200._cabi import EogsRasterError
    dev, P, W, H = cuda_dev, 2001, 96, 80
    sc = S.make_scene(P, "trained", 8).to(dev)
    view = S.make_camera(8).to(dev)
    colors = S.colors_precomp(S.make_scene(P, "trained", 8), S.make_camera(8)).to(dev)
    bg = S.background(8).to(dev)
    empty = torch.empty(0, device=dev)
    store = torch.zeros(4 * P + 1, device=dev)
    rot_view = store[1:].view(P, 4)                    # 4-byte aligned only
    rot_view.copy_(sc.rotations)
    assert rot_view.data_ptr() % 16 != 0
    st_a = E.rasterize_forward_raw(bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0, empty, view, H, W)
    st_b = E.rasterize_forward_raw(bg, sc.means3D, colors, sc.opacities, sc.scales, rot_view, 1.0, empty, view, H, W)
    assert torch.equal(st_a.color, st_b.color)
    dcol, dinv = (t.to(dev) for t in S.upstream_grads(5, H, W, 8, True))
    g_ref = E.rasterize_backward_raw(st_b, bg, sc.means3D, colors, sc.opacities, sc.scales, rot_view, 1.0, empty, view, view,
                                     dcol, dinv)
    bad = torch.zeros(4 * P + 1, device=dev)[1:].view(P, 4)
    with pytest.raises(EogsRasterError, match="16-byte aligned"):
        E.rasterize_backward_raw(st_b, bg, sc.means3D, colors, sc.opacities, sc.scales, rot_view, 1.0, empty, view, view,
                                 dcol, dinv, out={"rotations": bad})
    good = torch.zeros(P, 4, device=dev)
    g = E.rasterize_backward_raw(st_b, bg, sc.means3D, colors, sc.opacities, sc.scales, rot_view, 1.0, empty, view, view,
                                 dcol, dinv, out={"rotations": good})
    torch.cuda.synchronize()                            # the context is still healthy
    assert g[6].data_ptr() == good.data_ptr() and rel(g[6], g_ref[6]) < 1e-5
