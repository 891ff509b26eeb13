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
