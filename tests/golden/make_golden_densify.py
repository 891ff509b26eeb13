"""Golden vectors for SURVEY.md section 8 row N2 (densification): the REFERENCE's own

    scene/gaussian_model.py:223-271   GaussianModel.training_setup        (Adam, eps 1e-15, one group per tensor)
    scene/gaussian_model.py:451-539   _prune_optimizer / prune_points / cat_tensors_to_optimizer
    scene/gaussian_model.py:540-668   densification_postfix / densify_and_split / densify_and_clone
    scene/gaussian_model.py:672-704   densify_and_prune
    scene/gaussian_model.py:719-723   add_densification_stats

executed here on the CPU of the build container from /root/reference (tests/golden/ref_import.py maps the hard-coded
device="cuda" of the tensor factories to the CPU).  The model takes two Adam steps first so the moments that the
tensor surgery has to carry along are non-trivial.

The split samples: `torch.normal(mean=0, std=stds)` (:597) draws N(0,1) and scales by std; the proxy draws the
standard normals itself (recorded as `split_noise`, in the reference's row order) and returns mean + std * noise, so
the GPU test can hand the same draws to FlatGaussianAdam.densify_and_prune(noise=...).

    python tests/golden/make_golden_densify.py          # writes tests/golden/densify_ref.npz
"""
from __future__ import annotations

import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))

import ref_import                                   # noqa: E402

OUT = HERE / "densify_ref.npz"
NAMES = (("xyz", "_xyz"), ("f_dc", "_features_dc"), ("opacity", "_opacity"), ("scaling", "_scaling"), ("rotation", "_rotation"))
LRS = dict(position_lr_init=1.6e-4, feature_lr=2.5e-3, opacity_lr=5e-2, scaling_lr=5e-3, rotation_lr=1e-3)
CASES = {   # tag: (P, seed, max_screen_size)
    "a": (800, 21, None),
    "b": (400, 22, 20),
}
ARGS = dict(grad_threshold=2e-4, min_opacity=0.3, screen_size_threshold=1.0, scene_extent=1.0)
PERCENT_DENSE = 0.02


def make_model(R, P, seed):
    g = torch.Generator().manual_seed(seed)
    pc = R.GaussianModel(0)
    pc._xyz = torch.nn.Parameter(torch.randn(P, 3, generator=g))
    pc._features_dc = torch.nn.Parameter(torch.randn(P, 1, 3, generator=g))
    pc._features_rest = torch.nn.Parameter(torch.zeros(P, 0, 3))
    pc._opacity = torch.nn.Parameter(torch.randn(P, 1, generator=g))
    pc._scaling = torch.nn.Parameter(torch.randn(P, 3, generator=g) * 0.8 - 4.0)      # exp ~ 0.002 .. 0.2
    pc._rotation = torch.nn.Parameter(torch.randn(P, 4, generator=g))
    pc._exposure = torch.nn.Parameter(torch.eye(3, 4)[None])
    pc.max_radii2D = torch.zeros(P)
    pc.spatial_lr_scale = 1.0
    pc.training_setup(SimpleNamespace(percent_dense=PERCENT_DENSE, **LRS))
    return pc, g


def state_of(pc, prefix, out):
    for name, attr in NAMES:
        p = getattr(pc, attr)
        st = pc.optimizer.state[p]
        out[f"{prefix}_{name}"] = p.detach().numpy().copy()
        out[f"{prefix}_m_{name}"] = st["exp_avg"].numpy().copy()
        out[f"{prefix}_v_{name}"] = st["exp_avg_sq"].numpy().copy()


def generate() -> dict:
    R = ref_import.load()
    proxy = R.gaussian_model.torch
    out = {"percent_dense": PERCENT_DENSE, **{"lr_" + k: v for k, v in LRS.items()}, **{"arg_" + k: v for k, v in ARGS.items()}}
    for tag, (P, seed, max_screen_size) in CASES.items():
        pc, g = make_model(R, P, seed)
        for name, attr in NAMES:
            out[f"{tag}_init_{name}"] = getattr(pc, attr).detach().numpy().copy()
        for it in range(2):
            for name, attr in NAMES:
                p = getattr(pc, attr)
                p.grad = torch.randn(p.shape, generator=g) * 1e-3
                out[f"{tag}_grad{it}_{name}"] = p.grad.numpy().copy()
            pc.optimizer.step()
            pc.optimizer.zero_grad(set_to_none=True)
        state_of(pc, f"{tag}_before", out)
        # statistics as train_pan.py:681-690 accumulates them: add_densification_stats over a few "views"
        for view in range(3):
            vsp = SimpleNamespace(grad=torch.randn(P, 3, generator=g) * 2.5e-4)
            upd = torch.rand(P, generator=g) < 0.6
            radii = (torch.rand(P, generator=g) * 30).int() * upd
            pc.max_radii2D[upd] = torch.max(pc.max_radii2D[upd], radii[upd].float())
            pc.add_densification_stats(vsp, upd)
        out[f"{tag}_accum"] = pc.xyz_gradient_accum.numpy().copy()
        out[f"{tag}_denom"] = pc.denom.numpy().copy()
        out[f"{tag}_max_radii2D"] = pc.max_radii2D.numpy().copy()
        noise_log = []

        def normal(mean, std, **kw):
            z = torch.randn(std.shape, generator=g)
            noise_log.append(z)
            return mean + std * z

        proxy.normal = normal
        try:
            pc.densify_and_prune(ARGS["grad_threshold"], ARGS["min_opacity"], ARGS["screen_size_threshold"], max_screen_size,
                                 radii, ARGS["scene_extent"])
        finally:
            del proxy.normal
        assert len(noise_log) == 1
        out[f"{tag}_split_noise"] = noise_log[0].numpy().copy()
        out[f"{tag}_max_screen_size"] = -1 if max_screen_size is None else max_screen_size
        state_of(pc, f"{tag}_after", out)
        assert pc.xyz_gradient_accum.abs().sum() == 0 and pc.denom.shape[0] == pc.get_xyz.shape[0]
        print(tag, "P", P, "->", pc.get_xyz.shape[0], "split samples", noise_log[0].shape[0])
    return out


if __name__ == "__main__":
    data = generate()
    np.savez_compressed(OUT, **data)
    print(OUT, OUT.stat().st_size, "bytes;", len(data), "arrays")
