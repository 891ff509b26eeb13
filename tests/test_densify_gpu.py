"""GPU: FlatGaussianAdam.densify_and_prune (csrc/optim.cu densify_* kernels) against (1) the reference's own run
(tests/golden/densify_ref.npz, generated from /root/reference by tests/golden/make_golden_densify.py) and (2), at
larger sizes and on the CUDA random stream, a torch restatement of
GaussianModel.densify_and_clone / densify_and_split / densify_and_prune + cat_tensors_to_optimizer /
_prune_optimizer (scene/gaussian_model.py:451-539, 573-704) on plain tensors, with the same random stream."""
import pytest
import torch

from eogs2_b200 import optim as O
from test_optim_gpu import LRS, SHAPES, grads_like

pytestmark = pytest.mark.gpu


def build_rotation(r):                                      # utils/general_utils.py:82-105
    q = r / torch.sqrt((r * r).sum(1))[:, None]
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.zeros((q.size(0), 3, 3), device=r.device)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - w * z); R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y); R[:, 2, 1] = 2 * (y * z + w * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def reference_densify_and_prune(p, m, v, accum, denom, grad_threshold, min_opacity, screen_size_threshold,
                                max_screen_size, scene_extent, percent_dense, N, generator):
    """p / m / v: dicts of parameters and Adam moments.  Returns the new dicts."""
    grads = accum / denom
    grads[grads.isnan()] = 0.0
    scaling = lambda: torch.exp(p["scaling"])              # noqa: E731  get_scaling
    # densify_and_clone (:626-668)
    sel = (torch.norm(grads, dim=-1) >= grad_threshold) & (scaling().max(dim=1).values <= percent_dense * scene_extent)
    for n in p:
        new = p[n][sel]
        p[n] = torch.cat([p[n], new]); m[n] = torch.cat([m[n], torch.zeros_like(new)]); v[n] = torch.cat([v[n], torch.zeros_like(new)])
    # densify_and_split (:573-624)
    n_init = p["xyz"].shape[0]
    padded = torch.zeros(n_init, device=grads.device)
    padded[:grads.shape[0]] = grads.squeeze()
    sel = (padded >= grad_threshold) & (scaling().max(dim=1).values > percent_dense * scene_extent)
    stds = scaling()[sel].repeat(N, 1)
    samples = torch.normal(mean=torch.zeros((stds.size(0), 3), device=stds.device), std=stds, generator=generator)
    rots = build_rotation(p["rotation"][sel]).repeat(N, 1, 1)
    new = {"xyz": torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + p["xyz"][sel].repeat(N, 1),
           "scaling": torch.log(scaling()[sel].repeat(N, 1) / (0.8 * N)),
           "rotation": p["rotation"][sel].repeat(N, 1), "f_dc": p["f_dc"][sel].repeat(N, 1, 1),
           "opacity": p["opacity"][sel].repeat(N, 1)}
    for n in p:
        p[n] = torch.cat([p[n], new[n]]); m[n] = torch.cat([m[n], torch.zeros_like(new[n])]); v[n] = torch.cat([v[n], torch.zeros_like(new[n])])
    prune_filter = torch.cat([sel, torch.zeros(N * int(sel.sum()), device=sel.device, dtype=torch.bool)])
    for d in (p, m, v):
        for n in d:
            d[n] = d[n][~prune_filter]
    # densify_and_prune tail (:690-700); max_radii2D was reset to zeros by densification_postfix (:571)
    prune_mask = (torch.sigmoid(p["opacity"]) < min_opacity).squeeze()
    if max_screen_size:
        big_vs = torch.zeros_like(prune_mask)
        big_ws = scaling().max(dim=1).values > 0.1 * screen_size_threshold
        prune_mask = prune_mask | big_vs | big_ws
    for d in (p, m, v):
        for n in d:
            d[n] = d[n][~prune_mask]
    return p, m, v


@pytest.mark.parametrize("P,max_screen_size,seed", [(20_011, None, 1), (5_000, 20, 2), (64, None, 3)])
def test_densify_and_prune_matches_reference(cuda_dev, P, max_screen_size, seed):
    dev = cuda_dev
    g = torch.Generator().manual_seed(seed)
    init = {n: torch.randn((P,) + s, generator=g).to(dev) for n, s in SHAPES.items()}
    init["scaling"] = (torch.randn((P, 3), generator=g) * 0.8 - 4.0).to(dev)         # exp ~ 0.002 .. 0.2
    flat = O.FlatGaussianAdam(init, LRS)
    for it in range(2):                                     # non-trivial moments
        gr = grads_like(dev, init, 50 + it, 1e-3)
        flat.step(flat.pack(gr))
    p = {n: flat.params[n].detach().clone() for n in flat.names}
    m = {n: flat.exp_avg[slice(*flat.slices[n])].view_as(p[n]).clone() for n in flat.names}
    v = {n: flat.exp_avg_sq[slice(*flat.slices[n])].view_as(p[n]).clone() for n in flat.names}
    accum = (torch.rand((P, 1), generator=g) * 4e-4).to(dev)
    denom = torch.randint(0, 3, (P, 1), generator=g).float().to(dev)                 # zeros -> NaN -> 0
    args = dict(grad_threshold=2e-4, min_opacity=0.3, screen_size_threshold=1.0, max_screen_size=max_screen_size,
                scene_extent=1.0, percent_dense=0.02, N=2)
    rp, rm, rv = reference_densify_and_prune(p, m, v, accum.clone(), denom, generator=torch.Generator(dev).manual_seed(99), **args)
    new = flat.densify_and_prune(accum, denom, generator=torch.Generator(dev).manual_seed(99), **args)
    assert flat.last_densify["cloned"] > 0 and flat.last_densify["split"] > 0 or P < 100
    for n in flat.names:
        assert new[n].shape == rp[n].shape, (n, new[n].shape, rp[n].shape)
        a, b = flat.slices[n]
        if n in ("xyz", "scaling"):
            assert torch.allclose(new[n], rp[n], rtol=1e-6, atol=1e-6), n
        else:
            assert torch.equal(new[n], rp[n]), n
        assert torch.equal(flat.exp_avg[a:b].view_as(new[n]), rm[n]), n
        assert torch.equal(flat.exp_avg_sq[a:b].view_as(new[n]), rv[n]), n
    # the model keeps training after densification
    for n in flat.params:
        flat.params[n].grad = torch.ones_like(flat.params[n])
    flat.step()
    assert all(torch.isfinite(q).all() for q in flat.params.values())


def test_densify_is_rank_reproducible(cuda_dev):
    """Two replicas with the same generator seed stay bit-identical (the data-parallel contract)."""
    dev = cuda_dev
    g = torch.Generator().manual_seed(5)
    init = {n: torch.randn((3000,) + s, generator=g).to(dev) for n, s in SHAPES.items()}
    init["scaling"] = (torch.randn((3000, 3), generator=g) - 3.0).to(dev)
    accum = (torch.rand((3000, 1), generator=g) * 4e-4).to(dev); denom = torch.ones(3000, 1, device=dev)
    outs = []
    for _ in range(2):
        flat = O.FlatGaussianAdam(init, LRS)
        outs.append(flat.densify_and_prune(accum, denom, 2e-4, 0.005, 1.0, None, 1.0, 0.02,
                                           generator=torch.Generator(dev).manual_seed(1234)))
    for n in outs[0]:
        assert torch.equal(outs[0][n], outs[1][n])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_densify_and_adam_match_the_reference_run(cuda_dev, tag):
    """tests/golden/densify_ref.npz: the reference's own GaussianModel (training_setup, two Adam steps,
    add_densification_stats, densify_and_prune; scene/gaussian_model.py:223-271,451-704,719-723) executed on the CPU
    of the build container by tests/golden/make_golden_densify.py.  Same initial tensors, gradients, statistics and
    split draws here: the fused Adam must land on the reference's parameters / moments, and the densification on its
    rows — order, copies, children and carried moments."""
    import numpy as np
    from pathlib import Path
    dev = cuda_dev
    z = np.load(Path(__file__).parent / "golden" / "densify_ref.npz")
    t = lambda k: torch.from_numpy(z[k]).to(dev)                                   # noqa: E731
    names = list(SHAPES)
    lrs = {"xyz": float(z["lr_position_lr_init"]), "f_dc": float(z["lr_feature_lr"]), "opacity": float(z["lr_opacity_lr"]),
           "scaling": float(z["lr_scaling_lr"]), "rotation": float(z["lr_rotation_lr"])}
    flat = O.FlatGaussianAdam({n: t(f"{tag}_init_{n}") for n in names}, lrs)
    for it in range(2):
        flat.step(flat.pack({n: t(f"{tag}_grad{it}_{n}") for n in names}))
    for n in names:
        a, b = flat.slices[n]
        ref_p, ref_m, ref_v = t(f"{tag}_before_{n}"), t(f"{tag}_before_m_{n}"), t(f"{tag}_before_v_{n}")
        assert torch.allclose(flat.params[n], ref_p, rtol=1e-6, atol=1e-7), n
        assert torch.allclose(flat.exp_avg[a:b].view_as(ref_m), ref_m, rtol=1e-6, atol=1e-12), n
        assert torch.allclose(flat.exp_avg_sq[a:b].view_as(ref_v), ref_v, rtol=1e-6, atol=1e-15), n
        # continue from the reference's exact state so the row comparison below is not blurred by Adam rounding
        flat.flat[a:b].copy_(ref_p.reshape(-1)); flat.exp_avg[a:b].copy_(ref_m.reshape(-1)); flat.exp_avg_sq[a:b].copy_(ref_v.reshape(-1))
    mss = int(z[f"{tag}_max_screen_size"])
    new = flat.densify_and_prune(t(f"{tag}_accum"), t(f"{tag}_denom"), float(z["arg_grad_threshold"]), float(z["arg_min_opacity"]),
                                 float(z["arg_screen_size_threshold"]), None if mss < 0 else mss, float(z["arg_scene_extent"]),
                                 float(z["percent_dense"]), N=2, noise=t(f"{tag}_split_noise"))
    assert flat.last_densify["cloned"] > 0 and 2 * flat.last_densify["split"] == z[f"{tag}_split_noise"].shape[0]
    for n in names:
        ref_p, ref_m, ref_v = t(f"{tag}_after_{n}"), t(f"{tag}_after_m_{n}"), t(f"{tag}_after_v_{n}")
        assert new[n].shape == ref_p.shape, (n, new[n].shape, ref_p.shape)
        a, b = flat.slices[n]
        if n in ("xyz", "scaling"):                                                # children: R(q) * (std * z) + mean, log(std / 1.6)
            assert torch.allclose(new[n], ref_p, rtol=1e-6, atol=1e-6), n
        else:
            assert torch.equal(new[n], ref_p), n
        assert torch.equal(flat.exp_avg[a:b].view_as(ref_m), ref_m), n
        assert torch.equal(flat.exp_avg_sq[a:b].view_as(ref_v), ref_v), n
