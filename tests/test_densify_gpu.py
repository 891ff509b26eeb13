"""GPU: FlatGaussianAdam.densify_and_prune (csrc/optim.cu densify_* kernels) against a torch restatement of
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
