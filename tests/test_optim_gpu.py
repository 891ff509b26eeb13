"""GPU: FlatGaussianAdam (eogs2_b200/optim.py, csrc/optim.cu) against torch.optim.Adam with EOGS++'s setup
(one group per parameter, eps 1e-15, scene/gaussian_model.py:223-271) and against boolean-mask pruning of
parameters and optimizer state (_prune_optimizer, :466-486)."""
import numpy as np
import pytest
import torch

from eogs2_b200 import optim as O

pytestmark = pytest.mark.gpu
SHAPES = {"xyz": (3,), "f_dc": (1, 3), "opacity": (1,), "scaling": (3,), "rotation": (4,)}
LRS = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "opacity": 5e-2, "scaling": 5e-3, "rotation": 1e-3}


def make(dev, P, seed):
    g = torch.Generator().manual_seed(seed)
    return {n: torch.randn((P,) + s, generator=g).to(dev) for n, s in SHAPES.items()}


def grads_like(dev, params, seed, scale):
    g = torch.Generator().manual_seed(seed)
    return {n: (torch.randn(p.shape, generator=g) * scale).to(dev) for n, p in params.items()}


def torch_adam(params):
    ps = {n: torch.nn.Parameter(p.clone()) for n, p in params.items()}
    opt = torch.optim.Adam([{"params": [ps[n]], "lr": LRS[n], "name": n} for n in ps], lr=0.0, eps=1e-15)
    return ps, opt


def close(a, b, tol):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    return np.linalg.norm(a - b) <= tol * (np.linalg.norm(b) + 1e-30)


def test_flat_adam_matches_torch_adam_and_pruning(cuda_dev):
    dev, P = cuda_dev, 10_007
    init = make(dev, P, 1)
    ref_p, ref_opt = torch_adam(init)
    flat = O.FlatGaussianAdam(init, LRS)
    assert all(flat.params[n].shape == init[n].shape and flat.params[n].is_leaf for n in init)
    for it in range(6):
        gr = grads_like(dev, init, 100 + it, 1e-3 * (it + 1))
        if it == 3:                                        # xyz learning-rate schedule
            ref_opt.param_groups[0]["lr"] = 4e-5
            flat.set_lr("xyz", 4e-5)
        for n in ref_p:
            ref_p[n].grad = gr[n].clone()
        ref_opt.step()
        if it % 2 == 0:                                    # gradients through .grad of the views ...
            for n in flat.params:
                flat.params[n].grad = gr[n].clone()
            flat.step()
        else:                                              # ... or as one flat bucket (the all-reduced DP buffer)
            flat.step(flat.pack(gr))
        for n in ref_p:
            assert close(flat.params[n], ref_p[n], 2e-6), (it, n)

    # prune: keep a random 60 %
    g = torch.Generator().manual_seed(7)
    keep = (torch.rand(P, generator=g) < 0.6).to(dev)
    new = flat.prune(keep)
    for gi, grp in enumerate(ref_opt.param_groups):
        n = grp["name"]
        st = ref_opt.state[grp["params"][0]]
        assert new[n].shape[0] == int(keep.sum())
        assert close(new[n], ref_p[n][keep], 2e-6)
        a, b = flat.slices[n]
        assert close(flat.exp_avg[a:b].view_as(new[n]), st["exp_avg"][keep], 2e-6)
        assert close(flat.exp_avg_sq[a:b].view_as(new[n]), st["exp_avg_sq"][keep], 2e-6)
    # exactness of the compaction itself: gathered rows are copies
    before = {n: flat.params[n].detach().clone() for n in flat.names}
    keep2 = torch.ones(flat.P, dtype=torch.bool, device=dev); keep2[::3] = False
    new2 = flat.prune(keep2)
    for n in flat.names:
        assert torch.equal(new2[n], before[n][keep2])
    # and a step after pruning still works
    for n in flat.params:
        flat.params[n].grad = torch.ones_like(flat.params[n])
    flat.step()
    assert flat.P == int(keep2.sum()) and all(torch.isfinite(p).all() for p in flat.params.values())


def test_prune_edge_cases(cuda_dev):
    dev = cuda_dev
    flat = O.FlatGaussianAdam(make(dev, 50, 2), LRS)
    flat.prune(torch.zeros(50, dtype=torch.bool, device=dev))          # everything pruned
    assert flat.P == 0 and flat.params["xyz"].shape == (0, 3)
    flat.step()                                                         # no-op on an empty model
    flat2 = O.FlatGaussianAdam(make(dev, 33, 3), LRS)
    ref = flat2.params["rotation"].detach().clone()
    flat2.prune(torch.ones(33, dtype=torch.bool, device=dev))           # nothing pruned
    assert torch.equal(flat2.params["rotation"], ref)
    with pytest.raises(Exception):
        O.FlatGaussianAdam({k: v.cpu() for k, v in make(dev, 4, 4).items()}, LRS)
