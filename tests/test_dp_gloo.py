"""CPU, world_size 2, gloo: the data-parallel-over-views harness (eogs2_b200/dp.py).
A 2-rank step must produce the gradients, and after Adam the parameters, of a 1-process run that
accumulates the same views; parameters must stay identical across ranks."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from eogs2_b200 import dp  # noqa: E402


def make_params(seed=0, P=257):
    g = torch.Generator().manual_seed(seed)
    shapes = {"xyz": (P, 3), "f_dc": (P, 3), "opacity": (P, 1), "scaling": (P, 3), "rotation": (P, 4)}
    return {n: torch.randn(*s, generator=g).requires_grad_(True) for n, s in shapes.items()}


def make_cameras(n=5):
    g = torch.Generator().manual_seed(42)
    return [torch.randn(4, 4, generator=g) for _ in range(n)]


def render_fn(params, cam):
    """Stand-in for render()+loss: any differentiable function of all parameters and the camera."""
    p = params["xyz"] @ cam[:3, :3] + cam[3, :3]
    s = torch.exp(params["scaling"]).sum(1, keepdim=True) * torch.sigmoid(params["opacity"])
    q = params["rotation"] / params["rotation"].norm(dim=1, keepdim=True)
    return ((p * s).sum(1) * q[:, 0] + params["f_dc"].sum(1)).pow(2).mean()


def single_process_reference(steps=3):
    params, cams = make_params(), make_cameras()
    opt = dp.replicated_adam(params, {"xyz": 1e-2, "f_dc": 5e-3})
    grads0 = None
    for it in range(steps):
        opt.zero_grad(set_to_none=True)
        for c in cams:
            render_fn(params, c).backward()
        if it == 0:
            grads0 = {n: p.grad.clone() for n, p in params.items()}
        opt.step()
    return grads0, {n: p.detach().clone() for n, p in params.items()}


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = dp.init_distributed("gloo")
    params, cams = make_params(), make_cameras()
    extra = torch.zeros(4, 4, requires_grad=True)
    bucket = dp.GradBucket(params, {"cam0": extra})
    opt = dp.replicated_adam(params, {"xyz": 1e-2, "f_dc": 5e-3})
    grads0 = None
    for it in range(3):
        dp.dp_backward(params, cams, render_fn, r, w, bucket)
        if it == 0:
            grads0 = {n: p.grad.clone() for n, p in params.items()}
        opt.step()
    torch.save({"grads0": grads0, "params": {n: p.detach() for n, p in params.items()},
                "shard": dp.shard_views(len(cams), r, w), "bucket_numel": bucket.flat.numel()}, out + f".{rank}")
    dist.destroy_process_group()


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_two_rank_step_equals_single_process_accumulation(tmp_path):
    out = str(tmp_path / "res")
    mp.spawn(worker, args=(2, free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    g_ref, p_ref = single_process_reference()
    assert r0["shard"] == [0, 2, 4] and r1["shard"] == [1, 3]
    # 14 floats per Gaussian + the camera extras, every segment starting on a 16-byte boundary (odd P: 6 floats of padding)
    sizes = [("xyz", 771), ("f_dc", 771), ("opacity", 257), ("scaling", 771), ("rotation", 1028), ("extra:cam0", 16)]
    slices, total = dp.segment_layout(sizes)
    assert r0["bucket_numel"] == total == 257 * 14 + 16 + 6 and all(a % 4 == 0 for a, _ in slices.values())
    for n in g_ref:
        assert torch.allclose(r0["grads0"][n], g_ref[n], rtol=1e-5, atol=1e-7), n
        assert torch.equal(r0["grads0"][n], r1["grads0"][n]), n         # identical on every rank
        assert torch.equal(r0["params"][n], r1["params"][n]), n         # replicas never drift
        assert torch.allclose(r0["params"][n], p_ref[n], rtol=1e-4, atol=1e-6), n


def stats_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = dp.init_distributed("gloo")
    P, views = 101, stats_views()
    st = dp.DensificationStats(P, "cpu")
    for it in range(2):                                   # two iterations between densifications
        for i in dp.shard_views(len(views), r, w):
            st.update(*views[i])
        st.all_reduce()
    torch.save({"max_radii2D": st.max_radii2D, "accum": st.xyz_gradient_accum, "denom": st.denom}, out + f".{rank}")
    dist.destroy_process_group()


def stats_views(P=101, n=5):
    g = torch.Generator().manual_seed(3)
    views = []
    for _ in range(n):
        radii = torch.randint(0, 30, (P,), generator=g).int() * (torch.rand(P, generator=g) < 0.7)
        views.append((torch.randn(P, 3, generator=g), radii, (radii > 0).nonzero()))
    return views


def test_densification_statistics_match_a_single_gpu_run(tmp_path):
    """max_radii2D (MAX) and xyz_gradient_accum / denom (SUM) after a 2-rank run equal the reference's single-GPU
    bookkeeping over the same views (train_pan.py:681-690, gaussian_model.py:719-723), on every rank."""
    out = str(tmp_path / "stats")
    mp.spawn(stats_worker, args=(2, free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    P, views = 101, stats_views()
    max_radii2D, accum, denom = torch.zeros(P), torch.zeros(P, 1), torch.zeros(P, 1)
    for it in range(2):
        for grad, radii, vis in views:                    # the reference's three lines, view after view
            v = vis.reshape(-1)
            max_radii2D[v] = torch.max(max_radii2D[v], radii[v].float())
            accum[v] += torch.norm(grad[v, :2], dim=-1, keepdim=True)
            denom[v] += 1
    for r in (r0, r1):
        assert torch.equal(r["max_radii2D"], max_radii2D) and torch.equal(r["denom"], denom)
        assert torch.allclose(r["accum"], accum, rtol=1e-6, atol=1e-7)
    assert torch.equal(r0["accum"], r1["accum"])          # replicas hold identical statistics


def test_shard_views_partition():
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 19):
            parts = [dp.shard_views(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
