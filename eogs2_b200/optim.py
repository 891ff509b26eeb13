"""Replicated Adam and prune compaction on one flat parameter buffer.

EOGS++ (`GaussianModel.training_setup`, scene/gaussian_model.py:223-271) builds one Adam group per
parameter tensor (eps 1e-15) and rebuilds every parameter and both moment tensors with boolean-mask
indexing on each prune (`_prune_optimizer` / `prune_points`, :466-505).  `FlatGaussianAdam` keeps the
parameters segment by segment — xyz | f_dc | opacity | scaling | rotation (| others) — in ONE flat
fp32 buffer, the layout of the data-parallel gradient bucket (`dp.GradBucket`), so that

  * `step()` is one kernel for all groups (per-group learning rates, `set_lr` for the xyz schedule)
    reading the gradients — `.grad` of the views, or the all-reduced flat bucket — in place;
  * `prune(keep_mask)` is a scan of the keep flags + one gather per segment for parameters and both
    moments, and hands back fresh parameter views (what `prune_points` assigns to `self._xyz`, ...).

The parameter tensors it exposes are leaf views into the flat buffer, so the rasterizer reads them
without copies.  Every rank runs the same step on the same reduced gradients: replicas stay
bit-identical without a broadcast.  No CPU path (the kernels live in libeogs_raster.so).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _cabi
from .dp import PARAM_ORDER, segment_layout


class FlatGaussianAdam:
    def __init__(self, params: Dict[str, torch.Tensor], lrs: Dict[str, float], betas=(0.9, 0.999), eps: float = 1e-15):
        any_p = next(iter(params.values()))
        if not any_p.is_cuda:
            raise _cabi.EogsRasterError("FlatGaussianAdam needs CUDA tensors: the optimiser step has no CPU path")
        self.device = any_p.device
        self.names = [n for n in PARAM_ORDER if n in params] + sorted(n for n in params if n not in PARAM_ORDER)
        if len(self.names) > 8:
            raise ValueError("at most 8 parameter segments")
        self.P = int(any_p.shape[0])
        self.row_shape = {n: tuple(params[n].shape[1:]) for n in self.names}
        self.width = {n: int(params[n][0].numel()) if self.P else int(torch.Size(self.row_shape[n]).numel()) for n in self.names}
        self.lrs = {n: float(lrs.get(n, 0.0)) for n in self.names}
        self.betas, self.eps, self.t = betas, float(eps), 0
        self._alloc(self.P)
        for n in self.names:
            a, b = self.slices[n]
            self.flat[a:b].copy_(params[n].detach().reshape(-1).to(torch.float32))
        self._make_views()

    # ---- layout ---------------------------------------------------------------------------------
    def _alloc(self, P: int) -> None:
        # 16-byte aligned segment starts for ANY P (dp.segment_layout): the rasterizer reads the rotation view as float4
        self.slices, off = segment_layout([(n, P * self.width[n]) for n in self.names])
        self.numel = off
        with torch.cuda.device(self.device):
            self.flat = torch.zeros(off, dtype=torch.float32, device=self.device)
            self.exp_avg = torch.zeros_like(self.flat)
            self.exp_avg_sq = torch.zeros_like(self.flat)
        self.P = P

    def _make_views(self) -> None:
        self.params: Dict[str, torch.Tensor] = {}
        for n in self.names:
            a, b = self.slices[n]
            self.params[n] = self.flat[a:b].view((self.P,) + self.row_shape[n]).detach().requires_grad_(True)

    def set_lr(self, name: str, lr: float) -> None:
        """e.g. the exponential xyz schedule (update_learning_rate, scene/gaussian_model.py:273-279)."""
        self.lrs[name] = float(lr)

    def pack(self, tensors: Dict[str, torch.Tensor]) -> torch.Tensor:
        """{name: per-parameter tensor} -> one flat fp32 tensor in this optimiser's (16-byte aligned) segment layout,
        e.g. gradients to hand to step(flat_grads=...).  Segments that are absent stay zero."""
        with torch.cuda.device(self.device):
            out = torch.zeros(self.numel, dtype=torch.float32, device=self.device)
        for n, t in tensors.items():
            a, b = self.slices[n]
            out[a:b].copy_(t.reshape(-1))
        return out

    def zero_grad(self) -> None:
        for p in self.params.values():
            p.grad = None

    # ---- step -----------------------------------------------------------------------------------
    def step(self, flat_grads: Optional[torch.Tensor] = None) -> None:
        """One Adam step of every group.  flat_grads: a [numel] fp32 tensor in this layout (the all-reduced
        bucket of dp.GradBucket); by default the views' .grad are packed.  A view whose .grad is None is stepped with a
        ZERO gradient (its moments decay and it keeps moving along them) — torch.optim.Adam, which the reference uses,
        skips such a parameter entirely; the rasterizer's backward always produces every gradient, so the two only
        differ for segments the caller never renders."""
        lib = _cabi.load()
        if self.numel == 0:
            return
        if flat_grads is None:
            with torch.cuda.device(self.device):
                flat_grads = torch.zeros_like(self.flat)
            for n in self.names:
                g = self.params[n].grad
                if g is not None:
                    a, b = self.slices[n]
                    flat_grads[a:b].copy_(g.reshape(-1))
        if flat_grads.numel() < self.numel or flat_grads.dtype != torch.float32 or not flat_grads.is_contiguous():
            raise _cabi.EogsRasterError("flat_grads must be a contiguous float32 tensor covering the parameter layout")
        self.t += 1
        k = len(self.names)
        # a segment ends where the next one begins: the (zero-gradient, zero-valued) alignment padding rides along
        starts = [self.slices[n][0] for n in self.names]
        ends = (C.c_ulonglong * k)(*(starts[1:] + [self.numel]))
        lrs = (C.c_float * k)(*[self.lrs[n] for n in self.names])
        with torch.cuda.device(self.device), torch.no_grad():
            _cabi.check(lib.eogs_adam_step(
                torch.cuda.current_stream(self.device).cuda_stream, self.numel, k, ends, lrs, self.betas[0], self.betas[1],
                self.eps, self.t, self.flat.data_ptr(), flat_grads.data_ptr(), self.exp_avg.data_ptr(),
                self.exp_avg_sq.data_ptr()), "eogs_adam_step")

    # ---- prune ----------------------------------------------------------------------------------
    def prune(self, keep_mask: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Keep the rows where keep_mask is True (prune_points passes ~mask, gaussian_model.py:488-490).
        Returns the new parameter views; moments are compacted the same way."""
        lib = _cabi.load()
        if keep_mask.numel() != self.P:
            raise ValueError("keep_mask must have one entry per Gaussian")
        dev, P = self.device, self.P
        with torch.cuda.device(dev), torch.no_grad():
            stream = torch.cuda.current_stream(dev).cuda_stream
            keep = keep_mask.to(device=dev, dtype=torch.uint8).contiguous()
            offsets = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
            tmp_bytes = lib.eogs_prune_temp_bytes(P)
            tmp = torch.empty(tmp_bytes, dtype=torch.uint8, device=dev)
            count = torch.zeros(1, dtype=torch.int32, device=dev)
            _cabi.check(lib.eogs_prune_offsets(stream, P, keep.data_ptr(), offsets.data_ptr(), tmp.data_ptr(), tmp_bytes,
                                               count.data_ptr()), "eogs_prune_offsets")
            new_P = int(count.item())                    # one sync, like boolean-mask indexing in the reference
            old = (self.flat, self.exp_avg, self.exp_avg_sq, dict(self.slices))
            self._alloc(new_P)
            for src, dst in zip(old[:3], (self.flat, self.exp_avg, self.exp_avg_sq)):
                for n in self.names:
                    a, _ = old[3][n]
                    na, _ = self.slices[n]
                    if P and new_P:
                        _cabi.check(lib.eogs_prune_gather(stream, P, self.width[n], keep.data_ptr(), offsets.data_ptr(),
                                                          src.data_ptr() + 4 * a, dst.data_ptr() + 4 * na), "eogs_prune_gather")
        self._make_views()
        return self.params

    # ---- densify --------------------------------------------------------------------------------
    def _scan_flags(self, lib, stream, flags: torch.Tensor):
        P = flags.numel()
        offsets = torch.empty(max(P, 1), dtype=torch.int32, device=self.device)
        tmp_bytes = lib.eogs_prune_temp_bytes(P)
        tmp = torch.empty(tmp_bytes, dtype=torch.uint8, device=self.device)
        count = torch.zeros(1, dtype=torch.int32, device=self.device)
        _cabi.check(lib.eogs_prune_offsets(stream, P, flags.data_ptr(), offsets.data_ptr(), tmp.data_ptr(), tmp_bytes,
                                           count.data_ptr()), "eogs_prune_offsets")
        return offsets, count

    def densify_and_prune(self, xyz_gradient_accum: torch.Tensor, denom: torch.Tensor, grad_threshold: float,
                          min_opacity: float, screen_size_threshold: float, max_screen_size, scene_extent: float,
                          percent_dense: float = 0.01, N: int = 2,
                          generator: Optional[torch.Generator] = None,
                          noise: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """GaussianModel.densify_and_prune (scene/gaussian_model.py:672-704) on the flat buffers: clone the small
        Gaussians with a large view-space gradient, split the large ones into N samples of their own
        distribution (parents removed), then prune by opacity (and, when max_screen_size is given, by world size
        > 0.1 * screen_size_threshold).  Arguments follow the reference's order, with the two statistics tensors
        in place of `self` state and without `radii` (only stored in tmp_radii there).
        Row order is the reference's: [survivors | clones | split children, N blocks].  New rows get zero Adam
        moments (cat_tensors_to_optimizer, :507-539).  The caller resets xyz_gradient_accum / denom /
        max_radii2D to zeros of the new length like densification_postfix (:569-571).

        `generator`: CUDA generator for the split samples; seed it identically on every rank (e.g. from the
        iteration number) and the replicas stay bit-identical without a broadcast.
        `noise`: the (N * K_split, 3) standard-normal draws themselves, in the reference's row order (N blocks of the
        selected rows), instead of the generator — how tests/test_densify_gpu.py replays the reference's own run."""
        for need in ("xyz", "opacity", "scaling", "rotation"):
            if need not in self.slices:
                raise ValueError(f"densify_and_prune needs the '{need}' segment")
        lib, dev, P = _cabi.load(), self.device, self.P
        if xyz_gradient_accum.numel() != P or denom.numel() != P:
            raise ValueError("xyz_gradient_accum / denom must have one entry per Gaussian")
        with torch.cuda.device(dev), torch.no_grad():
            stream = torch.cuda.current_stream(dev).cuda_stream
            accum = xyz_gradient_accum.detach().to(torch.float32).reshape(-1).contiguous()
            den = denom.detach().to(torch.float32).reshape(-1).contiguous()
            clone_flag = torch.zeros(max(P, 1), dtype=torch.uint8, device=dev)
            split_flag = torch.zeros(max(P, 1), dtype=torch.uint8, device=dev)
            sa, _ = self.slices["scaling"]
            _cabi.check(lib.eogs_densify_select(stream, P, accum.data_ptr(), den.data_ptr(), self.flat.data_ptr() + 4 * sa,
                                                float(grad_threshold), float(percent_dense * scene_extent),
                                                clone_flag.data_ptr(), split_flag.data_ptr()), "eogs_densify_select")
            clone_off, clone_cnt = self._scan_flags(lib, stream, clone_flag[:P])
            split_off, split_cnt = self._scan_flags(lib, stream, split_flag[:P])
            Kc, Ks = (int(v) for v in torch.cat([clone_cnt, split_cnt]).tolist())      # one sync for both counts
            Pn = P + Kc + N * Ks
            if noise is None:
                noise = torch.randn((N * Ks, 3), device=dev, dtype=torch.float32, generator=generator)
            else:
                if tuple(noise.shape) != (N * Ks, 3):
                    raise ValueError(f"noise must have shape ({N * Ks}, 3) for {Ks} split Gaussians, got {tuple(noise.shape)}")
                noise = noise.to(device=dev, dtype=torch.float32).contiguous()

            old = (self.flat, self.exp_avg, self.exp_avg_sq, dict(self.slices))
            self._alloc(Pn)                                   # zeros: new rows start with zero moments
            for which, (src, dst) in enumerate(zip(old[:3], (self.flat, self.exp_avg, self.exp_avg_sq))):
                for n in self.names:
                    a, b = old[3][n]
                    na, _ = self.slices[n]
                    w = self.width[n]
                    dst[na:na + P * w].copy_(src[a:b])
                    if which != 0 or P == 0:
                        continue
                    if Kc:
                        _cabi.check(lib.eogs_prune_gather(stream, P, w, clone_flag.data_ptr(), clone_off.data_ptr(),
                                                          src.data_ptr() + 4 * a, dst.data_ptr() + 4 * (na + P * w)),
                                    "eogs_prune_gather(clone)")
                    for rep in range(N if Ks else 0):
                        row0 = P + Kc + rep * Ks
                        _cabi.check(lib.eogs_prune_gather(stream, P, w, split_flag.data_ptr(), split_off.data_ptr(),
                                                          src.data_ptr() + 4 * a, dst.data_ptr() + 4 * (na + row0 * w)),
                                    "eogs_prune_gather(split)")
            if Ks:
                row0 = P + Kc
                ptr = lambda n, wdt: self.flat.data_ptr() + 4 * (self.slices[n][0] + row0 * wdt)   # noqa: E731
                _cabi.check(lib.eogs_densify_split_children(stream, N * Ks, N, ptr("xyz", 3), ptr("scaling", 3),
                                                            ptr("rotation", 4), noise.data_ptr()),
                            "eogs_densify_split_children")
            keep = torch.empty(max(Pn, 1), dtype=torch.uint8, device=dev)
            ws = 0.1 * float(screen_size_threshold) if max_screen_size else -1.0
            _cabi.check(lib.eogs_densify_keep(stream, Pn, P, split_flag.data_ptr(),
                                              self.flat.data_ptr() + 4 * self.slices["opacity"][0],
                                              self.flat.data_ptr() + 4 * self.slices["scaling"][0],
                                              float(min_opacity), ws, keep.data_ptr()), "eogs_densify_keep")
        self.last_densify = {"cloned": Kc, "split": Ks, "rows_before_prune": Pn}
        return self.prune(keep[:Pn].bool())
