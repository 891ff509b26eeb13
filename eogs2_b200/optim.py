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
from .dp import PARAM_ORDER


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
        self.slices, off = {}, 0
        for n in self.names:
            self.slices[n] = (off, off + P * self.width[n])
            off += P * self.width[n]
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

    def zero_grad(self) -> None:
        for p in self.params.values():
            p.grad = None

    # ---- step -----------------------------------------------------------------------------------
    def step(self, flat_grads: Optional[torch.Tensor] = None) -> None:
        """One Adam step of every group.  flat_grads: a [numel] fp32 tensor in this layout (the all-reduced
        bucket of dp.GradBucket); by default the views' .grad are packed (missing grads count as zero)."""
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
        ends = (C.c_ulonglong * k)(*[self.slices[n][1] for n in self.names])
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
