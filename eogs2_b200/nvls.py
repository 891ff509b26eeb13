"""Gradient exchange of the data-parallel step through the NVSwitch: a flat fp32 bucket in
symmetric memory, all-reduced by the library's own NVLS kernel (csrc/nvls.cu: multimem.ld_reduce + multimem.st on
this rank's 1/N slice) instead of ncclAllReduce.

`SymmetricBucket(numel, device)` allocates the bucket with torch.distributed._symmetric_memory (one allocation per
rank, rendezvoused into a multicast mapping) — plumbing; the collective itself is the sm_100a kernel.
`all_reduce()` = symmetric-memory barrier (every rank's backward has written its gradients) -> kernel -> barrier
(every slice has landed on every rank), all stream-ordered on the current stream, no host synchronisation.
Where multicast is unavailable (one GPU, no NVSwitch, older driver) `SymmetricBucket.create` returns None and
callers keep the NCCL path of eogs2_b200/dp.py — a different collective, never a CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import _cabi


class SymmetricBucket:
    def __init__(self, flat: torch.Tensor, handle, group):
        self.flat, self.handle, self.group = flat, handle, group
        self.rank, self.world = int(handle.rank), int(handle.world_size)
        self._phase = 0
        # every rank's copy of the bucket as mapped into this process (peer-to-peer variant of the collective)
        ptrs = [int(p) for p in getattr(handle, "buffer_ptrs", [])]
        self._peer_ptrs = (C.c_void_p * self.world)(*ptrs) if len(ptrs) == self.world and self.world <= 8 and all(ptrs) else None

    @classmethod
    def create(cls, numel: int, device: torch.device, group=None) -> Optional["SymmetricBucket"]:
        """Collective over `group` (default: WORLD).  None when NVLS multicast cannot be set up."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
            return None
        if numel % 4:
            raise ValueError("the bucket must hold a multiple of 4 floats (16-byte multimem accesses)")
        try:
            import torch.distributed._symmetric_memory as symm_mem
            group = group if group is not None else dist.group.WORLD
            with torch.cuda.device(device):
                flat = symm_mem.empty(numel, dtype=torch.float32, device=device)
                handle = symm_mem.rendezvous(flat, group)
            ok = int(getattr(handle, "multicast_ptr", 0) or 0) != 0
        except Exception:                                    # noqa: BLE001 — any set-up failure means "not available"
            ok, flat, handle = False, None, None
        # every rank must take the same path
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            return None
        flat.zero_()
        return cls(flat, handle, group)

    def all_reduce(self) -> torch.Tensor:
        """In-place SUM over ranks of self.flat, on the current stream."""
        lib = _cabi.load()
        dev = self.flat.device
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            # two barrier channels alternate so that consecutive all-reduces never share signal slots
            self.handle.barrier(channel=self._phase, timeout_ms=20000)
            _cabi.check(lib.eogs_nvls_allreduce(C.c_void_p(stream), C.c_void_p(int(self.handle.multicast_ptr)),
                                                C.c_ulonglong(self.flat.numel()), self.rank, self.world),
                        "eogs_nvls_allreduce")
            self.handle.barrier(channel=self._phase + 1, timeout_ms=20000)
            self._phase ^= 2
        return self.flat

    @property
    def has_p2p(self) -> bool:
        return self._peer_ptrs is not None

    def all_reduce_p2p(self) -> torch.Tensor:
        """Same contract through the peer-to-peer kernel (csrc/nvls.cu: p2p_allreduce_kernel): this rank sums its slice out
        of every rank's copy over NVLink and stores it into every copy — the faster one on 2 GPUs."""
        if self._peer_ptrs is None:
            raise _cabi.EogsRasterError("peer buffer pointers are not available: use all_reduce() or NCCL")
        lib = _cabi.load()
        dev = self.flat.device
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            self.handle.barrier(channel=self._phase, timeout_ms=20000)
            _cabi.check(lib.eogs_p2p_allreduce(C.c_void_p(stream), self._peer_ptrs, C.c_ulonglong(self.flat.numel()),
                                               self.rank, self.world), "eogs_p2p_allreduce")
            self.handle.barrier(channel=self._phase + 1, timeout_ms=20000)
            self._phase ^= 2
        return self.flat


def make_grad_exchange(numel: int, device: torch.device, group=None, calibrate: bool = True):
    """The data-parallel step's gradient exchange: returns (flat_bucket, all_reduce, name).

    `flat_bucket` is the fp32 tensor the backward kernels write into (`rasterize_backward_raw(..., out=views)`),
    `all_reduce()` sums it over ranks in place on the current stream.  With NVLS available the bucket lives in
    symmetric memory and both collectives are timed once on it (a few launches each, max over ranks): the own
    multimem kernel wins on an 8-GPU NVSwitch domain (0.197 vs 0.270 ms for 64 MB, profiles/r03u), ncclAllReduce
    on 2 GPUs, so the faster one is used and named.  Without NVLS: a plain tensor and ncclAllReduce."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
        flat = torch.zeros(numel, dtype=torch.float32, device=device)
        return flat, (lambda: flat), "none (single rank)"
    sym = SymmetricBucket.create(numel, device, group)
    if sym is None:
        flat = torch.zeros(numel, dtype=torch.float32, device=device)
        return flat, (lambda: dist.all_reduce(flat, group=group)), "ncclAllReduce (NVLS multicast unavailable)"
    flat = sym.flat

    def nccl():
        dist.all_reduce(flat, group=group)
    choice = "nvls"
    if calibrate:
        def timed(fn, reps=8):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(device)
            dist.barrier(group=group)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize(device)
            return e0.elapsed_time(e1) / reps
        use_p2p = sym.has_p2p and sym.world <= 4                          # (larger worlds: the switch's reduction wins)
        t = torch.tensor([timed(sym.all_reduce), timed(nccl), timed(sym.all_reduce_p2p) if use_p2p else 1e9],
                         dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)            # same decision on every rank
        best = int(torch.argmin(t).item())
        choice = ("nvls", "nccl", "p2p")[best]
        flat.zero_()
        timing = f" (calibration: own NVLS kernel {float(t[0]):.3f} ms, NCCL {float(t[1]):.3f} ms" + \
                 (f", own peer-to-peer kernel {float(t[2]):.3f} ms)" if use_p2p else ")")
    else:
        timing = ""
    if choice == "p2p":
        return flat, sym.all_reduce_p2p, "own peer-to-peer kernel: NVLink loads + stores over symmetric memory" + timing
    if choice == "nvls":
        return flat, sym.all_reduce, "own NVLS kernel: multimem.ld_reduce + multimem.st over symmetric memory" + timing
    return flat, nccl, "ncclAllReduce" + timing
