"""Data parallelism over views for the rasterizer path.

The reference trains on one GPU, one camera per optimiser step (train_pan.py:236-670,
utils/general_utils.py:155 pins cuda:0) and has no communication code.  Views are independent
units of work, so the path shards naturally: Gaussians are replicated, rank r rasterises the
cameras r, r+k, r+2k, ... of the iteration's view batch (forward + backward), and there is ONE
exchange step per iteration — an all-reduce (SUM) of a single flat fp32 bucket holding the
gradients of the Adam parameter groups (xyz 3, f_dc 3, opacity 1, scaling 3, rotation 4 =
14 floats per Gaussian, scene/gaussian_model.py:228-259) plus the camera-matrix gradients —
before the replicated optimiser step.  One process per GPU (torchrun), `torch.distributed` with
the NCCL backend over NVLink 5 / NVSwitch on a B200 box; the same code runs on `gloo` for the
CPU tests.  A k-GPU step equals a 1-GPU step that accumulates the same k views (up to the
all-reduce's summation order).

Nothing here touches the kernels: `render_fn(params, camera) -> loss` is whatever the caller
uses (the EOGS++ render() + losses, or bench.py's synthetic loss).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Sequence

import torch
import torch.distributed as dist

PARAM_ORDER = ("xyz", "f_dc", "opacity", "scaling", "rotation")   # Adam groups, gaussian_model.py:228-259


SEGMENT_ALIGN = 4      # floats: every segment of a flat buffer starts on a 16-byte boundary


def segment_layout(sizes: Sequence[tuple]) -> tuple:
    """[(name, numel), ...] -> ({name: (begin, end)}, total).  Segment starts are aligned to 16 bytes, whatever the
    number of Gaussians: the kernels read quaternions (and write their gradients) as float4, and a rotation segment
    that starts after 10 * P floats would be misaligned for every odd P (after any prune).  The same rule lays out
    the gradient bucket (GradBucket) and the optimiser's flat parameter buffer (optim.FlatGaussianAdam), so the
    all-reduced bucket can be handed to the optimiser step as is."""
    slices, off = {}, 0
    for name, k in sizes:
        off = (off + SEGMENT_ALIGN - 1) // SEGMENT_ALIGN * SEGMENT_ALIGN
        slices[name] = (off, off + int(k))
        off += int(k)
    return slices, off


def init_distributed(backend: str | None = None) -> tuple[int, int, int]:
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun); returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_views(num_views: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership of the iteration's views: rank r takes r, r+world, ..."""
    return list(range(rank, num_views, world))


@dataclass
class GradBucket:
    """One flat fp32 buffer for every gradient that must be exchanged, laid out as
    [xyz | f_dc | opacity | scaling | rotation | extras...]: a single collective per step."""
    params: Dict[str, torch.Tensor]
    extras: Dict[str, torch.Tensor] = field(default_factory=dict)   # e.g. per-camera viewmatrix params
    # "nccl": dist.all_reduce on a plain tensor.  "auto" (CUDA + an initialised process group): the bucket is placed
    # in symmetric memory and the library's own NVLS kernel is used when a start-up calibration finds it faster
    # than ncclAllReduce (eogs2_b200/nvls.py: 8 GPUs 0.20 vs 0.27 ms for 64 MB).
    exchange: str = "nccl"

    def __post_init__(self):
        self.names = [n for n in PARAM_ORDER if n in self.params] + \
                     sorted(n for n in self.params if n not in PARAM_ORDER)
        self.slices, off = segment_layout([(n, self.params[n].numel()) for n in self.names] +
                                          [("extra:" + n, self.extras[n].numel()) for n in sorted(self.extras)])
        any_p = next(iter(self.params.values()))
        self._exchange, self.exchange_name = None, "ncclAllReduce"
        if self.exchange == "auto" and any_p.is_cuda and dist.is_initialized() and dist.get_world_size() > 1:
            from .nvls import make_grad_exchange
            padded = (off + 3) // 4 * 4                       # 16-byte multimem accesses
            whole, self._exchange, self.exchange_name = make_grad_exchange(padded, any_p.device)
            self._whole = whole                               # keeps the symmetric allocation alive
            self.flat = whole[:off]
        else:
            self.flat = torch.zeros(off, dtype=torch.float32, device=any_p.device)

    def pack(self) -> torch.Tensor:
        for n in self.names:
            a, b = self.slices[n]
            g = self.params[n].grad
            if g is None:
                self.flat[a:b].zero_()
            else:
                self.flat[a:b].copy_(g.reshape(-1))
        for n, t in self.extras.items():
            a, b = self.slices["extra:" + n]
            if t.grad is None:
                self.flat[a:b].zero_()
            else:
                self.flat[a:b].copy_(t.grad.reshape(-1))
        return self.flat

    def all_reduce(self, average: bool = False) -> torch.Tensor:
        if dist.is_initialized() and dist.get_world_size() > 1:
            if self._exchange is not None:
                self._exchange()
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if average:
                self.flat.div_(dist.get_world_size())
        return self.flat

    def unpack(self) -> None:
        for n in self.names:
            a, b = self.slices[n]
            p = self.params[n]
            p.grad = self.flat[a:b].view_as(p).clone()
        for n, t in self.extras.items():
            a, b = self.slices["extra:" + n]
            t.grad = self.flat[a:b].view_as(t).clone()


def dp_backward(params: Dict[str, torch.Tensor], cameras: Sequence, render_fn: Callable,
                rank: int, world: int, bucket: GradBucket | None = None, average: bool = False) -> float:
    """One data-parallel gradient evaluation: this rank renders its share of `cameras` forward +
    backward (gradients accumulate in .grad), then the bucket is all-reduced and written back so
    that every rank holds the gradient of sum_views loss — what a single GPU would have after
    accumulating all the views.  Returns this rank's summed loss."""
    for p in params.values():
        p.grad = None
    if bucket is not None:
        for t in bucket.extras.values():
            t.grad = None
    total = 0.0
    for i in shard_views(len(cameras), rank, world):
        loss = render_fn(params, cameras[i])
        loss.backward()
        total += float(loss.detach())
    if bucket is None:
        bucket = GradBucket(params)
    bucket.pack()
    bucket.all_reduce(average)
    bucket.unpack()
    return total


class DensificationStats:
    """The per-view statistics EOGS++ accumulates between densifications, kept consistent across data-parallel ranks.

    On one GPU the reference updates, after every view (train_pan.py:681-690, scene/gaussian_model.py:719-723):

        max_radii2D[vis]         = max(max_radii2D[vis], radii[vis])
        xyz_gradient_accum[vis] += || viewspace_points.grad[vis, :2] ||
        denom[vis]              += 1

    With views sharded over ranks each rank sees only its own views, so replicas that densify on their local statistics
    diverge (different clone / split decisions -> different Gaussian counts).  `update()` applies the reference's three
    lines to this rank's LOCAL deltas; `all_reduce()` — one exchange step next to the gradient all-reduce — merges the
    deltas of all ranks (MAX for the radii, SUM for the two accumulators) into the replicated totals, after which every
    rank holds exactly what a single GPU would hold after processing the same views (the per-view gradient norms are
    taken BEFORE any reduction: a norm of summed gradients would be a different statistic).  Only needed when
    `only_prune=False`."""

    def __init__(self, num_points: int, device):
        self.max_radii2D = torch.zeros(num_points, device=device)
        self.xyz_gradient_accum = torch.zeros(num_points, 1, device=device)
        self.denom = torch.zeros(num_points, 1, device=device)
        self._d_radii = torch.zeros(num_points, device=device)
        self._d_sum = torch.zeros(num_points, 2, device=device)          # [gradient-norm sum, view count]

    def update(self, viewspace_point_grad: torch.Tensor, radii: torch.Tensor, visibility_filter: torch.Tensor) -> None:
        """One rendered view of THIS rank (arguments as in train_pan.py:681-690)."""
        vis = visibility_filter.reshape(-1)
        self._d_radii[vis] = torch.max(self._d_radii[vis], radii[vis].to(self._d_radii.dtype))
        self._d_sum[vis, 0] += torch.norm(viewspace_point_grad[vis, :2], dim=-1)
        self._d_sum[vis, 1] += 1

    def all_reduce(self, group=None) -> None:
        """Merge the deltas of all ranks into the replicated totals (2 small collectives: MAX [P], SUM [P, 2])."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self._d_radii, op=dist.ReduceOp.MAX, group=group)
            dist.all_reduce(self._d_sum, op=dist.ReduceOp.SUM, group=group)
        self.max_radii2D = torch.max(self.max_radii2D, self._d_radii)
        self.xyz_gradient_accum += self._d_sum[:, 0:1]
        self.denom += self._d_sum[:, 1:2]
        self._d_radii.zero_()
        self._d_sum.zero_()

    def reset(self, num_points: int) -> None:
        """After densify_and_prune (densification_postfix, gaussian_model.py:569-571)."""
        self.__init__(num_points, self.max_radii2D.device)


def replicated_adam(params: Dict[str, torch.Tensor], lrs: Dict[str, float]) -> torch.optim.Optimizer:
    """Adam with eps = 1e-15 and one group per parameter, as GaussianModel.training_setup
    (scene/gaussian_model.py:228-262).  Every rank steps the same optimiser on the same reduced
    gradients, so parameters stay bit-identical across ranks without a broadcast."""
    groups = [{"params": [params[n]], "lr": lrs.get(n, 1e-3), "name": n} for n in params]
    return torch.optim.Adam(groups, lr=0.0, eps=1e-15)


class HostInputPipeline:
    """Double-buffered host -> device staging of a step's input tensors.

    `submit(host_tensors)` enqueues the copies of the NEXT step's pinned host tensors on a side
    stream into one of two preallocated device slots; `get()` makes the current stream wait for
    them and returns the device tensors.  Calling `get()` then `submit()` at the top of every step
    overlaps the PCIe transfer of step s+1 with the kernels of step s.  Plain torch plumbing
    (streams + events): it knows nothing about the rasterizer and works for any dict of tensors.

    A slot is overwritten two submits later; the copy first waits (on the device) for everything
    the consumer stream had enqueued when the following `get()` was called, so the usual
    "use the tensors within the step" pattern is race-free.
    """

    def __init__(self, device: torch.device):
        self.device = device
        self.copy_stream = torch.cuda.Stream(device)
        self.slots: List[Dict[str, torch.Tensor] | None] = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [None, None]             # event on the consumer stream after which the slot may be rewritten
        self.next_slot = 0
        self.pending: List[int] = []

    def submit(self, host_tensors: Dict[str, torch.Tensor]) -> None:
        s = self.next_slot
        self.next_slot ^= 1
        if self.slots[s] is None or any(self.slots[s][k].shape != v.shape or self.slots[s][k].dtype != v.dtype
                                        for k, v in host_tensors.items()):
            # Allocate ON the copy stream: a block taken from the consumer stream's pool may have been
            # freed there a moment ago (e.g. the rasterizer's binning scratch) with kernels that still
            # use it queued on the consumer stream — the copy stream would overwrite it under them.
            with torch.cuda.device(self.device), torch.cuda.stream(self.copy_stream):
                self.slots[s] = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device)
                                 for k, v in host_tensors.items()}
            for t in self.slots[s].values():
                t.record_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.copy_stream):
            if self.free[s] is not None:
                self.copy_stream.wait_event(self.free[s])
            for k, v in host_tensors.items():
                self.slots[s][k].copy_(v, non_blocking=True)
            self.ready[s].record(self.copy_stream)
        self.pending.append(s)

    def get(self) -> Dict[str, torch.Tensor]:
        if not self.pending:
            raise RuntimeError("HostInputPipeline.get() without a pending submit()")
        s = self.pending.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[s])
        # the OTHER slot's previous consumer work is already enqueued on `cur`: mark it reusable
        ev = torch.cuda.Event()
        ev.record(cur)
        self.free[s ^ 1] = ev
        return self.slots[s]
