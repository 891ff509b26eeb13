"""Multi-GPU check + timing of the NVLS all-reduce kernel against ncclAllReduce on the same data.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/nvls_check.py"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from eogs2_b200.nvls import SymmetricBucket      # noqa: E402

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 16 * 1_000_000 + 16
bucket = SymmetricBucket.create(n, dev)
out = {"world": world, "nvls": bucket is not None}
if bucket is not None:
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    errs = []
    for trial in range(3):
        x = torch.randn(n, device=dev, generator=g)
        ref = x.clone()
        dist.all_reduce(ref)
        bucket.flat.copy_(x)
        bucket.all_reduce()
        torch.cuda.synchronize()
        errs.append(float((bucket.flat - ref).abs().max() / ref.abs().max()))
        same = bucket.flat.clone()
        dist.broadcast(same, 0)
        assert torch.equal(same, bucket.flat), "ranks disagree on the reduced bucket"
    out["max_rel_err_vs_nccl"] = max(errs)
    out["p2p"] = bucket.has_p2p
    if bucket.has_p2p:                                # the peer-to-peer kernel: same checks
        errs = []
        for trial in range(3):
            x = torch.randn(n, device=dev, generator=g)
            ref = x.clone()
            dist.all_reduce(ref)
            bucket.flat.copy_(x)
            bucket.all_reduce_p2p()
            torch.cuda.synchronize()
            errs.append(float((bucket.flat - ref).abs().max() / ref.abs().max()))
            same = bucket.flat.clone()
            dist.broadcast(same, 0)
            assert torch.equal(same, bucket.flat), "ranks disagree on the reduced bucket (p2p)"
        out["p2p_max_rel_err_vs_nccl"] = max(errs)

    def timeit(fn, reps=30):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    y = torch.randn(n, device=dev)
    out["nccl_ms"] = timeit(lambda: dist.all_reduce(y))
    out["nvls_ms"] = timeit(bucket.all_reduce)
    if bucket.has_p2p:
        out["p2p_ms"] = timeit(bucket.all_reduce_p2p)
    import ctypes as C
    from eogs2_b200 import _cabi
    lib = _cabi.load()

    def kernel_only():
        _cabi.check(lib.eogs_nvls_allreduce(C.c_void_p(torch.cuda.current_stream().cuda_stream),
                                            C.c_void_p(int(bucket.handle.multicast_ptr)), C.c_ulonglong(n), rank, world), "k")

    def barriers_only():
        bucket.handle.barrier(channel=0, timeout_ms=20000); bucket.handle.barrier(channel=1, timeout_ms=20000)
    out["barriers_ms"] = timeit(barriers_only)
    out["kernel_ms"] = timeit(kernel_only)
    out["bytes"] = 4 * n
# the data-parallel harness with the calibrated exchange (dp.GradBucket(exchange="auto")): 14 P + 16 floats, padded
from eogs2_b200 import dp      # noqa: E402
gp = torch.Generator(device=dev).manual_seed(7)
shapes = {"xyz": 3, "f_dc": 3, "opacity": 1, "scaling": 3, "rotation": 4}
params = {k: torch.randn(1001, w, device=dev, generator=gp).requires_grad_(True) for k, w in shapes.items()}
cam = torch.zeros(4, 4, device=dev, requires_grad=True)
gb = dp.GradBucket(params, {"cam": cam}, exchange="auto")
for k, p_ in params.items():
    p_.grad = torch.full_like(p_, float(rank + 1))
cam.grad = torch.full_like(cam, float(10 * (rank + 1)))
gb.pack(); gb.all_reduce(); gb.unpack()
torch.cuda.synchronize()
tot = world * (world + 1) / 2
ok = all(bool((p_.grad == tot).all()) for p_ in params.values()) and bool((cam.grad == 10 * tot).all())
out["grad_bucket_auto"] = {"ok": ok, "exchange": gb.exchange_name, "numel": gb.flat.numel()}
assert ok, "GradBucket(exchange='auto') did not sum the gradients"
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
