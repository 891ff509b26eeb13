"""CPU, world_size 2, gloo: the host logic of the gradient exchange selector (eogs2_b200/nvls.py).  Without NVLS
multicast (CPU tensors here) every rank must agree on the NCCL/gloo path and the returned all-reduce must sum the
bucket in place; a single rank gets a no-op."""
import os
import socket
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo")
    from eogs2_b200.nvls import SymmetricBucket, make_grad_exchange
    assert SymmetricBucket.create(1024, torch.device("cpu")) is None            # agreed on by both ranks (MIN all-reduce)
    flat, exchange, name = make_grad_exchange(1024, torch.device("cpu"))
    flat.copy_(torch.arange(1024, dtype=torch.float32) * (rank + 1))
    exchange()
    torch.save({"flat": flat.clone(), "name": name}, out + f".{rank}")
    dist.destroy_process_group()


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_exchange_falls_back_consistently_without_multicast(tmp_path):
    out = str(tmp_path / "res")
    mp.spawn(worker, args=(2, free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    want = torch.arange(1024, dtype=torch.float32) * 3
    assert torch.equal(r0["flat"], want) and torch.equal(r1["flat"], want)
    assert r0["name"] == r1["name"] and "ncclAllReduce" in r0["name"] and "unavailable" in r0["name"]


def test_single_rank_exchange_is_a_no_op():
    from eogs2_b200.nvls import make_grad_exchange
    flat, exchange, name = make_grad_exchange(64, torch.device("cpu"))
    flat.fill_(2.0)
    exchange()
    assert torch.equal(flat, torch.full((64,), 2.0)) and "single rank" in name


def test_bucket_must_be_a_multiple_of_four_floats():
    import pytest
    from eogs2_b200.nvls import SymmetricBucket
    # the check only matters on the multi-rank path; without a process group create() returns None first
    assert SymmetricBucket.create(6, torch.device("cpu")) is None
    with pytest.raises(Exception):
        import ctypes as C
        from eogs2_b200 import _cabi
        _cabi.check(_cabi.load().eogs_nvls_allreduce(None, C.c_void_p(16), 6, 0, 2), "eogs_nvls_allreduce")
