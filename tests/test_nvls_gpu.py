"""GPU (>= 2 devices on one NVSwitch domain; skipped otherwise): the NVLS all-reduce kernel (csrc/nvls.cu,
eogs2_b200/nvls.py) equals ncclAllReduce on the same data and every rank ends with the same bucket.
Runs tools/nvls_check.py under torchrun on 2 ranks."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nvls_allreduce_matches_nccl_on_two_ranks():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(ROOT / "tools" / "nvls_check.py")],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    if not out["nvls"]:
        pytest.skip("NVLS multicast not available on this box")
    assert out["max_rel_err_vs_nccl"] <= 1e-6
    assert out["p2p"] and out["p2p_max_rel_err_vs_nccl"] <= 1e-6        # the peer-to-peer kernel (the 2-GPU exchange)
    assert out["grad_bucket_auto"]["ok"]


def test_no_multicast_pointer_is_an_error_not_a_fallback(cuda_dev):
    import ctypes as C
    from eogs2_b200 import _cabi
    lib = _cabi.load()
    rc = lib.eogs_nvls_allreduce(C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(None), 1024, 0, 2)
    assert rc != 0 and b"NCCL" in lib.eogs_last_error()


def test_p2p_allreduce_rejects_bad_arguments(cuda_dev):
    """The peer-to-peer variant: no buffer table, a null peer, a misaligned peer or an impossible world are errors
    (never a silent fallback); world = 1 with the rank's own buffer is the identity."""
    import ctypes as C
    from eogs2_b200 import _cabi
    lib = _cabi.load()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    x = torch.arange(1024, dtype=torch.float32, device=cuda_dev)
    assert lib.eogs_p2p_allreduce(stream, C.c_void_p(None), 1024, 0, 2) != 0
    two = (C.c_void_p * 2)(x.data_ptr(), None)
    assert lib.eogs_p2p_allreduce(stream, two, 1024, 0, 2) != 0 and b"peer buffer 1" in lib.eogs_last_error()
    mis = (C.c_void_p * 2)(x.data_ptr(), x.data_ptr() + 4)
    assert lib.eogs_p2p_allreduce(stream, mis, 1020, 0, 2) != 0
    assert lib.eogs_p2p_allreduce(stream, two, 1024, 0, 9) != 0
    one = (C.c_void_p * 1)(x.data_ptr())
    assert lib.eogs_p2p_allreduce(stream, one, 1024, 0, 1) == 0
    torch.cuda.synchronize()
    assert torch.equal(x, torch.arange(1024, dtype=torch.float32, device=cuda_dev))
