"""GPU: the depth-order stage alone (csrc/binning.cu: depth_order_impl) through the C-ABI test hook
eogs_debug_depth_order — the order of the P Gaussians by (depth bit pattern, id) that the tile lists inherit (the depth
half of the reference's 64-bit instance sort, DGR/cuda_rasterizer/rasterizer_impl.cu:306-311).  Oracle: a stable sort
of the same keys (numpy argsort(kind="stable") on int64); bit-for-bit agreement, ties and culled sentinels included."""
import ctypes as C

import numpy as np
import pytest
import torch

from eogs2_b200 import _cabi

pytestmark = pytest.mark.gpu
SENTINEL = 0xFFFFFFFF


def depth_order(dev, keys_u32: np.ndarray) -> np.ndarray:
    lib = _cabi.load()
    P = int(keys_u32.size)
    k = torch.from_numpy(keys_u32.view(np.int32).copy()).to(dev)
    order = torch.full((max(P, 1),), -1, dtype=torch.int32, device=dev)
    scratch = torch.empty(lib.eogs_debug_depth_order_bytes(P), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(lib.eogs_debug_depth_order(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), P,
                                               C.c_void_p(k.data_ptr()), C.c_void_p(order.data_ptr()),
                                               C.c_void_p(scratch.data_ptr())), "eogs_debug_depth_order")
        torch.cuda.synchronize(dev)
    return order[:P].cpu().numpy().astype(np.int64)


def oracle(keys_u32: np.ndarray) -> np.ndarray:
    return np.argsort(keys_u32.astype(np.int64), kind="stable")


def make_keys(kind: str, P: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "depths":                     # positive floats of one scene: 200 - altitude
        k = (200.0 - rng.uniform(-30.0, 75.0, P)).astype(np.float32).view(np.uint32)
    elif kind == "depths_culled":            # the same with culled Gaussians (sentinel key)
        k = (200.0 - rng.uniform(-30.0, 75.0, P)).astype(np.float32).view(np.uint32)
        k[rng.random(P) < 0.3] = SENTINEL
    elif kind == "wide":                     # depths over many binades
        k = np.exp(rng.uniform(-20, 20, P)).astype(np.float32).view(np.uint32)
    elif kind == "few":                      # massive ties: stability decides
        k = rng.choice(np.array([0x3F800000, 0x3F800001, 0x40000000, SENTINEL], np.uint32), P)
    elif kind == "constant":
        k = np.full(P, 0x42C80000, np.uint32)
    elif kind == "constant_culled":          # one depth + culled ones: only the sentinel flag separates them
        k = np.full(P, 0x42C80000, np.uint32)
        k[rng.random(P) < 0.5] = SENTINEL
    elif kind == "all_culled":
        k = np.full(P, SENTINEL, np.uint32)
    elif kind == "random32":                 # every bit varies, sign bit included: the plain 32-bit path
        k = rng.integers(0, 2 ** 32, P, dtype=np.uint64).astype(np.uint32)
    elif kind == "top_ones":                 # a valid key whose varying bits are all ones next to sentinels
        k = rng.integers(0x3F800000, 0x3F800000 + 512, P, dtype=np.uint64).astype(np.uint32)
        k[::7] = 0x3F8001FF
        k[3::5] = SENTINEL
    else:
        raise ValueError(kind)
    return k


SIZES = [1, 2, 31, 32, 33, 4095, 4096, 4097, 40_000, 65_536, 65_537, 300_000, 1_000_003]
KINDS = ["depths", "depths_culled", "wide", "few", "constant", "constant_culled", "all_culled", "random32", "top_ones"]


@pytest.mark.parametrize("kind", KINDS)
def test_depth_order_equals_a_stable_sort(cuda_dev, kind):
    for n, P in enumerate(SIZES):
        keys = make_keys(kind, P, 100 + n)
        got = depth_order(cuda_dev, keys)
        want = oracle(keys)
        assert np.array_equal(got, want), (kind, P, int(np.argmax(got != want)))


def test_depth_order_is_deterministic_across_calls(cuda_dev):
    keys = make_keys("depths_culled", 777_777, 5)
    a = depth_order(cuda_dev, keys)
    for _ in range(3):
        assert np.array_equal(depth_order(cuda_dev, keys), a)
