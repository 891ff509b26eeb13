"""GPU: distCUDA2 through the C ABI (eogs_knn_dist2) is bit-identical to the CPU oracle, to the golden
fixture of the compiled reference, and — at 1 M points — to the reference itself (oracle/_ref/libknn_ref.so)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from knn_cases import GOLDEN_CASES, points
from oracle import c_oracle as O
from oracle import ref_knn

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden" / "knn_ref.npz"


def ours(p: np.ndarray) -> np.ndarray:
    from simple_knn._C import distCUDA2
    return distCUDA2(torch.from_numpy(p).cuda()).cpu().numpy()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_matches_reference_golden(name):
    g = np.load(GOLDEN)
    assert np.array_equal(bits(ours(g[f"{name}_points"])), bits(g[f"{name}_dist2"]))


@pytest.mark.parametrize("kind,P,seed", [("uniform", 6000, 1), ("clustered", 5000, 2), ("planar", 4097, 3),
                                         ("offset", 3000, 4), ("line", 2049, 5), ("uniform", 31, 6),
                                         ("uniform", 32, 7), ("uniform", 1024, 8), ("uniform", 32769, 9)])
def test_matches_oracle(kind, P, seed):
    p = points(kind, P, seed)
    assert np.array_equal(bits(ours(p)), bits(O.dist2(p)))


def test_empty_and_all_duplicates():
    from simple_knn._C import distCUDA2
    assert distCUDA2(torch.zeros(0, 3, device="cuda")).shape == (0,)
    d = distCUDA2(torch.ones(5000, 3, device="cuda"))
    assert float(d.abs().max()) == 0.0


def test_non_contiguous_input_and_side_stream():
    from simple_knn._C import distCUDA2
    p = torch.from_numpy(points("uniform", 5000, 21)).cuda()
    wide = torch.zeros(5000, 6, device="cuda")
    wide[:, ::2] = p
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        a = distCUDA2(wide[:, ::2])
    s.synchronize()
    assert torch.equal(a, distCUDA2(p))


@pytest.mark.skipif(not ref_knn.available(), reason="oracle/_ref/libknn_ref.so not built")
@pytest.mark.parametrize("kind,P", [("uniform", 1_000_000), ("clustered", 300_000), ("offset", 200_000)])
def test_full_size_bit_exact_against_the_compiled_reference(kind, P):
    from simple_knn._C import distCUDA2
    p = torch.from_numpy(points(kind, P, 1337)).cuda()
    mine = distCUDA2(p)
    ref = ref_knn.distCUDA2(p)
    assert torch.equal(mine.view(torch.int32), ref.view(torch.int32))


def test_full_size_properties():
    """Size-independent checks at 1 M points: permutation equivariance (exact) and the scale-initialisation
    use of the result (scene/gaussian_model.py:179-186) is finite."""
    from simple_knn._C import distCUDA2
    p = torch.from_numpy(points("uniform", 1_000_000, 7)).cuda()
    d = distCUDA2(p)
    perm = torch.randperm(p.shape[0], device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    assert torch.equal(distCUDA2(p[perm]), d[perm])
    scales = torch.log(torch.sqrt(torch.clamp_min(d, 1e-7)))
    assert torch.isfinite(scales).all() and float(d.min()) > 0
