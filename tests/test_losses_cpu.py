"""CPU: the restatement of the reference's photometric loss used by the GPU parity test (tests/test_losses_gpu.py:
ref_ssim / ref_photometric) reproduces tests/golden/loss_ref.npz, which tests/golden/make_golden_loss.py produced
by running the REFERENCE's own utils/loss_utils.py (l1_loss, ssim) and loss/shadow.py:21-29 on the CPU."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch

HERE = Path(__file__).resolve().parent
GOLDEN = HERE / "golden" / "loss_ref.npz"


def load(name):
    spec = importlib.util.spec_from_file_location(name, HERE / "golden" / f"{name}.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


MG = load("make_golden_loss")


@pytest.mark.parametrize("name", sorted(MG.CASES))
def test_restatement_matches_the_reference_functions(name):
    import test_losses_gpu as TL
    g = np.load(GOLDEN)
    img, gt, lam = MG.inputs(name)
    assert np.array_equal(img.numpy(), g[f"{name}_image"]) and np.array_equal(gt.numpy(), g[f"{name}_gt"])
    torch.set_num_threads(1)
    x = img.clone().requires_grad_(True)
    loss = TL.ref_photometric(x, gt, lam)
    loss.backward()
    want = g[f"{name}_out"]
    assert abs(float(loss.detach()) - want[0]) <= 1e-7
    assert abs(float(TL.ref_ssim(img, gt)) - want[1]) <= 1e-7
    assert np.abs(x.grad.numpy() - g[f"{name}_grad"]).max() <= 1e-9 + 1e-6 * np.abs(g[f"{name}_grad"]).max()


def test_window_matches_the_reference_gaussian():
    from eogs2_b200 import losses as L
    from math import exp
    ref = torch.Tensor([exp(-((x - 11 // 2) ** 2) / float(2 * 1.5 ** 2)) for x in range(11)])     # loss_utils.py:26-33
    ref = ref / ref.sum()
    assert torch.equal(L.gaussian_window(11, 1.5), ref)
