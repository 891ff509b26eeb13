"""Generates tests/golden/loss_ref.npz by importing the REFERENCE's own loss functions
(src/gaussiansplatting/utils/loss_utils.py: l1_loss, ssim; loss/shadow.py:21-29 photometric_L) in this
container (CPU, float32) — the reference is Python here, so it can be run as is:

    python tests/golden/make_golden_loss.py        # needs /root/reference; writes tests/golden/loss_ref.npz

The fixture pins (a) the line-by-line restatement used by the GPU parity test (tests/test_losses_gpu.py: ref_ssim /
ref_photometric; CPU test tests/test_losses_cpu.py) and (b) the fused CUDA loss itself (GPU test), forward values
and the gradient with respect to the rendered image."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference/src/gaussiansplatting/utils/loss_utils.py")

CASES = {
    # name: (C, H, W, lambda_dssim, seed)
    "rgb_small": (3, 48, 64, 0.2, 1),
    "pan_ragged": (1, 37, 29, 0.2, 2),
    "pure_dssim": (3, 24, 24, 1.0, 3),
    "five_channels_l1": (5, 20, 50, 0.0, 4),
}


def inputs(name):
    C, H, W, lam, seed = CASES[name]
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(C, H, W, generator=g)
    img = (gt + 0.15 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    img[:, : H // 4] = gt[:, : H // 4]                      # a region with image == gt
    return img, gt, lam


def main():
    spec = importlib.util.spec_from_file_location("ref_loss_utils", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    torch.set_num_threads(1)
    save = {}
    for name in CASES:
        img, gt, lam = inputs(name)
        x = img.clone().requires_grad_(True)
        l1 = m.l1_loss(x, gt)
        s = m.ssim(x, gt)
        loss = (1.0 - lam) * l1 + lam * (1.0 - s)              # photometric_L.forward, loss/shadow.py:26-29
        loss.backward()
        save[f"{name}_image"], save[f"{name}_gt"] = img.numpy(), gt.numpy()
        save[f"{name}_out"] = np.array([float(loss.detach()), float(s.detach()), float(l1.detach())], np.float64)
        save[f"{name}_grad"] = x.grad.numpy()
        print(name, float(loss.detach()), float(s.detach()), float(l1.detach()))
    np.savez_compressed(ROOT / "tests" / "golden" / "loss_ref.npz", **save)


if __name__ == "__main__":
    main()
