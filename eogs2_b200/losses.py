"""Photometric loss of EOGS++ on the sm_100a kernels.

Mirror of `l1_loss`, `ssim` (utils/loss_utils.py:18-85) and `photometric_L` (loss/shadow.py:21-29):

    L = (1 - lambda_dssim) * mean|image - gt| + lambda_dssim * (1 - ssim(image, gt))

The reference evaluates it with five depthwise `F.conv2d`, ~15 elementwise kernels and autograd;
`photometric_loss` is ONE forward and ONE backward kernel (`csrc/ssim_loss.cu`) whose gradient comes
out in the planar [C,H,W] layout the rasterizer's backward reads.  No CPU path.
"""
from __future__ import annotations

import ctypes as C
from math import exp

import torch

from . import _cabi
from .rasterizer import _f32c, _ptr

WINDOW_SIZE = 11


def gaussian_window(window_size: int = WINDOW_SIZE, sigma: float = 1.5) -> torch.Tensor:
    """utils/loss_utils.py:26-33, bit for bit: python doubles -> float32 tensor -> / sum (float32)."""
    g = torch.Tensor([exp(-((x - window_size // 2) ** 2) / float(2 * sigma ** 2)) for x in range(window_size)])
    return g / g.sum()


_WINDOW = None


def _window_c():
    global _WINDOW
    if _WINDOW is None:
        w = gaussian_window().tolist()
        _WINDOW = (C.c_float * WINDOW_SIZE)(*w)
    return _WINDOW


class _Photometric(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, lambda_dssim):
        lib = _cabi.load()
        if not image.is_cuda:
            raise _cabi.EogsRasterError("image must be a CUDA tensor: the photometric loss has no CPU path")
        dev = image.device
        if image.dim() != 3 or image.shape != gt.shape:
            raise RuntimeError("image and gt must both be [C,H,W]")
        Cn, H, W = (int(x) for x in image.shape)
        img = _f32c(image, "image", dev)
        ref = _f32c(gt, "gt_image", dev)
        with torch.cuda.device(dev):
            maps = torch.empty((3, Cn, H, W), dtype=torch.float32, device=dev)
            scal = torch.empty(5, dtype=torch.float32, device=dev)          # sums[2] | out[3]
            _cabi.check(lib.eogs_photometric_forward(
                torch.cuda.current_stream(dev).cuda_stream, Cn, H, W, _window_c(), _ptr(img), _ptr(ref),
                float(lambda_dssim), maps.data_ptr(), scal.data_ptr(), scal[2:].data_ptr()), "eogs_photometric_forward")
        ctx.save_for_backward(img, ref, maps)
        ctx.lambda_dssim = float(lambda_dssim)
        # the very tensor objects that are returned must be the ones marked: a view made afterwards would be a fresh
        # output that autograd attaches a grad_fn to (and whose incoming gradient backward() would silently drop)
        loss, ssim_mean, l1_mean = scal[2], scal[3], scal[4]
        ctx.mark_non_differentiable(ssim_mean, l1_mean)
        return loss, ssim_mean, l1_mean

    @staticmethod
    def backward(ctx, g_loss, _g_ssim, _g_l1):
        lib = _cabi.load()
        img, ref, maps = ctx.saved_tensors
        dev = img.device
        Cn, H, W = (int(x) for x in img.shape)
        with torch.cuda.device(dev):
            g = None if g_loss is None else _f32c(g_loss.reshape(1), "grad", dev)
            d_img = torch.empty_like(img)
            _cabi.check(lib.eogs_photometric_backward(
                torch.cuda.current_stream(dev).cuda_stream, Cn, H, W, _window_c(), _ptr(img), _ptr(ref),
                ctx.lambda_dssim, maps.data_ptr(), _ptr(g), d_img.data_ptr()), "eogs_photometric_backward")
        return d_img, None, None


def photometric_loss(image: torch.Tensor, gt_image: torch.Tensor, lambda_dssim: float = 0.2, return_parts: bool = False):
    """(1 - lambda) * l1_loss(image, gt) + lambda * (1 - ssim(image, gt)); differentiable w.r.t. `image`.
    return_parts=True also returns the (detached) mean SSIM and mean L1 that the training loop logs
    (train_pan.py:423,476-485)."""
    loss, ssim_mean, l1_mean = _Photometric.apply(image, gt_image, lambda_dssim)
    return (loss, ssim_mean, l1_mean) if return_parts else loss


class photometric_L(torch.nn.Module):
    """loss/shadow.py:21-29 (the Ll1 argument is accepted for signature compatibility; the fused kernel
    computes the L1 term itself from the same images)."""

    def __init__(self, lambda_dssim):
        super().__init__()
        self.lambda_dssim = lambda_dssim

    def forward(self, image, gt_image, Ll1=None):
        return photometric_loss(image, gt_image, self.lambda_dssim)


def l1_loss(network_output, gt):
    return photometric_loss(network_output, gt, 0.0)


def ssim(img1, img2):
    """Mean SSIM (utils/loss_utils.py:45-54 with size_average=True); differentiable w.r.t. img1."""
    return 1.0 - photometric_loss(img1, img2, 1.0)
