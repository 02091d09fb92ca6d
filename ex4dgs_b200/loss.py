"""Fused photometric loss: the loss block that follows the rasterizer in the
reference's training step, train.py:144-151, in two kernels instead of the ~50 of
utils/loss_utils.py (l1_loss :22-25; ssim/_ssim :33-81, called twice per step when l1_accum is on).

    loss, Ll1, ssim_value, l1_errors, ssim_errors = photometric_loss(image, gt_image, lambda_dssim)

    loss        = (1 - lambda_dssim) * Ll1 + lambda_dssim * (1 - ssim_value)       train.py:146
    l1_errors   = (image - gt_image).abs().mean(dim=0)                              train.py:149
    ssim_errors = ssim(image, gt_image, reduce=False).mean(dim=0)                   train.py:150

Only `loss` carries a gradient (to `image`); Ll1 / ssim_value are the detached scalars the
reference logs, the two maps are the detached per-pixel errors its backtrack hook stacks with acc[0]
(train.py:151).  CUDA only; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _scalar_of(buf: torch.Tensor, i: int) -> torch.Tensor:
    """0-dim tensor on element i of `buf` that is NOT an autograd view of it (it shares the storage through set_()):
    the reference's training loop does `loss += ...` on the value it gets back (train.py:152-166), which autograd forbids
    on a view created inside a custom Function ("Output 0 of ... is a view and is being modified inplace")."""
    t = torch.empty(0, dtype=buf.dtype, device=buf.device)
    t.set_(buf.untyped_storage(), buf.storage_offset() + i, (), ())
    return t


class _PhotometricLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt_image, lambda_dssim):
        if not image.is_cuda:
            raise RuntimeError("ex4dgs_b200: the fused loss is CUDA-only")
        if image.dim() != 3 or image.shape[0] != 3 or gt_image.shape != image.shape:
            raise ValueError("image and gt_image must both be [3,H,W]")
        lib = _lib.load()
        dev = image.device
        img = image.detach().float().contiguous()
        gt = gt_image.detach().float().contiguous()
        _, H, W = img.shape
        scratch = torch.empty(lib.ex4dgs_loss_scratch_bytes(W, H), dtype=torch.uint8, device=dev)
        out3 = torch.empty(3, device=dev)
        l1_err = torch.empty(H, W, device=dev)
        ssim_err = torch.empty(H, W, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.ex4dgs_loss_forward(W, H, img.data_ptr(), gt.data_ptr(), float(lambda_dssim), scratch.data_ptr(),
                                         out3.data_ptr(), l1_err.data_ptr(), ssim_err.data_ptr(), C.c_void_p(stream))
        if rc < 0:
            raise RuntimeError("ex4dgs_loss_forward failed (%d): %s" % (rc, _lib.last_error()))
        ctx.save_for_backward(img, gt, scratch)
        ctx.lambda_dssim = float(lambda_dssim)
        loss, ll1, ss = _scalar_of(out3, 0), _scalar_of(out3, 1), _scalar_of(out3, 2)
        ctx.mark_non_differentiable(ll1, ss, l1_err, ssim_err)
        return loss, ll1, ss, l1_err, ssim_err

    @staticmethod
    def backward(ctx, g_loss, *_unused):
        lib = _lib.load()
        img, gt, scratch = ctx.saved_tensors
        _, H, W = img.shape
        dev = img.device
        g = g_loss.detach().float().reshape(1).contiguous()
        grad = torch.empty_like(img)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.ex4dgs_loss_backward(W, H, img.data_ptr(), gt.data_ptr(), ctx.lambda_dssim, scratch.data_ptr(),
                                          g.data_ptr(), grad.data_ptr(), C.c_void_p(stream))
        if rc < 0:
            raise RuntimeError("ex4dgs_loss_backward failed (%d): %s" % (rc, _lib.last_error()))
        return grad, None, None


class _L1Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, network_output, gt):
        if not network_output.is_cuda:
            raise RuntimeError("ex4dgs_b200: the fused loss is CUDA-only")
        if gt.shape != network_output.shape:
            raise ValueError("l1_loss: shapes differ: %s vs %s" % (tuple(network_output.shape), tuple(gt.shape)))
        lib = _lib.load()
        dev = network_output.device
        a = network_output.detach().float().contiguous()
        b = gt.detach().float().contiguous()
        out = torch.empty(1, device=dev)
        sc = _l1_scratch.get(dev)
        if sc is None:
            sc = _l1_scratch[dev] = torch.empty(int(lib.ex4dgs_l1_scratch_bytes()), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        if a.numel() == 0:
            return torch.full((), float("nan"), device=dev)            # torch: mean of an empty tensor
        with torch.cuda.device(dev):
            rc = lib.ex4dgs_l1_forward(a.numel(), a.data_ptr(), b.data_ptr(), sc.data_ptr(), out.data_ptr(), C.c_void_p(stream))
        if rc < 0:
            raise RuntimeError("ex4dgs_l1_forward failed (%d): %s" % (rc, _lib.last_error()))
        ctx.save_for_backward(a, b)
        ctx.shape = network_output.shape
        return _scalar_of(out, 0)

    @staticmethod
    def backward(ctx, g_loss):
        lib = _lib.load()
        a, b = ctx.saved_tensors
        dev = a.device
        g = g_loss.detach().float().reshape(1).contiguous()
        grad = torch.empty_like(a)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.ex4dgs_l1_backward(a.numel(), a.data_ptr(), b.data_ptr(), g.data_ptr(), grad.data_ptr(), C.c_void_p(stream))
        if rc < 0:
            raise RuntimeError("ex4dgs_l1_backward failed (%d): %s" % (rc, _lib.last_error()))
        return grad.view(ctx.shape), None


_l1_scratch = {}


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """utils/loss_utils.py:22-25 `torch.abs((network_output - gt)).mean()` - one kernel each way instead of torch's
    seven element-wise passes.  The gradient flows to `network_output` (the rendered image); `gt` is data."""
    return _L1Loss.apply(network_output, gt)


def photometric_loss(image: torch.Tensor, gt_image: torch.Tensor, lambda_dssim: float = 0.2):
    """(loss, Ll1, ssim_value, l1_errors[H,W], ssim_errors[H,W]); see the module docstring."""
    return _PhotometricLoss.apply(image, gt_image, lambda_dssim)


def backtrack_hook_tensor(acc: torch.Tensor, l1_errors: torch.Tensor, ssim_errors: torch.Tensor) -> torch.Tensor:
    """The [3,H,W] tensor train.py:151 hands to the rasterizer as the `flow` gradient."""
    return torch.stack([acc[0], l1_errors, ssim_errors])
