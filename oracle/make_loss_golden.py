"""Generates tests/golden/loss_fixture.npz by running the reference's OWN loss functions
(/root/reference/utils/loss_utils.py, imported unmodified; CPU float32) on small seeded images.
Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_loss_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
from utils.loss_utils import l1_loss, ssim  # noqa: E402

CASES = {"a": (37, 53, 0.2), "b": (16, 32, 0.2), "c": (7, 9, 0.5), "d": (64, 80, 0.0)}


def main():
    out = {}
    g = torch.Generator().manual_seed(20240925)
    for name, (H, W, lam) in CASES.items():
        gt = torch.rand(3, H, W, generator=g)
        # rendered image = blurred/noisy version of the target, plus flat regions (sigma ~ 0) and exact ties
        img = (gt + 0.15 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
        img[:, : H // 3, : W // 4] = 0.25
        gt[:, : H // 4, : W // 3] = 0.25
        img = img.clone().requires_grad_(True)
        ll1 = l1_loss(img, gt)
        loss = (1.0 - lam) * ll1 + lam * (1.0 - ssim(img, gt))
        loss.backward()
        out[name + "_img"] = img.detach().numpy()
        out[name + "_gt"] = gt.numpy()
        out[name + "_lambda"] = np.float64(lam)
        out[name + "_loss"] = loss.detach().numpy()
        out[name + "_Ll1"] = ll1.detach().numpy()
        out[name + "_ssim_map"] = ssim(img.detach(), gt, reduce=False).numpy()
        out[name + "_l1_errors"] = (img.detach() - gt).abs().mean(dim=0).numpy()
        out[name + "_grad"] = img.grad.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "loss_fixture.npz"), **out)
    print("wrote loss_fixture.npz:", {k: v.shape for k, v in out.items() if k.endswith("_img")})


if __name__ == "__main__":
    main()
