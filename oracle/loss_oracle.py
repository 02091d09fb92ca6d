"""CPU restatement of the reference's photometric loss (TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's baseline legs may import this; the product never does).

Follows utils/loss_utils.py of the reference: l1_loss (:22-25), gaussian/create_window (:33-42: the
1-D window is built in float32 from Python floats, normalised by its float32 sum, and the 2-D window
is its float32 outer product), _ssim (:56-81: five zero-padded depthwise 11x11 convolutions, C1 =
0.01^2, C2 = 0.03^2), and train.py:144-151 for how the pieces are combined.  Pinned against the
reference's own functions by tests/golden/loss_fixture.npz (oracle/make_loss_golden.py imports
utils/loss_utils.py from /root/reference and stores its outputs).

Everything is evaluated in `dtype` (float64 by default: the yardstick the CUDA kernels are compared
with; float32 reproduces the reference's own arithmetic up to the convolution's summation order).
"""
from math import exp

import torch
import torch.nn.functional as F


def window_2d(dtype=torch.float64):
    g = torch.tensor([exp(-(x - 11 // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)], dtype=torch.float32)
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).to(dtype)               # float32 outer product, like create_window


def ssim_map(img1, img2, dtype=torch.float64):
    """[3,H,W] x [3,H,W] -> [3,H,W] (loss_utils.py:56-72, reduce=False, size_average=True)."""
    x, y = img1.to(dtype).unsqueeze(0), img2.to(dtype).unsqueeze(0)
    ch = x.shape[1]
    w = window_2d(dtype).to(x.device).expand(ch, 1, 11, 11).contiguous()
    conv = lambda t: F.conv2d(t, w, padding=5, groups=ch)
    mu1, mu2 = conv(x), conv(y)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    sigma1_sq = conv(x * x) - mu1_sq
    sigma2_sq = conv(y * y) - mu2_sq
    sigma12 = conv(x * y) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return m[0]


def photometric_loss(image, gt_image, lambda_dssim=0.2, dtype=torch.float64):
    """(loss, Ll1, ssim_value, l1_errors[H,W], ssim_errors[H,W], dloss_dimage[3,H,W]) in `dtype`."""
    img = image.detach().to(dtype).clone().requires_grad_(True)
    gt = gt_image.detach().to(dtype)
    ll1 = (img - gt).abs().mean()                                   # l1_loss
    m = ssim_map(img, gt, dtype)
    ss = m.mean()
    loss = (1.0 - lambda_dssim) * ll1 + lambda_dssim * (1.0 - ss)   # train.py:146
    loss.backward()
    return (loss.detach(), ll1.detach(), ss.detach(), (img - gt).abs().mean(dim=0).detach(),
            m.mean(dim=0).detach(), img.grad.detach())
