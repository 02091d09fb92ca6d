"""Diagnostic (not a test): CUDA-event time of the fused loss vs the torch ops of the reference's loss block."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ex4dgs_b200.loss import photometric_loss
from oracle import loss_oracle

H, W = 1014, 1352
gt = torch.rand(3, H, W, device="cuda")
img0 = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1)


def fused():
    x = img0.clone().requires_grad_(True)
    loss, _, _, l1e, sse = photometric_loss(x, gt, 0.2)
    loss.backward()


def torch_ref():
    x = img0.clone().requires_grad_(True)
    ll1 = (x - gt).abs().mean()
    loss = 0.8 * ll1 + 0.2 * (1.0 - loss_oracle.ssim_map(x, gt, torch.float32).mean())
    l1e = (x - gt).abs().mean(dim=0)
    sse = loss_oracle.ssim_map(x.detach(), gt, torch.float32).mean(dim=0)
    loss.backward()


for name, fn in (("fused", fused), ("torch", torch_ref)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(name, "ms per fwd+bwd:", e0.elapsed_time(e1) / 50)
