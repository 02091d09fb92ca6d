"""GPU: the fused photometric loss (ex4dgs_b200/loss.py -> ex4dgs_loss_forward/backward) against the
float64 loss oracle and against the reference outputs stored in tests/golden/loss_fixture.npz.

Tolerances (floating point; stated here as the brief asks): the SSIM map is ill-conditioned in float32
where both images are flat (sigma = E[x^2] - mu^2 cancels against C2 = 9e-4); the reference's own
float32 result deviates from exact arithmetic by up to ~3e-5 on the map, so the map is held to 1e-4
absolute, the scalar loss to 2e-6 absolute, the gradient to 1e-3 of its largest entry (the north-star
gradient tolerance), and the L1 error map to 1e-6."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loss_oracle  # noqa: E402
from oracle import getters_oracle as GO  # noqa: E402

pytestmark = pytest.mark.gpu
FIX = np.load(os.path.join(ROOT, "tests", "golden", "loss_fixture.npz"))
CASES = sorted({k.split("_")[0] for k in FIX.files})


def run_cuda(img, gt, lam, g_loss=None):
    from ex4dgs_b200.loss import photometric_loss
    x = img.cuda().clone().requires_grad_(True)
    loss, ll1, ss, l1e, sse = photometric_loss(x, gt.cuda(), lam)
    assert not ll1.requires_grad and not l1e.requires_grad and not sse.requires_grad
    if g_loss is None:
        loss.backward()
    else:
        (loss * g_loss).backward()
    return loss.detach().cpu(), ll1.cpu(), ss.cpu(), l1e.cpu(), sse.cpu(), x.grad.cpu()


def check(img, gt, lam, ours, g_scale=1.0):
    loss, ll1, ss, l1e, sse, grad = loss_oracle.photometric_loss(img, gt, lam, torch.float64)
    assert abs(float(ours[0]) - float(loss)) < 2e-6
    assert abs(float(ours[1]) - float(ll1)) < 1e-6
    assert abs(float(ours[2]) - float(ss)) < 5e-6
    assert (ours[3].double() - l1e).abs().max() < 1e-6
    assert (ours[4].double() - sse).abs().max() < 1e-4
    scale = grad.abs().max() * g_scale
    assert (ours[5].double() - grad * g_scale).abs().max() <= 1e-3 * scale


@pytest.mark.parametrize("case", CASES)
def test_loss_matches_reference_fixture(built, case):
    img, gt = torch.from_numpy(FIX[case + "_img"]), torch.from_numpy(FIX[case + "_gt"])
    lam = float(FIX[case + "_lambda"])
    ours = run_cuda(img, gt, lam)
    assert abs(float(ours[0]) - float(FIX[case + "_loss"])) < 2e-6
    np.testing.assert_allclose(ours[3].numpy(), FIX[case + "_l1_errors"], atol=1e-6)
    np.testing.assert_allclose(ours[4].numpy(), FIX[case + "_ssim_map"].mean(0), atol=1e-4)
    g = FIX[case + "_grad"]
    assert np.abs(ours[5].numpy() - g).max() <= 1e-3 * np.abs(g).max()
    check(img, gt, lam, ours)


@pytest.mark.parametrize("H,W", [(1, 1), (5, 3), (16, 32), (17, 33), (31, 100), (129, 67)])
def test_loss_ragged_sizes(built, H, W):
    g = torch.Generator().manual_seed(H * 1000 + W)
    gt = torch.rand(3, H, W, generator=g)
    img = (gt + 0.1 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    check(img, gt, 0.2, run_cuda(img, gt, 0.2))


def test_loss_identical_images_and_upstream_gradient(built):
    g = torch.Generator().manual_seed(7)
    gt = torch.rand(3, 40, 48, generator=g)
    ours = run_cuda(gt, gt, 0.2)
    assert float(ours[0]) < 1e-6 and abs(float(ours[2]) - 1.0) < 1e-6       # ssim(x, x) = 1, L1 = 0
    assert ours[5].abs().max() < 1e-6                                       # stationary point; sign(0) = 0
    img = (gt + 0.2 * torch.randn(3, 40, 48, generator=g)).clamp(0, 1)
    check(img, gt, 0.35, run_cuda(img, gt, 0.35, g_loss=2.5), g_scale=2.5)


def test_loss_full_size_and_determinism(built):
    H, W = 1014, 1352
    g = torch.Generator().manual_seed(11)
    gt = torch.rand(3, H, W, generator=g)
    img = (gt + 0.1 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    a = run_cuda(img, gt, 0.2)
    b = run_cuda(img, gt, 0.2)
    for u, v in zip(a, b):
        assert torch.equal(u, v)                                            # fixed-order sums
    check(img, gt, 0.2, a)


def test_loss_drives_the_rasterizer_backward(built):
    """train.py:139-172 shape: render -> loss (+ hook tensor as the flow gradient) -> backward."""
    import ex4dgs_b200 as m
    from ex4dgs_b200 import synth
    from ex4dgs_b200.loss import photometric_loss, backtrack_hook_tensor
    sc = synth.make_config("tiny")
    fi = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in GO.flat_inputs(sc, 3.0).items()}
    cam = sc.cam
    rs = m.GaussianRasterizationSettings(
        image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, kernel_size=cam.kernel_size,
        subpixel_offset=torch.zeros(cam.H, cam.W, 2, device="cuda"), bg=sc.bg.cuda(), scale_modifier=1.0,
        viewmatrix=cam.viewmatrix.cuda(), projmatrix=cam.projmatrix.cuda(), sh_degree=sc.sh_degree,
        campos=cam.campos.cuda(), prefiltered=False, min_depth=cam.min_depth, max_depth=cam.max_depth, debug=False)
    means = fi["means3D"].clone().requires_grad_(True)
    opac = fi["opacities"].clone().requires_grad_(True)
    out = m.GaussianRasterizer(rs)(means3D=means, means2D=torch.zeros_like(means), dir3D=torch.zeros_like(means),
                                   opacities=opac, shs=fi["shs"], scales=fi["scales"], rotations=fi["rotations"])
    color, flow, acc = out[0], out[3], out[4]
    gt = torch.rand(3, cam.H, cam.W, generator=torch.Generator().manual_seed(3)).cuda()
    loss, _, _, l1e, sse = photometric_loss(color, gt, 0.2)
    hook = backtrack_hook_tensor(acc, l1e, sse)
    flow.register_hook(lambda grad: hook)
    (loss + flow.mean() * 0).backward()
    assert torch.isfinite(means.grad).all() and means.grad.abs().sum() > 0
    assert torch.isfinite(opac.grad).all()


@pytest.mark.parametrize("shape", [(3, 1014, 1352), (3, 37, 53), (7,), (1, 4099)])
def test_l1_loss_equals_torch(shape):
    """ex4dgs_b200.loss.l1_loss == utils/loss_utils.py:22-25 `torch.abs((network_output - gt)).mean()`: value to float
    rounding (double accumulation here), gradient bit for bit (sgn(a - b) * g / n, sgn(0) = 0), unaligned views, an
    upstream gradient other than 1, bit-reproducible."""
    from ex4dgs_b200.loss import l1_loss
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(sum(shape))
    n = int(np.prod(shape))
    base_a = torch.rand(n + 1, generator=g).to(dev)
    base_b = torch.rand(n + 1, generator=g).to(dev)
    base_b[1 + 5 % n] = base_a[1 + 5 % n]                          # an exact zero difference
    for off in (0, 1):                                             # off = 1: views at +4 bytes -> the scalar path
        b = base_b.clone()[off:off + n].view(shape)
        a1 = base_a.clone()[off:off + n].view(shape).requires_grad_(True)
        a2 = base_a.clone()[off:off + n].view(shape).requires_grad_(True)
        assert (a2.data_ptr() % 16 == 0) == (off == 0)
        ref = torch.abs((a1 - b)).mean()
        (ref * 0.37).backward()
        ours = l1_loss(a2, b)
        (ours * 0.37).backward()
        assert abs(float(ours.detach()) - float(ref.detach())) <= 2e-6 * abs(float(ref.detach()))
        assert torch.equal(a2.grad, a1.grad)
        assert torch.equal(l1_loss(a2.detach(), b), ours.detach())
    with pytest.raises(RuntimeError, match="CUDA-only"):
        l1_loss(torch.zeros(3), torch.zeros(3))


def test_loss_value_can_be_modified_in_place(built):
    """train.py:152-166 does `loss += ...` on the value the loss returns: the fused losses must not hand out autograd
    views (torch: "Output 0 of ... is a view and is being modified inplace")."""
    from ex4dgs_b200.loss import photometric_loss, l1_loss
    g = torch.Generator().manual_seed(2)
    img = torch.rand(3, 40, 56, generator=g).cuda().requires_grad_(True)
    gt = torch.rand(3, 40, 56, generator=g).cuda()
    extra = torch.ones((), device="cuda", requires_grad=True)
    loss, ll1, ss, _, _ = photometric_loss(img, gt, 0.2)
    want = float(loss)
    loss += 0.25 * extra
    loss += 0.0 * ll1
    loss.backward()
    assert abs(float(loss) - (want + 0.25)) < 1e-6 and img.grad is not None and float(extra.grad) == 0.25
    g1 = img.grad.clone()
    img.grad = None
    l1 = l1_loss(img, gt)
    l1 += 1.0
    l1.backward()
    assert img.grad is not None and float(img.grad.abs().max()) > 0 and not torch.equal(g1, img.grad)
