"""CPU: the loss oracle (oracle/loss_oracle.py) against the outputs of the reference's own
utils/loss_utils.py stored in tests/golden/loss_fixture.npz (oracle/make_loss_golden.py)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loss_oracle  # noqa: E402

FIX = np.load(os.path.join(ROOT, "tests", "golden", "loss_fixture.npz"))
CASES = sorted({k.split("_")[0] for k in FIX.files})


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_loss_oracle_matches_reference_outputs(case, dtype):
    img, gt = torch.from_numpy(FIX[case + "_img"]), torch.from_numpy(FIX[case + "_gt"])
    lam = float(FIX[case + "_lambda"])
    loss, ll1, ss, l1e, sse, grad = loss_oracle.photometric_loss(img, gt, lam, dtype)
    # float32 = the reference's own arithmetic (measured: identical on CPU); float64 differs by the
    # float32 cancellation noise of sigma = E[x^2] - mu^2 in the reference (measured <= 2.7e-5 on the map)
    f32 = dtype == torch.float32
    t_map, t_grad = (1e-6, 1e-5) if f32 else (1e-4, 1e-4)
    assert abs(float(loss) - float(FIX[case + "_loss"])) < (1e-7 if f32 else 1e-6)
    assert abs(float(ll1) - float(FIX[case + "_Ll1"])) < 1e-6
    np.testing.assert_allclose(l1e.numpy(), FIX[case + "_l1_errors"], atol=1e-6)
    m = loss_oracle.ssim_map(img, gt, dtype).numpy()
    np.testing.assert_allclose(m, FIX[case + "_ssim_map"], atol=t_map)
    np.testing.assert_allclose(sse.numpy(), FIX[case + "_ssim_map"].mean(0), atol=t_map)
    g_ref = FIX[case + "_grad"]
    scale = np.abs(g_ref).max()
    assert np.abs(grad.numpy() - g_ref).max() < t_grad * scale


def test_window_is_the_reference_window():
    w = loss_oracle.window_2d(torch.float32)
    assert w.shape == (11, 11) and abs(float(w.sum()) - 1.0) < 1e-6
    assert torch.equal(w, w.t())
