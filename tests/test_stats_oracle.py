"""CPU: the restatement in oracle/stats_oracle.py against the fixture produced by the reference's own
CGaussianModel methods / train.py expressions (oracle/make_stats_golden.py), and host-side argument checks."""
import os

import numpy as np
import pytest
import torch

from oracle import stats_oracle as SO
from oracle.make_stats_golden import fresh_state, make_inputs

FIX = os.path.join(os.path.dirname(__file__), "golden", "stats_fixture.npz")


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_stats_restatement_equals_reference_methods(tag):
    fx = np.load(FIX)
    Ns, Nd, l1_accum, seed, steps = [int(v) for v in fx["%s_meta" % tag]]
    state = fresh_state(Ns, Nd)
    for it in make_inputs(Ns, Nd, seed, steps):
        SO.iteration_stats(state, Ns, it["radii"], it["grad"], it["err"] if l1_accum else None, it["timestamp"], it["densify"])
    for k in SO.ALL_NAMES:
        want = fx["%s_%s" % (tag, k)]
        got = state[k].numpy()
        assert got.shape == want.shape, k
        assert np.array_equal(got, want), k            # element-wise float32 arithmetic in the same order: bit-exact
    if l1_accum:                                       # the fixture exercises every branch
        assert (fx["%s_xyz_error_min_timestamp" % tag] >= 0).any() and (fx["%s_min_radii2D" % tag] < 1000).any()
        assert (fx["%s_error_denom" % tag] < fx["%s_denom" % tag]).any()


def test_regularizer_restatement_equals_train_py_expressions():
    fx = np.load(FIX)
    sr, mr = [float(v) for v in fx["reg_weights"]]
    terms, gd, gm = SO.regularizers(torch.from_numpy(fx["reg_disp"]), torch.from_numpy(fx["reg_motion"]), sr, mr)
    assert np.allclose(terms.numpy(), fx["reg_terms"], rtol=2e-6, atol=0)
    assert np.allclose(gd.numpy(), fx["reg_gdisp"], rtol=1e-5, atol=1e-12)
    assert np.allclose(gm.numpy(), fx["reg_gmotion"], rtol=1e-5, atol=1e-12)
    assert float(np.abs(fx["reg_gdisp"][5]).max()) == 0.0 and float(np.abs(fx["reg_gmotion"][3, 4]).max()) == 0.0
    # weights of 0 switch the terms off (train.py tests `opt.*_reg > 0`)
    t0, g0, m0 = SO.regularizers(torch.from_numpy(fx["reg_disp"]), torch.from_numpy(fx["reg_motion"]), 0.0, 0.0)
    assert float(t0.abs().sum()) == 0.0 and float(g0.abs().sum()) == 0.0 and float(m0.abs().sum()) == 0.0


def test_stats_are_cuda_only():
    from types import SimpleNamespace
    from ex4dgs_b200 import stats
    m = SimpleNamespace(_xyz=torch.zeros(4, 3))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        stats.iteration_stats(m, torch.zeros(4, dtype=torch.int32), torch.zeros(4, 3), None, 0.0)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        stats.regularizers_(torch.zeros(4, 3), torch.zeros(2, 5, 3), 1e-4, 1e-4)
