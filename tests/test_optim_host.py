"""CPU tests of row N4 (ex4dgs_b200/optim.py): the per-step scalars the fused RAdam kernel receives are pinned
against torch/optim/radam.py's own expressions, the optimizer object keeps torch.optim.RAdam's state layout
(so checkpoints interchange), the product refuses CPU tensors, and the data-parallel gradient exchange is
exercised with a world_size-2 gloo group."""
import ctypes as C
import math
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ex4dgs_b200 import optim as fopt


def _torch_scalars(lr, step, beta1, beta2):
    """The non-capturable branch of torch.optim.radam._multi_tensor_radam, verbatim expressions."""
    rho_inf = 2 / (1 - beta2) - 1
    rho_t = rho_inf - 2 * step * (beta2 ** step) / (1 - beta2 ** step)
    rect = ((rho_t - 4) * (rho_t - 2) * rho_inf / ((rho_inf - 4) * (rho_inf - 2) * rho_t)) ** 0.5 if rho_t > 5 else 0
    unrectified = 0 if rect > 0 else 1.0
    bc1 = 1 - beta1 ** step
    unrect_step_size = (lr * unrectified / bc1) * -1
    bias_correction2 = ((1 - beta2 ** step) ** 0.5) * (lr * rect / bc1) * -1
    return bias_correction2, unrect_step_size, rect > 0


def test_step_scalars_match_torch_expressions(built):
    from ex4dgs_b200 import _lib
    lib = _lib.load()
    S, U, R = C.c_float(), C.c_float(), C.c_int()
    for beta1, beta2 in ((0.9, 0.999), (0.8, 0.99)):
        for lr in (1.6e-4, 0.0025, 0.05):
            for step in list(range(1, 40)) + [100, 1000, 7000, 30000, 120000]:
                assert lib.ex4dgs_radam_scalars(lr, step, beta1, beta2, C.byref(S), C.byref(U), C.byref(R)) == 0
                s, u, r = _torch_scalars(lr, float(step), beta1, beta2)
                assert bool(R.value) == bool(r)
                assert S.value == torch.tensor(s, dtype=torch.float32).item(), (step, S.value, s)
                assert U.value == torch.tensor(u, dtype=torch.float32).item(), (step, U.value, u)
    # the first five steps of the default betas are the un-rectified (SGD-with-momentum) phase
    assert [_torch_scalars(1e-3, float(t), 0.9, 0.999)[2] for t in range(1, 8)] == [False] * 5 + [True] * 2
    assert lib.ex4dgs_radam_scalars(1e-3, 0, 0.9, 0.999, C.byref(S), C.byref(U), C.byref(R)) < 0


def test_state_layout_is_torch_radam_compatible():
    ps = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]
    groups = [{"params": [ps[0]], "lr": 1.6e-4, "name": "xyz"}, {"params": [ps[1]], "lr": 0.05, "name": "opacity"}]
    ref = torch.optim.RAdam([dict(g) for g in groups], lr=0.001)
    ours = fopt.FusedRAdam([dict(g) for g in groups], lr=0.001)
    for k in ("lr", "betas", "eps", "weight_decay", "name"):
        assert [g[k] for g in ours.param_groups] == [g[k] for g in ref.param_groups]
    # a state dict produced by torch's RAdam loads into the fused one and back
    for p in ps:
        p.grad = torch.randn_like(p)
    ref.step()
    sd = ref.state_dict()
    ours.load_state_dict(sd)
    st = ours.state[ps[0]]
    assert set(st) == {"step", "exp_avg", "exp_avg_sq"} and float(st["step"]) == 1.0
    ref.load_state_dict(ours.state_dict())
    with pytest.raises(ValueError):
        fopt.FusedRAdam(ps, weight_decay=0.1)
    with pytest.raises(ValueError):
        fopt.FusedRAdam(ps, betas=(1.0, 0.999))


def test_no_cpu_fallback(built):
    p = torch.nn.Parameter(torch.randn(4))
    p.grad = torch.ones(4)
    opt = fopt.FusedRAdam([p])
    with pytest.raises(RuntimeError):
        opt.step()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    ps = [torch.nn.Parameter(torch.zeros(6, 3)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2))]
    ps[0].grad = torch.randn(6, 3, generator=g)
    ps[1].grad = torch.randn(5, generator=g)          # ps[2] has no gradient on any rank
    scale = fopt.allreduce_gradients(ps)
    q.put((rank, scale, [None if p.grad is None else p.grad.clone() for p in ps]))
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    assert fopt.allreduce_gradients([torch.nn.Parameter(torch.zeros(2))]) == 1.0     # not distributed: identity
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp0 = sum(torch.randn(6, 3, generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
    for rank, scale, grads in res:
        assert scale == 0.5
        assert torch.allclose(grads[0], exp0) and grads[2] is None
    assert torch.equal(res[0][2][1], res[1][2][1])
