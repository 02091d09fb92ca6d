"""GPU parity of the fused RAdam step (row N4) against the real thing: torch.optim.RAdam on the same
device (its CUDA "foreach" implementation is what the reference's train.py:250 executes), over the
reference's 15-group layout with per-group learning rates, odd sizes, the un-rectified first five
steps, a group without gradient, unaligned views and a learning-rate change between steps."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from ex4dgs_b200 import optim as fopt  # noqa: E402

# (name, shape, lr) after scene/c_gaussian_model.py:430-449 with the N3V config's learning rates
GROUPS = [("xyz", (1237, 3), 1.6e-4), ("f_dc", (1237, 1, 3), 0.0025), ("f_rest", (1237, 15, 3), 0.0025 / 20),
          ("opacity", (1237, 1), 0.05), ("scaling", (1237, 3), 0.005), ("rotation", (1237, 4), 0.001),
          ("xyz_disp", (1237, 3), 1.6e-4), ("motion_xyz", (411, 36, 3), 1.6e-4), ("motion_f_dc", (411, 1, 3), 0.0025),
          ("motion_f_rest", (411, 15, 3), 0.0025 / 20), ("motion_scaling", (411, 3), 0.005), ("motion_opacity", (411, 1), 0.05),
          ("motion_opacity_center", (411, 2, 1), 1e-4), ("motion_opacity_var", (411, 2, 1), 1e-4),
          ("motion_rotation", (411, 36, 4), 0.001)]


def _make(dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [(n, torch.randn(*s, generator=g).to(dev), lr) for n, s, lr in GROUPS]


def _ulps(a, b):
    ai = a.contiguous().view(torch.int32).long()
    bi = b.contiguous().view(torch.int32).long()
    return (ai - bi).abs().max().item()


def test_fused_radam_matches_torch_radam():
    dev = torch.device("cuda:0")
    init = _make(dev)
    pr = [torch.nn.Parameter(t.clone()) for _, t, _ in init]
    po = [torch.nn.Parameter(t.clone()) for _, t, _ in init]
    ref = torch.optim.RAdam([{"params": [p], "lr": lr, "name": n} for p, (n, _, lr) in zip(pr, init)], lr=0.001)
    ours = fopt.FusedRAdam([{"params": [p], "lr": lr, "name": n} for p, (n, _, lr) in zip(po, init)], lr=0.001)
    g = torch.Generator().manual_seed(7)
    worst = 0
    for step in range(1, 13):
        for i, (a, b) in enumerate(zip(pr, po)):
            if i == 13 and step % 2:            # a group that sometimes has no gradient (optimizer skips it)
                a.grad = b.grad = None
                continue
            gr = (torch.randn(a.shape, generator=g) * (10.0 ** torch.randint(-6, 1, (1,), generator=g).item())).to(dev)
            if i == 7:
                gr[:, :30] = 0                   # keyframe gradients are dense but mostly zero
            a.grad, b.grad = gr.clone(), gr.clone()
        if step == 8:                            # update_learning_rate rewrites group['lr'] (c_gaussian_model.py:461-470)
            for grp in (ref.param_groups[0], ours.param_groups[0]):
                grp["lr"] = 9.7e-5
        ref.step()
        ours.step()
        for (n, _, _), a, b in zip(init, pr, po):
            # same operations in the same order: equal up to the last bit or two of float rounding
            assert torch.allclose(a, b, rtol=2e-6, atol=5e-8), (step, n, (a - b).abs().max().item())
            sa, sb = ref.state[a], ours.state[b]
            if sa:
                assert torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=1e-6, atol=1e-12), (step, n)
                assert torch.allclose(sa["exp_avg_sq"], sb["exp_avg_sq"], rtol=1e-6, atol=1e-20), (step, n)
                assert float(sa["step"]) == float(sb["step"])
            worst = max(worst, _ulps(a.detach(), b.detach()))
    print("max parameter difference after 12 steps: %d ulp" % worst)     # informational (parameters near 0 inflate ulps)


def test_fused_radam_grad_scale_views_and_edge_sizes():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    for numel in (1, 3, 4095, 4096, 4097, 70001):
        base = torch.randn(numel + 1, generator=g).to(dev)
        a = torch.nn.Parameter(base[1:].clone())          # fresh allocation: 16-byte aligned
        storage = base.clone()
        b = torch.nn.Parameter(storage[1:])               # view at +4 bytes: the scalar (unaligned) path
        gr = torch.randn(numel, generator=g).to(dev)
        ref = torch.optim.RAdam([a], lr=0.01)
        ours = fopt.FusedRAdam([b], lr=0.01)
        for _ in range(7):
            a.grad = gr / 4
            b.grad = gr.clone()
            ref.step()
            ours.step(grad_scale=0.25)                    # mean over 4 ranks folded into the kernel
        assert torch.allclose(a, b, rtol=2e-6, atol=5e-8), (numel, (a - b).abs().max().item())
    # empty tensors and gradient-less groups are no-ops
    e = torch.nn.Parameter(torch.zeros(0, 3, device=dev))
    e.grad = torch.zeros(0, 3, device=dev)
    fopt.FusedRAdam([e]).step()


def test_fused_radam_nan_guards():
    """train.py:244-253 folded into the step: nan_to_num on a named group's gradient == torch's nan_to_num followed
    by RAdam; a NaN written into a `check_nan` group raises its flag (and only its flag); groups outside the two
    name lists behave exactly as before."""
    dev = torch.device("cuda:0")
    init = _make(dev, seed=5)
    pr = [torch.nn.Parameter(t.clone()) for _, t, _ in init]
    po = [torch.nn.Parameter(t.clone()) for _, t, _ in init]
    ref = torch.optim.RAdam([{"params": [p], "lr": lr, "name": n} for p, (n, _, lr) in zip(pr, init)], lr=0.001)
    ours = fopt.FusedRAdam([{"params": [p], "lr": lr, "name": n} for p, (n, _, lr) in zip(po, init)], lr=0.001,
                           check_nan=("xyz", "motion_xyz"), sanitize_grad=("motion_opacity_var",))
    names = [n for n, _, _ in init]
    iv = names.index("motion_opacity_var")
    g = torch.Generator().manual_seed(11)
    for step in range(1, 9):
        for i, (a, b) in enumerate(zip(pr, po)):
            gr = (torch.randn(a.shape, generator=g) * 1e-2).to(dev)
            if i == iv:
                gr.view(-1)[3] = float("nan")
                gr.view(-1)[5] = float("inf")
                gr.view(-1)[9] = float("-inf")
                a.grad, b.grad = gr.nan_to_num(), gr.clone()      # the reference sanitises before optimizer.step()
            else:
                a.grad, b.grad = gr.clone(), gr.clone()
        ref.step()
        ours.step()
        for n, a, b in zip(names, pr, po):
            assert torch.allclose(a, b, rtol=2e-6, atol=5e-8, equal_nan=True), (step, n)
        assert ours.nan_detected() == {"xyz": False, "motion_xyz": False}
        assert not torch.isnan(po[iv]).any()
    ours.poll_nan()
    # a NaN gradient in one keyframe of one dynamic Gaussian -> NaN parameter -> flag of that group only
    for a, b in zip(pr, po):
        b.grad = torch.zeros_like(b)
    po[names.index("motion_xyz")].grad[17, 4, 1] = float("nan")
    po[names.index("f_rest")].grad[0, 0, 0] = float("nan")       # not a checked group
    ours.step()
    torch.cuda.synchronize()
    assert ours.poll_nan() == {"xyz": False, "motion_xyz": False}     # first poll after the step: copy only queued
    torch.cuda.synchronize()
    assert ours.poll_nan() == {"xyz": False, "motion_xyz": True}
    assert ours.nan_detected() == {"xyz": False, "motion_xyz": True}
    assert torch.isnan(po[names.index("motion_xyz")][17, 4, 1])
    ours.clear_nan()
    assert ours.nan_detected() == {"xyz": False, "motion_xyz": False}


def test_overlapped_allreduce_step_equals_plain_step(built):
    """FusedRAdam.step(allreduce_group=...) - reductions queued up front, the step kernel launched per bucket - on a
    one-rank NCCL group must give exactly the plain step (sum over one rank, mean = the gradient itself)."""
    import torch.distributed as dist
    from ex4dgs_b200.optim import FusedRAdam
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29577", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
        created = True
    try:
        g = torch.Generator().manual_seed(9)
        shapes = [(1000, 3), (1000, 1, 3), (1000, 15, 3), (1000, 1), (700, 36, 3), (700, 36, 4), (5,), (700, 2)]
        def make():
            ps = [torch.nn.Parameter(torch.randn(*s, generator=torch.Generator().manual_seed(i)).cuda()) for i, s in enumerate(shapes)]
            return ps, FusedRAdam([{"params": [p], "lr": 1e-2 * (i + 1), "name": "g%d" % i} for i, p in enumerate(ps)])
        pa, oa = make()
        pb, ob = make()
        for it in range(7):
            for x, y in zip(pa, pb):
                gr = torch.randn(x.shape, generator=g).cuda()
                x.grad, y.grad = gr.clone(), gr.clone()
            oa.step()
            ob.step(allreduce_group=dist.group.WORLD, buckets=3)
        torch.cuda.synchronize()
        for x, y in zip(pa, pb):
            assert torch.equal(x, y)
            assert torch.equal(oa.state[x]["exp_avg_sq"], ob.state[y]["exp_avg_sq"])
    finally:
        if created:
            dist.destroy_process_group()
