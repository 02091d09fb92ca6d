"""Fused optimizer step.

`FusedRAdam` is a drop-in for the optimizer the reference builds in CGaussianModel.training_setup
(`torch.optim.RAdam(l, lr=0.001)` over 15 named single-tensor groups, scene/c_gaussian_model.py:430-449)
and steps once per iteration (train.py:250-251): same constructor arguments, same `param_groups`
(the reference rewrites `group['lr']` per iteration, :461-470, and swaps `group['params'][0]` when it
densifies/prunes, :674-760), same `state[p] = {'step', 'exp_avg', 'exp_avg_sq'}` layout and therefore
the same `state_dict()` - a checkpoint written with one loads into the other.  `step()` is ONE CUDA
kernel for all tensors (csrc/optim.cu) instead of torch's nine foreach passes.  CUDA only.

`allreduce_gradients` is the data-parallel exchange the reference lacks (it is single-GPU): a SUM
all-reduce of every gradient over the process group (NCCL on GPUs); the division by the world size is
folded into the optimizer kernel (`grad_scale`).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import torch
import torch.distributed as dist

from . import _lib


class FusedRAdam(torch.optim.Optimizer):
    """torch.optim.RAdam semantics (betas, eps; weight_decay must be 0 as in the reference), one fused launch."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 check_nan=(), sanitize_grad=()):
        """check_nan / sanitize_grad: group names (the reference names every group, c_gaussian_model.py:430-449).
        The two per-iteration guards of train.py:244-253 are folded into the step kernel: the gradient of a
        `sanitize_grad` group goes through nan_to_num first (train.py:246-248 does that for "motion_opacity_var"),
        and a NaN parameter written into a `check_nan` group raises that group's flag (prune_nan_points,
        c_gaussian_model.py:1229-1241, tests "xyz" and "motion_xyz" with reductions + host waits after every step).
        Read the flags with nan_detected() (waits for the device) or poll_nan() (no wait, one step late)."""
        self._check_nan = frozenset(check_nan)
        self._sanitize = frozenset(sanitize_grad)
        self._flags = {}          # device -> int32[32 * launches-per-step] flag words, [device tensor, pinned copy, names]
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: %r" % (lr,))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: %r" % (eps,))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: %r" % (betas[0],))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: %r" % (betas[1],))
        if weight_decay != 0.0:
            raise ValueError("FusedRAdam implements weight_decay = 0 (what the reference uses)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0, allreduce_group=None, buckets: int = 6):
        """allreduce_group: data-parallel training - SUM all-reduce every gradient over that process group and step with the
        mean gradient, OVERLAPPED: all reductions are queued on the communication stream up front (in parameter order) and
        the step kernel is launched per bucket of tensors (`buckets` of roughly equal bytes) as soon as that bucket's
        reductions have completed, so the HBM-bound update of bucket k runs beside the NVLink-bound reduction of bucket
        k+1 (the un-overlapped form is allreduce_gradients() followed by step(grad_scale=...))."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        # groups may differ in betas/eps: one launch per distinct (beta1, beta2, eps); the reference has one
        batches = {}
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("RAdam does not support sparse gradients")
                if not p.is_cuda:
                    raise RuntimeError("ex4dgs_b200: FusedRAdam is CUDA-only")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedRAdam needs contiguous float32 parameters")
                st = self.state[p]
                if len(st) == 0:            # torch/optim/radam.py _init_group
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad if (p.grad.dtype == torch.float32 and p.grad.is_contiguous()) else p.grad.float().contiguous()
                batches.setdefault((float(beta1), float(beta2), float(group["eps"]), p.device), []).append(
                    (p, g, st["exp_avg"], st["exp_avg_sq"], float(group["lr"]), int(st["step"].item()), group.get("name")))
        works = {}
        if allreduce_group is not None and dist.is_available() and dist.is_initialized():
            for items in batches.values():
                for it in items:
                    works[id(it[0])] = dist.all_reduce(it[1], op=dist.ReduceOp.SUM, group=allreduce_group, async_op=True)
            grad_scale = grad_scale / dist.get_world_size(allreduce_group)
        slot = {}
        for (beta1, beta2, eps, dev), items in batches.items():
            stream = torch.cuda.current_stream(dev).cuda_stream
            # launch units: at most 32 tensors each; with an overlapped all-reduce, buckets of ~equal bytes in issue order
            if works:
                total = sum(it[0].numel() for it in items)
                target = max(1, total // max(1, buckets))
                units, cur, acc = [], [], 0
                for it in items:
                    cur.append(it)
                    acc += it[0].numel()
                    if acc >= target or len(cur) == 32:
                        units.append(cur)
                        cur, acc = [], 0
                if cur:
                    units.append(cur)
            else:
                units = [items[i:i + 32] for i in range(0, len(items), 32)]
            for chunk in units:
                for it in chunk:
                    w = works.get(id(it[0]))
                    if w is not None:
                        w.wait()              # the CURRENT stream waits for this tensor's reduction, the host does not
                arr = (_lib.RAdamTensor * len(chunk))()
                check = sanitize = 0
                for j, (a, (p, g, m, v, lr, step, name)) in enumerate(zip(arr, chunk)):
                    a.param, a.grad, a.exp_avg, a.exp_avg_sq = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
                    a.numel, a.lr, a.step = p.numel(), lr, step
                    check |= int(name in self._check_nan) << j
                    sanitize |= int(name in self._sanitize) << j
                flags_ptr = None
                if check:
                    fl = self._flags.get(dev)
                    base = slot.get(dev, 0)
                    if fl is None or fl[0].numel() < base + 32:
                        old = fl
                        fl = self._flags[dev] = [torch.zeros(base + 32, dtype=torch.int32, device=dev),
                                                 torch.zeros(base + 32, dtype=torch.int32).pin_memory(), {}]
                        if old is not None:
                            fl[0][:old[0].numel()] = old[0]
                            fl[2] = old[2]
                    for j, it in enumerate(chunk):
                        if (check >> j) & 1:
                            fl[2][it[6]] = base + j
                    flags_ptr = fl[0].data_ptr() + 4 * base
                    slot[dev] = base + 32
                with torch.cuda.device(dev):
                    rc = lib.ex4dgs_radam_step_ex(arr, len(chunk), beta1, beta2, eps, float(grad_scale), check, sanitize,
                                                  C.c_void_p(flags_ptr) if flags_ptr else None, C.c_void_p(stream))
                if rc < 0:
                    raise RuntimeError("ex4dgs_radam_step failed (%d): %s" % (rc, _lib.last_error()))
                # the kernel wrote parameters and moments through raw pointers: tell autograd (and every cache keyed on
                # Tensor._version, e.g. FusedGetters) that they changed, as torch's own optimizers do implicitly
                for (p, g, m, v, lr, step, name) in chunk:
                    torch.autograd.graph.increment_version(p)
                    torch.autograd.graph.increment_version(m)
                    torch.autograd.graph.increment_version(v)
        return loss

    def nan_detected(self) -> dict:
        """{group name: bool} for the check_nan groups: has any step so far written a NaN into the parameter?
        Waits for the device (like the reference's per-step `isnan().any()` test)."""
        out = {n: False for n in self._check_nan}
        for fl in self._flags.values():
            host = fl[0].cpu()
            for name, j in fl[2].items():
                out[name] = out[name] or bool(host[j].item())
        return out

    def poll_nan(self) -> dict:
        """Same without waiting: returns what the PREVIOUS poll's asynchronous copy delivered (valid once the
        stream has passed that point, e.g. after the training loop's `loss.item()`), then queues a new copy."""
        out = {n: False for n in self._check_nan}
        for fl in self._flags.values():
            for name, j in fl[2].items():
                out[name] = out[name] or bool(fl[1][j].item())
            fl[1].copy_(fl[0], non_blocking=True)
        return out

    def clear_nan(self) -> None:
        for fl in self._flags.values():
            fl[0].zero_()
            fl[1].zero_()


def allreduce_gradients(params: Iterable[torch.Tensor], group: Optional[dist.ProcessGroup] = None) -> float:
    """SUM all-reduce of every `.grad` (asynchronously issued, then waited).  Returns the factor
    1/world_size to hand to FusedRAdam.step(grad_scale=...) so that the update uses the mean gradient.
    Identity (returns 1.0) when torch.distributed is not initialised."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    works = []
    for p in params:
        if p.grad is not None:
            works.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in works:
        w.wait()
    return 1.0 / dist.get_world_size(group)
