"""Flat parameter / gradient / optimiser-state buckets of one network, and the fused optimiser step.

HBM layout (per network, 859.5 M parameters for SD1.5): one contiguous fp32 master bucket, one fp32 gradient
bucket (written in place by the wgrad kernels, reduced across ranks with ONE NCCL allreduce), one fp32 Adam
second-moment bucket (beta1 = 0 in the reference's recipe, so no first-moment bucket), an optional fp32 EMA
bucket and an optional bf16 shadow the tensor-core GEMMs read.  Every parameter is a view into the bucket at a
256-byte-aligned offset, with its diffusers name / logical shape (conv weights physically [O,3,3,I]).

Reference behaviour reproduced: training/sid_training_loop.py:291-292 (two Adam optimisers),
:458-462 / :541-549 (nan_to_num, fp16 clip, step), :553-565 (EMA), :316-323 (DDP mean-allreduce of gradients).
"""
import torch

from ._lib import lib, ptr, stream

_ALIGN = 64  # elements: 256 B fp32 / 128 B bf16


class FlatParams:
    def __init__(self, module, shadow=False):
        params = [p for p in module.parameters()]
        if not params:
            raise ValueError("module has no parameters")
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatParams needs the module on a CUDA device (no CPU path)")
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = total
        self.master = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = None       # allocated by ensure_grad(): frozen networks (teacher, G_ema) never pay for it
        self.shadow = torch.zeros(total, dtype=torch.bfloat16, device=dev) if shadow else None
        self.exp_avg_sq = None
        self.exp_avg = None
        self.step_count = 0
        self.reducer = None    # ddp.FlatDDP wrapping this network (its async reductions are joined before the optimiser step)
        self.params = params
        self.offsets = offs
        with torch.no_grad():
            for p, o in zip(params, offs):
                view = self._view_like(self.master, p, o)
                view.copy_(p)
                p.data = view
                p.grad = None
                p._shadow = self._view_like(self.shadow, p, o) if shadow else None
                p._flat = self
        if shadow:
            self.refresh_shadow()

    @staticmethod
    def _view_like(flat, p, off):
        """view of `flat` with p's logical shape and p's (dense, possibly channels_last) strides."""
        return flat.as_strided(p.shape, p.stride(), off)

    def ranges_of(self, modules):
        """merged [(start, end)] element ranges of the buckets covered by the parameters of `modules`."""
        index = {id(q): i for i, q in enumerate(self.params)}
        spans = []
        for m in modules:
            for q in m.parameters():
                i = index[id(q)]
                end = self.offsets[i + 1] if i + 1 < len(self.offsets) else self.numel
                spans.append((self.offsets[i], end))
        spans.sort()
        out = []
        for a, b in spans:
            if out and out[-1][1] == a:
                out[-1] = (out[-1][0], b)
            else:
                out.append((a, b))
        return out

    def refresh_shadow(self):
        if self.shadow is not None:
            lib.call("cast", ptr(self.master), ptr(self.shadow), self.numel, 0, 1, stream())

    def ensure_grad(self):
        """allocate the flat fp32 gradient bucket and point every parameter's .grad at its slice."""
        if self.grad is None:
            self.grad = torch.zeros(self.numel, dtype=torch.float32, device=self.master.device)
            for p, o in zip(self.params, self.offsets):
                p.grad = self._view_like(self.grad, p, o)
        return self.grad

    def zero_grad(self):
        if self.grad is None:
            self.ensure_grad()
        else:
            self.grad.zero_()

    def init_adam(self, beta1=0.0):
        self.exp_avg_sq = torch.zeros_like(self.master)
        self.exp_avg = torch.zeros_like(self.master) if beta1 != 0.0 else None
        self.step_count = 0

    def allreduce_grad(self, group=None, async_op=False):
        """Sum over ranks (the mean's 1/world is folded into adam_step's grad_scale)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.ensure_grad(), op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return None

    @staticmethod
    def adam_hyper(lr, betas, step, ema_beta=0.0):
        """the four per-step scalars of the fused optimiser pass (what sidlsg_adam_step derives from its arguments)"""
        return (float(lr), 1.0 - betas[0] ** step, (1.0 - betas[1] ** step) ** 0.5, float(ema_beta))

    def adam_step(self, lr, betas=(0.0, 0.999), eps=1e-8, grad_scale=1.0, clip=0.0, ema=None, ema_beta=0.0,
                  weight_decay=0.0, hyper=None):
        """nan_to_num + clip + Adam (+ EMA into `ema`, another FlatParams' master) (+ bf16 shadow), one launch."""
        if self.exp_avg_sq is None:
            self.init_adam(betas[0])
        if self.reducer is not None:
            self.reducer.finish()
        self.step_count += 1
        lib.call("adam_step", ptr(self.master), ptr(self.ensure_grad()), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                 ptr(ema.master) if ema is not None else None, ptr(self.shadow),
                 ptr(ema.shadow) if ema is not None else None, self.numel,
                 float(lr), float(betas[0]), float(betas[1]), float(eps), self.step_count, float(grad_scale),
                 float(clip), float(ema_beta), float(weight_decay), ptr(hyper), stream())

    def ema_into(self, ema, beta):
        lib.call("ema_update", ptr(self.master), ptr(ema.master), self.numel, float(beta), stream())
        ema.refresh_shadow()

    def copy_from(self, other):
        assert other.numel == self.numel
        self.master.copy_(other.master)
        self.refresh_shadow()

    def state_bytes(self):
        n = self.numel
        return n * 4 * (2 if self.grad is not None else 1) + (n * 2 if self.shadow is not None else 0) + (n * 4 if self.exp_avg_sq is not None else 0)


class FlatAdam:
    """`torch.optim.Adam` facade over a network's flat buckets: what `dnnlib.util.construct_class_by_name(params=...,
    class_name='torch.optim.Adam', lr=, betas=, eps=)` returns in the reference loop
    (/root/reference/training/sid_training_loop.py:291-292; sid_train.py:219-226) - `.zero_grad()`, `.step()`,
    `.state_dict()` / `.load_state_dict()` in torch.optim.Adam's own layout, `.param_groups`.  `step()` is the fused
    nan_to_num + clip + Adam (+ EMA) pass; `grad_scale`, `clip`, `ema`, `ema_beta` are per-call options."""

    def __init__(self, flat, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False, **_ignored):
        self.flat = flat
        self.param_groups = [dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, params=flat.params)]
        self.decoupled = decoupled        # AdamW (sid_train.py:223-226 optional)
        if flat.exp_avg_sq is None:
            flat.init_adam(betas[0])

    def zero_grad(self, set_to_none=True):
        del set_to_none
        self.flat.zero_grad()

    def step(self, grad_scale=1.0, clip=0.0, ema=None, ema_beta=0.0):
        g = self.param_groups[0]
        self.flat.adam_step(g["lr"], g["betas"], g["eps"], grad_scale=grad_scale, clip=clip, ema=ema, ema_beta=ema_beta,
                            weight_decay=g["weight_decay"] if self.decoupled else 0.0)

    def state_dict(self):
        from .training.checkpoint import adam_state_dict
        g, fl = self.param_groups[0], self.flat
        return adam_state_dict(fl.params, fl.exp_avg_sq, fl.offsets, fl.step_count, g["lr"], g["betas"], g["eps"],
                               exp_avg=fl.exp_avg)

    def load_state_dict(self, sd):
        from .training.checkpoint import load_adam_state_dict
        fl = self.flat
        fl.step_count = load_adam_state_dict(sd, fl.params, fl.exp_avg_sq, fl.offsets, exp_avg=fl.exp_avg)
