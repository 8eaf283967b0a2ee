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
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.shadow = torch.zeros(total, dtype=torch.bfloat16, device=dev) if shadow else None
        self.exp_avg_sq = None
        self.exp_avg = None
        self.step_count = 0
        self.params = params
        self.offsets = offs
        with torch.no_grad():
            for p, o in zip(params, offs):
                view = self._view_like(self.master, p, o)
                view.copy_(p)
                p.data = view
                p.grad = self._view_like(self.grad, p, o)
                p._shadow = self._view_like(self.shadow, p, o) if shadow else None
                p._flat = self
        if shadow:
            self.refresh_shadow()

    @staticmethod
    def _view_like(flat, p, off):
        """view of `flat` with p's logical shape and p's (dense, possibly channels_last) strides."""
        return flat.as_strided(p.shape, p.stride(), off)

    def refresh_shadow(self):
        if self.shadow is not None:
            lib.call("cast", ptr(self.master), ptr(self.shadow), self.numel, 0, 1, stream())

    def zero_grad(self):
        self.grad.zero_()

    def init_adam(self, beta1=0.0):
        self.exp_avg_sq = torch.zeros_like(self.master)
        self.exp_avg = torch.zeros_like(self.master) if beta1 != 0.0 else None
        self.step_count = 0

    def allreduce_grad(self, group=None, async_op=False):
        """Sum over ranks (the mean's 1/world is folded into adam_step's grad_scale)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return None

    def adam_step(self, lr, betas=(0.0, 0.999), eps=1e-8, grad_scale=1.0, clip=0.0, ema=None, ema_beta=0.0,
                  weight_decay=0.0):
        """nan_to_num + clip + Adam (+ EMA into `ema`, another FlatParams' master) (+ bf16 shadow), one launch."""
        if self.exp_avg_sq is None:
            self.init_adam(betas[0])
        self.step_count += 1
        lib.call("adam_step", ptr(self.master), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                 ptr(ema.master) if ema is not None else None, ptr(self.shadow), self.numel,
                 float(lr), float(betas[0]), float(betas[1]), float(eps), self.step_count, float(grad_scale),
                 float(clip), float(ema_beta), float(weight_decay), stream())

    def ema_into(self, ema, beta):
        lib.call("ema_update", ptr(self.master), ptr(ema.master), self.numel, float(beta), stream())

    def copy_from(self, other):
        assert other.numel == self.numel
        self.master.copy_(other.master)
        self.refresh_shadow()

    def state_bytes(self):
        n = self.numel
        return n * 4 * 2 + (n * 2 if self.shadow is not None else 0) + (n * 4 if self.exp_avg_sq is not None else 0)
