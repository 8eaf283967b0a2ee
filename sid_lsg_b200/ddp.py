"""Data-parallel wrapper for networks whose weight gradients live in a flat bucket (params.FlatParams).

Stands where the reference puts `torch.nn.parallel.DistributedDataParallel(module, device_ids=[device],
broadcast_buffers=False, find_unused_parameters=False)` (/root/reference/training/sid_training_loop.py:316-323):
same constructor keywords, `.module`, `no_sync()` (so `misc.ddp_sync(module_ddp, sync)` works unchanged,
/root/reference/torch_utils/misc.py:168-175), parameters broadcast from rank 0 at construction, gradients averaged over
ranks once per synchronised backward.

torch's DDP cannot wrap these networks: the wgrad kernels accumulate straight into the flat gradient bucket and
autograd never sees a weight gradient, so its reducer hooks would never fire.  Instead the UNet reports, from inside
backward, which stage's gradients are final (unet.forward: `mark`), and this class starts an NCCL allreduce of that
stage's bucket range on a side stream at once - the reduction of the up blocks runs under the backward of the down
blocks, and only the first stage's range is exposed.  `finish()` (called by FlatParams.adam_step) joins the side stream.
The SUM is taken here; the 1/world of the mean is folded into the optimiser kernel's `grad_scale`.
"""
import contextlib

import torch
import torch.distributed as dist


class FlatDDP(torch.nn.Module):
    def __init__(self, module, device_ids=None, broadcast_buffers=False, find_unused_parameters=False,
                 process_group=None, overlap=True):
        super().__init__()
        del device_ids, broadcast_buffers, find_unused_parameters   # accepted for call-site compatibility
        if getattr(module, "flat", None) is None:
            raise RuntimeError("FlatDDP wraps a network with flat buckets: call unet.flatten_() first")
        self.module = module
        self.group = process_group
        self.overlap = overlap
        self._sync = True
        self._armed = 0          # synchronised forwards recorded in the current autograd graph
        self._done_from = None   # lowest stage whose range has been handed to NCCL in this backward
        self._works = []
        self._callback_queued = False
        self._comm = None
        self._fwd_stream = None
        self.reduced_elems = 0   # elements reduced since construction (tests / logging)
        if self.world > 1:
            flat = module.flat
            dist.broadcast(flat.master, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                           group=self.group)
            flat.refresh_shadow()
        module.flat.reducer = self

    @property
    def world(self):
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    @contextlib.contextmanager
    def no_sync(self):
        prev, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = prev

    def forward(self, *args, **kwargs):
        m = self.module
        if self._sync and self.world > 1 and torch.is_grad_enabled() and any(p.requires_grad for p in m.flat.params[:1]):
            idx = self._armed
            self._armed += 1
            if m.flat.grad.is_cuda:
                self._fwd_stream = torch.cuda.current_stream()
            m._grad_ready = lambda stage, idx=idx: self._ready(stage, idx)
        return m(*args, **kwargs)

    # -- called from the autograd thread during backward ---------------------------------------------------------
    def _ready(self, stage, idx):
        if not self._callback_queued:
            self._callback_queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._end_of_backward)
        # a network called several times in one graph (multi-step generator): only the EARLIEST call's backward sees
        # final gradients
        if idx != 0 or not self.overlap:
            return
        stages = self.module.grad_stages()
        hi = len(stages) if self._done_from is None else self._done_from
        if stage < hi:
            for k in range(stage, hi):
                self._launch(stages[k])
            self._done_from = stage

    def _launch(self, ranges):
        grad = self.module.flat.grad
        for a, b in ranges:
            chunk = grad[a:b]
            if grad.is_cuda:
                if self._comm is None:
                    self._comm = torch.cuda.Stream(device=grad.device)
                ev = torch.cuda.Event()
                ev.record(self._fwd_stream)
                self._comm.wait_event(ev)
                with torch.cuda.stream(self._comm):
                    self._works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            else:
                self._works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.reduced_elems += b - a

    def _end_of_backward(self):
        """everything not yet reduced (stage 0 always; all stages without overlap)."""
        stages = self.module.grad_stages()
        hi = len(stages) if self._done_from is None else self._done_from
        for k in range(0, hi):
            self._launch(stages[k])
        self._done_from = None
        self._armed = 0
        self._callback_queued = False

    def finish(self):
        """make the current stream wait for the outstanding reductions (before the optimiser reads the bucket)."""
        for w in self._works:
            w.wait()
        self._works = []
