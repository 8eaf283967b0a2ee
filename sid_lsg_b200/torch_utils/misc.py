"""The two `torch_utils.misc` helpers on the distillation path (+ the parameter-copy helpers resume uses):
`InfiniteSampler` (/root/reference/torch_utils/misc.py:110-141), `ddp_sync` (:168-175),
`copy_params_and_buffers` (:154-162), `check_ddp_consistency` (:180-191).  Fresh implementations."""
import contextlib
import re

import numpy as np
import torch


class InfiniteSampler(torch.utils.data.Sampler):
    """Endless, rank-strided index stream over `dataset`.

    Order semantics of the reference (pinned by tests/golden/sampler_order.pt): a `RandomState(seed)` permutation;
    position k of the global stream belongs to rank `k % num_replicas`; after each position is visited it is swapped
    with a random earlier position inside a window of `window_size * len(dataset)` entries, so later epochs are
    re-shuffled incrementally.  Every rank advances the SAME random stream (one draw per global position), which is
    what keeps the ranks' shards disjoint."""

    def __init__(self, dataset, rank=0, num_replicas=1, shuffle=True, seed=0, window_size=0.5):
        if len(dataset) <= 0:
            raise AssertionError("InfiniteSampler: empty dataset")
        if num_replicas <= 0 or not (0 <= rank < num_replicas):
            raise AssertionError("InfiniteSampler: need 0 <= rank < num_replicas")
        if not (0 <= window_size <= 1):
            raise AssertionError("InfiniteSampler: window_size must be in [0, 1]")
        self.dataset, self.rank, self.num_replicas = dataset, rank, num_replicas
        self.shuffle, self.seed, self.window_size = shuffle, seed, window_size

    def __iter__(self):
        n = len(self.dataset)
        order = np.arange(n)
        rng, window = None, 0
        if self.shuffle:
            rng = np.random.RandomState(self.seed)
            rng.shuffle(order)
            window = int(np.rint(n * self.window_size))
        pos = 0
        while True:
            slot = pos % n
            if pos % self.num_replicas == self.rank:
                yield order[slot]
            if window >= 2:
                other = (slot - rng.randint(window)) % n
                order[slot], order[other] = order[other], order[slot]
            pos += 1


def params_and_buffers(module):
    assert isinstance(module, torch.nn.Module)
    return list(module.parameters()) + list(module.buffers())


def named_params_and_buffers(module):
    assert isinstance(module, torch.nn.Module)
    return list(module.named_parameters()) + list(module.named_buffers())


@torch.no_grad()
def copy_params_and_buffers(src_module, dst_module, require_all=False):
    """dst tensors <- same-named src tensors (resume path, sid_training_loop.py:296-304).  A destination living in
    flat buckets gets its bf16 shadow refreshed afterwards."""
    src = dict(named_params_and_buffers(src_module))
    for name, tensor in named_params_and_buffers(dst_module):
        if name not in src:
            if require_all:
                raise AssertionError("copy_params_and_buffers: %s missing in the source module" % name)
            continue
        tensor.copy_(src[name])
    flat = getattr(getattr(dst_module, "module", dst_module), "flat", None)
    if flat is not None:
        flat.refresh_shadow()


@contextlib.contextmanager
def ddp_sync(module, sync):
    """Gradient synchronisation on/off for one forward+backward, as the reference's context manager: a module that
    offers `no_sync()` (torch DistributedDataParallel, or this package's `parallel.FlatDDP`) skips its gradient
    allreduce inside the block when `sync` is false; any other module is left alone."""
    assert isinstance(module, torch.nn.Module)
    no_sync = getattr(module, "no_sync", None)
    if sync or no_sync is None:
        yield
    else:
        with no_sync():
            yield


def check_ddp_consistency(module, ignore_regex=None):
    """assert that every parameter / buffer equals rank 0's copy."""
    assert isinstance(module, torch.nn.Module)
    for name, tensor in named_params_and_buffers(module):
        fullname = type(module).__name__ + "." + name
        if ignore_regex is not None and re.fullmatch(ignore_regex, fullname):
            continue
        mine = tensor.detach()
        if mine.is_floating_point():
            mine = torch.nan_to_num(mine)
        theirs = mine.clone()
        torch.distributed.broadcast(tensor=theirs, src=0)
        if not bool((mine == theirs).all()):
            raise AssertionError("parameter differs across ranks: " + fullname)
