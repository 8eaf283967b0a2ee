"""Process-group shim with the reference's names (/root/reference/torch_utils/distributed.py:14-60).

Same env contract (MASTER_ADDR / MASTER_PORT / RANK / LOCAL_RANK / WORLD_SIZE, single-process defaults), same
helpers.  Differences, all additive:
  * without a CUDA device `init()` falls back to the gloo backend (the reference hard-requires NCCL + CUDA, :26-28),
    which is what lets the loop's host logic be tested on CPU at world_size 2;
  * `init()` is idempotent (a launcher such as bench.py may already have created the group);
  * the default MASTER_ADDR is 127.0.0.1 rather than 'localhost' (container hostnames may not resolve).
"""
import datetime
import os

import torch

from . import training_stats

_DEFAULT_ENV = (("MASTER_ADDR", "127.0.0.1"), ("MASTER_PORT", "29500"), ("RANK", "0"), ("LOCAL_RANK", "0"),
                ("WORLD_SIZE", "1"))


def init(backend=None, timeout_s=600):
    for key, val in _DEFAULT_ENV:
        os.environ.setdefault(key, val)
    use_cuda = torch.cuda.is_available()
    if not torch.distributed.is_initialized():
        if backend is None:
            backend = "nccl" if (use_cuda and os.name != "nt") else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", get_local_rank())
        torch.distributed.init_process_group(backend=backend, init_method="env://",
                                             timeout=datetime.timedelta(seconds=timeout_s), **kw)
    if use_cuda:
        torch.cuda.set_device(get_local_rank())
    sync_device = None
    if get_world_size() > 1:
        sync_device = torch.device("cuda", get_local_rank()) if use_cuda else torch.device("cpu")
    training_stats.init_multiprocessing(rank=get_rank(), sync_device=sync_device)


def get_rank():
    return torch.distributed.get_rank() if torch.distributed.is_initialized() else 0


def get_local_rank():
    return int(os.environ.get("LOCAL_RANK", "0"))


def get_world_size():
    return torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1


def should_stop():
    return False


def update_progress(cur, total):
    del cur, total


def print0(*args, **kwargs):
    if get_rank() == 0:
        print(*args, **kwargs)


def barrier():
    """torch.distributed.barrier() that is a no-op outside a process group (the reference calls the torch function
    directly, which requires `init()`; sid_training_loop.py:221,231,312,380)."""
    if torch.distributed.is_initialized():
        torch.distributed.barrier()
