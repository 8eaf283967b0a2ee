"""Scalar statistics of the loop with the reference's surface (/root/reference/torch_utils/training_stats.py):
`report`, `report0`, `init_multiprocessing`, `Collector` / `default_collector` (`update()`, `as_dict()`, `mean`, `std`,
`num`, `[name]`).  Per name three moments [count, sum, sum of squares] in float64; `Collector.update()` folds the
per-process counters into cumulative totals with ONE all_reduce per call when multi-process collection is on
(the reference's single per-tick collective, :255).  Fresh implementation."""
import math
import re

import torch

_rank = 0
_sync_device = None
_pending = {}      # name -> [count, sum, sumsq] since the last sync (this process)
_total = {}        # name -> cumulative [count, sum, sumsq] over all processes


def init_multiprocessing(rank, sync_device):
    global _rank, _sync_device
    _rank, _sync_device = rank, sync_device


def report(name, value):
    acc = _pending.setdefault(name, [0.0, 0.0, 0.0])
    v = torch.as_tensor(value)
    if v.numel():
        v = v.detach().to(torch.float64).flatten()
        acc[0] += float(v.numel())
        acc[1] += float(v.sum())
        acc[2] += float(v.square().sum())
    return value


def report0(name, value):
    report(name, value if _rank == 0 else [])
    return value


def _sync(names):
    if not names:
        return
    deltas = torch.tensor([_pending.get(n, [0.0, 0.0, 0.0]) for n in names], dtype=torch.float64)
    for n in names:
        _pending[n] = [0.0, 0.0, 0.0]
    if _sync_device is not None and torch.distributed.is_initialized():
        d = deltas.to(_sync_device)
        torch.distributed.all_reduce(d)
        deltas = d.cpu()
    for n, row in zip(names, deltas):
        tot = _total.setdefault(n, torch.zeros(3, dtype=torch.float64))
        tot += row


class Collector:
    """Averages over the window between the last two `update()` calls (keep_previous=True keeps the last value of a
    name that received no samples in the window)."""

    def __init__(self, regex=".*", keep_previous=True):
        self._regex = re.compile(regex)
        self._keep = keep_previous
        self._seen = {}     # name -> cumulative moments at the previous update()
        self._window = {}   # name -> moments of the current window

    def names(self):
        return [n for n in set(_pending) | set(_total) if self._regex.fullmatch(n)]

    def update(self):
        names = sorted(self.names())       # identical order on every rank: one collective
        _sync(names)
        for n in names:
            cur = _total[n].clone() if n in _total else torch.zeros(3, dtype=torch.float64)
            delta = cur - self._seen.get(n, torch.zeros(3, dtype=torch.float64))
            self._seen[n] = cur
            if not self._keep or float(delta[0]) != 0 or n not in self._window:
                self._window[n] = delta

    def num(self, name):
        return int(self._window[name][0]) if name in self._window else 0

    def mean(self, name):
        m = self._window.get(name)
        if m is None or int(m[0]) == 0:
            return float("nan")
        return float(m[1] / m[0])

    def std(self, name):
        m = self._window.get(name)
        if m is None or int(m[0]) == 0 or not math.isfinite(float(m[1])):
            return float("nan")
        if int(m[0]) == 1:
            return 0.0
        mean = float(m[1] / m[0])
        return math.sqrt(max(float(m[2] / m[0]) - mean * mean, 0.0))

    def as_dict(self):
        return {n: {"num": self.num(n), "mean": self.mean(n), "std": self.std(n)} for n in sorted(self._window)}

    def __getitem__(self, name):
        return self.mean(name)


default_collector = Collector()
