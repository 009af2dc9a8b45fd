"""Cheap scalar statistics across devices and processes.

Surface of the reference's torch_utils/training_stats.py (`init_multiprocessing` :37, `report` :59, `report0` :109,
`Collector` :119-246): `report(name, value)` folds any bag of scalars into three running moments (count, sum, sum of
squares) held on the value's own device, so calling it costs one tiny reduction and no synchronisation; `Collector.update()`
pulls the moments of the names it watches to one device, sums them over ranks with ONE all-reduce for all names
(`training_loop.py` calls it once per tick) and exposes per-interval mean / std.
"""
import re

import numpy as np
import torch

_MOMENTS = 3                       # count, sum, sum of squares
_ACC_DTYPE = torch.float64


class _State:
    rank = 0
    sync_device = None             # None = single process
    synced = False
    live = {}                      # name -> {device: moments accumulated since the last sync}
    total = {}                     # name -> cumulative moments over all ranks (CPU)


def init_multiprocessing(rank, sync_device):
    """Call after `torch.distributed.init_process_group()` and before the first `Collector.update()`."""
    assert not _State.synced
    _State.rank = rank
    _State.sync_device = sync_device


def report(name, value):
    slot = _State.live.setdefault(name, {})
    elems = torch.as_tensor(value)
    if elems.numel() == 0:
        return value
    e = elems.detach().flatten().to(torch.float32)
    m = torch.stack([torch.ones_like(e).sum(), e.sum(), e.square().sum()]).to(_ACC_DTYPE)
    acc = slot.get(m.device)
    if acc is None:
        slot[m.device] = m.clone()
    else:
        acc.add_(m)
    return value


def report0(name, value):
    """Only rank 0's scalars count; the other ranks still register the name so every rank syncs the same list."""
    report(name, value if _State.rank == 0 else [])
    return value


def _sync(names):
    if not names:
        return []
    _State.synced = True
    dev = _State.sync_device if _State.sync_device is not None else torch.device('cpu')
    rows = torch.zeros([len(names), _MOMENTS], dtype=_ACC_DTYPE, device=dev)
    for i, name in enumerate(names):
        for acc in _State.live[name].values():
            rows[i] += acc.to(dev)
            acc.zero_()
    if _State.sync_device is not None:
        torch.distributed.all_reduce(rows)
    rows = rows.cpu()
    out = []
    for i, name in enumerate(names):
        tot = _State.total.setdefault(name, torch.zeros([_MOMENTS], dtype=_ACC_DTYPE))
        tot.add_(rows[i])
        out.append((name, tot))
    return out


class Collector:
    """Averages of the scalars reported between the last two `update()` calls, for the names matching `regex`."""

    def __init__(self, regex='.*', keep_previous=True):
        self._regex = re.compile(regex)
        self._keep_previous = keep_previous
        self._seen = {}
        self._delta = {}
        self.update()
        self._delta.clear()

    def names(self):
        return [n for n in _State.live if self._regex.fullmatch(n)]

    def update(self):
        if not self._keep_previous:
            self._delta.clear()
        for name, tot in _sync(self.names()):
            prev = self._seen.setdefault(name, torch.zeros([_MOMENTS], dtype=_ACC_DTYPE))
            d = tot - prev
            prev.copy_(tot)
            if float(d[0]) != 0:
                self._delta[name] = d

    def _get(self, name):
        assert self._regex.fullmatch(name)
        return self._delta.setdefault(name, torch.zeros([_MOMENTS], dtype=_ACC_DTYPE))

    def num(self, name):
        return int(self._get(name)[0])

    def mean(self, name):
        d = self._get(name)
        return float('nan') if int(d[0]) == 0 else float(d[1] / d[0])

    def std(self, name):
        d = self._get(name)
        n = int(d[0])
        if n == 0 or not np.isfinite(float(d[1])):
            return float('nan')
        if n == 1:
            return 0.0
        mean = float(d[1] / d[0])
        return float(np.sqrt(max(float(d[2] / d[0]) - mean * mean, 0)))

    def as_dict(self):
        from .. import dnnlib
        return dnnlib.EasyDict({n: dnnlib.EasyDict(num=self.num(n), mean=self.mean(n), std=self.std(n)) for n in self.names()})

    def __getitem__(self, name):
        return self.mean(name)
