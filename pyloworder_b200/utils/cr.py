"""Channel timers with the reference's names (pyLOM/utils/cr.py:16-246): `@cr('POD.run')`,
`cr_start/cr_stop`, `cr_info()`.  Device work is timed with CUDA events on the current stream
(the reference uses MPI.Wtime without synchronising, so its GPU numbers are launch times)."""
import functools
import time

import torch

_channels = {}
_open = {}


class _Chan:
    __slots__ = ("n", "tsum", "tmax", "tmin", "pending")

    def __init__(self):
        self.n, self.tsum, self.tmax, self.tmin, self.pending = 0, 0.0, 0.0, float("inf"), []

    def add(self, dt):
        self.n += 1
        self.tsum += dt
        self.tmax = max(self.tmax, dt)
        self.tmin = min(self.tmin, dt)


_MAX_PENDING = 256       # event pairs kept per channel before they are folded into the totals


def _resolve(ch, only_done=False):
    keep = []
    for e0, e1 in ch.pending:
        if only_done and not e1.query():
            keep.append((e0, e1))
            continue
        e1.synchronize()
        ch.add(e0.elapsed_time(e1) * 1e-3)
    ch.pending = keep


def cr_start(name, suffix=0):
    key = f"{name}{suffix:02d}" if suffix else name
    if torch.cuda.is_available():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        _open[key] = e
    else:
        _open[key] = time.perf_counter()


def cr_stop(name, suffix=0):
    key = f"{name}{suffix:02d}" if suffix else name
    s = _open.pop(key, None)
    if s is None:
        return
    ch = _channels.setdefault(key, _Chan())
    if torch.cuda.is_available():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ch.pending.append((s, e))
        if len(ch.pending) > _MAX_PENDING:      # long runs that never call cr_info(): fold the finished pairs, bounded memory
            _resolve(ch, only_done=True)
            if len(ch.pending) > _MAX_PENDING:
                _resolve(ch)
    else:
        ch.add(time.perf_counter() - s)


def _arg_device(args):
    for x in args:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    return None


def cr(name):
    """Channel timer + device guard: the call runs with the CUDA device of its first device-tensor argument current,
    so streams, workspaces and the library's per-device state all belong to the device the data lives on (a tensor on
    cuda:1 while cuda:0 is current would otherwise launch on the wrong device)."""
    def deco(fn):
        @functools.wraps(fn)
        def wrap(*a, **k):
            dev = _arg_device(a)
            if dev is not None and dev.index != torch.cuda.current_device():
                with torch.cuda.device(dev):
                    return wrap(*a, **k)
            cr_start(name)
            try:
                return fn(*a, **k)
            finally:
                cr_stop(name)
        return wrap
    return deco


cr_nvtx = cr


def cr_time(name):
    ch = _channels.get(name)
    if ch is None:
        return 0.0
    _resolve(ch)
    return ch.tsum


def cr_reset():
    _channels.clear()
    _open.clear()


def cr_info(rank=0):
    """Print the channel table (rank 0), like pyLOM/utils/cr.py:147-197."""
    from .parall import pprint
    rows = []
    for k, ch in sorted(_channels.items()):
        _resolve(ch)
        if ch.n:
            rows.append((k, ch.n, ch.tmin, ch.tmax, ch.tsum / ch.n, ch.tsum))
    pprint(rank, "\ncr_info:")
    for k, n, tmin, tmax, tavg, tsum in rows:
        pprint(rank, f"  {k:<28s} n={n:<5d} min={tmin:.3e} max={tmax:.3e} avg={tavg:.3e} sum={tsum:.3e} s")
    return {k: dict(n=n, min=tmin, max=tmax, avg=tavg, sum=tsum) for k, n, tmin, tmax, tavg, tsum in rows}
