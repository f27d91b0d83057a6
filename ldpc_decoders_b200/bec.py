"""BEC channel and erasure decoder — GPU drop-ins for /root/reference/src/bec.py:11-18, 70-125."""
import numpy as np

from . import _lib
from .engine import engine_for, tables_of


class Channel:
    """bec.Channel (src/bec.py:11-18): symbol 2 marks an erasure, legacy global numpy RNG."""

    def __init__(self, p):
        self.p = p

    def send(self, x):
        tt = (np.random.random(x.shape) < self.p).astype(int)
        return np.clip(x + tt * 10, 0, 2)


class SPA:
    """bec.SPA (src/bec.py:70-122): integer message passing on symbols {0, 1, 2 = erasure}."""
    id_keys = ['max_iter']

    def __init__(self, p, _code, **kwargs):
        self.p = p
        self.max_iter = kwargs['max_iter']
        self.iter_cap = kwargs.get('iter_cap', 0)
        self.tables = tables_of(_code)
        self.xx, self.yy = self.tables.edge_chk, self.tables.edge_var
        self.engine = engine_for(self.tables, kwargs.get('device'))
        self._hist, self._frames = {}, 0

    def _count(self, iters):
        vals, cnt = np.unique(np.asarray(iters), return_counts=True)
        for v, c in zip(vals.tolist(), cnt.tolist()):
            self._hist[v] = self._hist.get(v, 0) + c
        self._frames += int(np.size(iters))

    def stats(self):
        top = max(self._hist) if self._hist else 0
        hist = [self._hist.get(i, 0) for i in range(top + 1)]
        tot = sum(i * c for i, c in enumerate(hist))
        return {'average': (tot / self._frames) if self._frames else 0., 'iter': hist}

    def decode_batch(self, Y, return_reason=False):
        Y = np.asarray(Y)
        if Y.ndim != 2 or Y.shape[1] != self.tables.n:
            raise ValueError("Y must be [B, n]")
        if Y.size and (Y.min() < 0 or Y.max() > 2):
            raise IndexError("BEC symbols must be 0, 1 or 2")     # the reference indexes a 3-entry table (bec.py:85)
        x_hat, iters, reason = self.engine.decode_host(_lib.CH_BEC, _lib.BEC, _lib.F32, 0.0,
                                                       np.ascontiguousarray(Y, np.uint8),
                                                       max_iter=self.max_iter, iter_cap=self.iter_cap)
        self._count(iters)
        x_hat = x_hat.astype(np.int64)
        return (x_hat, iters, reason) if return_reason else (x_hat, iters)

    def simulate_batch(self, x, B, seed, frame0=0, on_device=False):
        """On-device Monte-Carlo round: erase on the GPU, decode, count symbol errors (undecoded symbols count, main.py:41)."""
        from .biawgn import _simulate
        return _simulate(self, _lib.CH_BEC, self.p, x, B, seed, frame0, on_device)

    def simulate_round(self, x, B, seed, frame0, counters, nhist):
        from .biawgn import _simulate_round
        return _simulate_round(self, _lib.CH_BEC, self.p, x, B, seed, frame0, counters, nhist)

    def decode(self, y):
        y = np.asarray(y)
        x_hat, iters = self.decode_batch(y[None, :])
        if iters[0] == 0:
            return y                                              # bec.py:89: x_hat = y
        return x_hat[0]


class MSA(SPA):
    """bec.MSA is an alias of bec.SPA (src/bec.py:125)."""
    pass
