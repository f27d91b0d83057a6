"""BSC channel and decoder adapters — GPU drop-ins for /root/reference/src/bsc.py:11-39."""
import numpy as np

from . import _lib, bpa


class Channel:
    """bsc.Channel (src/bsc.py:11-16): flips each bit with probability p, legacy global numpy RNG."""

    def __init__(self, p):
        self.p = p

    def send(self, x):
        return (x + (np.random.random(x.shape) < self.p)) % 2


class LLR:
    """bsc.LLR (src/bsc.py:19-25): priors = llr * (1 - 2y), llr = log(1-p) - log(p)."""

    def __init__(self, p, dec, dtype=None):
        self.p = p
        self.llr, self.dec = np.log(1 - p) - np.log(p), dec
        self.dtype = np.dtype(np.float64 if dtype is None else dtype)
        self.stats = self.dec.stats

    def decode(self, y):
        y = np.asarray(y)
        return self.dec.decode(y, (self.llr * (1 - 2 * y)).astype(self.dtype, copy=False))

    def decode_batch(self, Y, return_reason=False):
        """Y [B,n] hard bits.  LLR map, transpose, decode and hard decisions all run on the GPU."""
        Y = np.ascontiguousarray(Y, np.uint8)
        dt = _lib.F32 if self.dtype == np.float32 else _lib.F64
        x_hat, iters, reason = self.dec.engine.decode_host(_lib.CH_BSC, self.dec._algo, dt, self.llr, Y,
                                                           max_iter=self.dec.max_iter, iter_cap=self.dec.iter_cap)
        self.dec._count(iters)
        x_hat = x_hat.astype(np.int64)
        return (x_hat, iters, reason) if return_reason else (x_hat, iters)


    def simulate_batch(self, x, B, seed, frame0=0, on_device=False):
        """On-device Monte-Carlo round (see biawgn.LLR.simulate_batch)."""
        from .biawgn import _simulate
        return _simulate(self, _lib.CH_BSC, self.p, x, B, seed, frame0, on_device)

    def simulate_round(self, x, B, seed, frame0, counters, nhist):
        from .biawgn import _simulate_round
        return _simulate_round(self, _lib.CH_BSC, self.p, x, B, seed, frame0, counters, nhist)


class SPA(LLR):
    id_keys = bpa.SPA.id_keys

    def __init__(self, p, _code, **kwargs):
        super().__init__(p, bpa.SPA(_code, **kwargs), kwargs.get('dtype'))


class MSA(LLR):
    id_keys = bpa.MSA.id_keys

    def __init__(self, p, _code, **kwargs):
        super().__init__(p, bpa.MSA(_code, **kwargs), kwargs.get('dtype'))
