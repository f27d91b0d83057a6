"""BIAWGN channel and decoder adapters — GPU drop-ins for /root/reference/src/biawgn.py:10-42."""
import numpy as np

from . import _lib, bpa

noise_var = lambda snr_in_db: 10 ** (-snr_in_db / 10)      # src/biawgn.py:10


class Channel:
    """biawgn.Channel (src/biawgn.py:13-18): {0,1} -> {-1,+1} plus N(0, sigma^2), legacy global numpy RNG."""

    def __init__(self, snr_in_db):
        self.std_dev = np.sqrt(noise_var(snr_in_db))

    def send(self, x):
        return (2 * x - 1) + np.random.normal(0, self.std_dev, x.shape)


class LLR:
    """biawgn.LLR (src/biawgn.py:21-28): priors = -2 y / noise_var."""

    def __init__(self, snr_in_db, dec, dtype=None):
        self.noise_var, self.dec = noise_var(snr_in_db), dec
        self.dtype = np.dtype(np.float64 if dtype is None else dtype)
        self.stats = self.dec.stats

    def decode(self, y):
        y = np.asarray(y)
        return self.dec.decode(y, (-2 * y / self.noise_var).astype(self.dtype, copy=False))

    def decode_batch(self, Y, return_reason=False):
        """Y [B,n] float64 or float32 received values; priors = dtype((-2 * float64(y)) / noise_var) on the GPU."""
        Y = np.ascontiguousarray(Y)
        if Y.dtype not in (np.float32, np.float64):
            Y = Y.astype(np.float64)
        dt = _lib.F32 if self.dtype == np.float32 else _lib.F64
        x_hat, iters, reason = self.dec.engine.decode_host(_lib.CH_BIAWGN, self.dec._algo, dt, self.noise_var, Y,
                                                           max_iter=self.dec.max_iter, iter_cap=self.dec.iter_cap)
        self.dec._count(iters)
        x_hat = x_hat.astype(np.int64)
        return (x_hat, iters, reason) if return_reason else (x_hat, iters)


    def simulate_batch(self, x, B, seed, frame0=0, on_device=False):
        """Draw B received frames of the word x on the GPU (global frame indices frame0 ..), decode, count bit errors.
        Returns numpy (bit_errs [B], iters [B]) — or, with on_device=True, the two CUDA int32 tensors themselves (no
        synchronisation; the caller gathers them); nothing but these two vectors ever crosses PCIe."""
        return _simulate(self, _lib.CH_BIAWGN, self.noise_var, x, B, seed, frame0, on_device)

    def simulate_round(self, x, B, seed, frame0, counters, nhist):
        """The same round with the counters kept on the GPU (Engine.mc_round): nothing crosses PCIe at all."""
        return _simulate_round(self, _lib.CH_BIAWGN, self.noise_var, x, B, seed, frame0, counters, nhist)


def _xdev(adapter, eng, x):
    """The transmitted word as a cached device tensor (the cache key is the word itself)."""
    import torch
    key = (np.asarray(x, np.uint8).tobytes(), eng.device)
    if getattr(adapter, "_xkey", None) != key:
        adapter._xdev = torch.from_numpy(np.ascontiguousarray(x, np.uint8)).to(eng._dev())
        adapter._xkey = key
        adapter._bufs = {}
    return adapter._xdev


def _simulate(adapter, channel, param, x, B, seed, frame0, on_device=False):
    dec = getattr(adapter, "dec", adapter)                 # bec.SPA is its own decoder core
    eng = dec.engine
    dt = _lib.F32 if getattr(adapter, "dtype", np.dtype(np.float32)) == np.float32 else _lib.F64
    xd = _xdev(adapter, eng, x)
    out = eng.simulate(channel, getattr(dec, "_algo", _lib.BEC), dt, param, B, seed, frame0, x=xd, max_iter=dec.max_iter,
                       iter_cap=dec.iter_cap, bufs=adapter._bufs)
    if on_device:
        return out["bit_errs"], out["iters"]
    errs, iters = out["bit_errs"].cpu().numpy(), out["iters"].cpu().numpy()
    dec._count(iters)
    return errs, iters


def _simulate_round(adapter, channel, param, x, B, seed, frame0, counters, nhist):
    dec = getattr(adapter, "dec", adapter)
    eng = dec.engine
    dt = _lib.F32 if getattr(adapter, "dtype", np.dtype(np.float32)) == np.float32 else _lib.F64
    eng.mc_round(channel, getattr(dec, "_algo", _lib.BEC), dt, param, B, seed, frame0, counters, nhist,
                 x=_xdev(adapter, eng, x), max_iter=dec.max_iter, iter_cap=dec.iter_cap)


class SPA(LLR):
    id_keys = bpa.SPA.id_keys

    def __init__(self, snr_in_db, _code, **kwargs):
        super().__init__(snr_in_db, bpa.SPA(_code, **kwargs), kwargs.get('dtype'))


class MSA(LLR):
    id_keys = bpa.MSA.id_keys

    def __init__(self, snr_in_db, _code, **kwargs):
        super().__init__(snr_in_db, bpa.MSA(_code, **kwargs), kwargs.get('dtype'))
