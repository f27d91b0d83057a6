"""GPU drop-ins for the reference's decoder cores ``bpa.SPA`` / ``bpa.MSA``.

Same constructor and call protocol as /root/reference/src/bpa.py:6-102::

    dec = MSA(parity_mtx, max_iter=10)          # bpa.py:9-15, 79-84
    x_hat = dec.decode(y, priors)               # bpa.py:17-63, one frame

plus the batch form the GPU exists for::

    x_hat, iters = dec.decode_batch(Y, priors)  # Y, priors: [B, n]

The arithmetic type follows ``priors.dtype`` exactly like the reference (float32 priors ->
float32 messages).  MSA is bit-exact at either type; SPA float64 mirrors the
reference formula, SPA float32 uses the cancellation-free hyperbolic-pair rule (csrc/ldpc_math.cuh cn_spa_sc).
There is no CPU fallback: constructing a decoder without a CUDA device raises.
"""
import numpy as np

from . import _lib
from .engine import engine_for, tables_of

#: safety bound on the reference's "max_iter <= 0 means unlimited" (bpa.py:28); frames that reach it
#: are reported with reason 'cap' instead of looping forever on a non-converging frame.
DEFAULT_ITER_CAP = 1000


class BPA:
    id_keys = ['max_iter']
    _algo = None

    def __init__(self, parity_mtx, **kwargs):
        self.max_iter = kwargs['max_iter']
        self.iter_cap = kwargs.get('iter_cap', DEFAULT_ITER_CAP)
        self.tables = tables_of(parity_mtx)
        self.parity_mtx = parity_mtx
        self.xx, self.yy = self.tables.edge_chk, self.tables.edge_var      # == np.where(parity_mtx), bpa.py:12
        self.engine = engine_for(self.tables, kwargs.get('device'))
        self._hist = {}
        self._frames = 0

    # ---- iteration statistics in the shape of admm.ADMM.stats() (src/admm.py:36-40; hook: main.py:34)
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

    # ---- the iteration-0 exit of bpa.py:29 on a real-valued y (BIAWGN).  x_hat = y there, so the test is
    # ((H @ y) % 2 == 0).all(); for continuous noise it never holds, but integer-valued y (noise-free
    # tests) do pass it and the reference then returns y itself.
    def _iter0_exit(self, Y):
        s = np.add.reduceat(Y[:, self.tables.edge_var], self.tables.chk_ptr[:-1].astype(np.int64), axis=1)
        s[:, self.tables.check_degrees == 0] = 0
        return ((s % 2) == 0).all(axis=1)

    def decode_batch(self, Y, priors, return_reason=False, strict_iter0=False):
        """Decode B frames.  Y [B,n] is what the reference calls y (hard bits for BSC, reals for
        BIAWGN); priors [B,n] float64 or float32.  Returns (x_hat int64 [B,n], iters int32 [B])."""
        Y = np.asarray(Y)
        priors = np.ascontiguousarray(priors)
        if priors.dtype not in (np.float32, np.float64):
            priors = priors.astype(np.float64)
        if Y.shape != priors.shape or Y.ndim != 2 or Y.shape[1] != self.tables.n:
            raise ValueError("Y and priors must both be [B, n]")
        dtype = _lib.F32 if priors.dtype == np.float32 else _lib.F64
        if Y.shape[0] == 0:                                 # an empty batch decodes to empty results
            empty = (np.empty((0, self.tables.n), np.int64), np.empty(0, np.int32), np.empty(0, np.uint8))
            return empty if return_reason else empty[:2]
        hard = np.issubdtype(Y.dtype, np.integer) or Y.dtype == np.bool_
        import torch
        dev = self.engine._dev()
        d_pri = torch.from_numpy(priors).to(dev)
        d_hard = torch.from_numpy(np.ascontiguousarray(Y, np.uint8)).to(dev) if hard else None
        out = self.engine.decode_device(self._algo, d_pri, y_hard=d_hard, max_iter=self.max_iter,
                                        iter_cap=self.iter_cap)
        x_hat = out['x_hat'].cpu().numpy().astype(np.int64)
        iters = out['iters'].cpu().numpy()
        reason = out['reason'].cpu().numpy()
        if not hard and strict_iter0:
            z = self._iter0_exit(Y)
            if z.any():
                iters = iters.copy()
                iters[z] = 0
                reason[z] = 0
                x_hat = x_hat.astype(Y.dtype)
                x_hat[z] = Y[z]
        self._count(iters)
        return (x_hat, iters, reason) if return_reason else (x_hat, iters)

    def decode(self, y, priors):
        """One frame, reference signature (bpa.py:17).  Returns y itself on a 0-iteration exit (bpa.py:20,24)."""
        y = np.asarray(y)
        x_hat, iters = self.decode_batch(y[None, :], np.asarray(priors)[None, :], strict_iter0=True)
        if iters[0] == 0:
            return y
        return x_hat[0]


class SPA(BPA):
    """bpa.SPA (src/bpa.py:66-75)."""
    _algo = _lib.SPA


class MSA(BPA):
    """bpa.MSA (src/bpa.py:78-102)."""
    _algo = _lib.MSA
