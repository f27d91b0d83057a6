"""oracle/oracle.py — ctypes wrapper of the CPU checker.  TEST INFRASTRUCTURE ONLY.

Loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  ``ldpc_decoders_b200`` never imports this module.

The decoders restated in ``ldpc_oracle.c`` follow /root/reference/src/bpa.py,
src/bec.py:70-122 and src/math_utils.py; parity is pinned by
tests/test_oracle_golden.py against fixtures made from the real reference.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libldpc_oracle.so")

MSA, SPA = 0, 1
REASONS = {0: "decoded", 1: "maximum", 2: "stopping", 4: "cap"}


def build(force=False):
    """Compile libldpc_oracle.so with the Makefile next to this file."""
    srcs = [os.path.join(_HERE, f) for f in ("ldpc_oracle.c", "bp_body.inc", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "libldpc_oracle.so"], check=True,
                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
        u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        vp = ctypes.c_void_p
        for sfx, ct in (("f64", np.float64), ("f32", np.float32)):
            fp = np.ctypeslib.ndpointer(ct, flags="C_CONTIGUOUS")
            f = getattr(L, "oracle_bp_" + sfx)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_int] * 4 + [i32p] * 4 + [ctypes.c_int, fp, vp, ctypes.c_int, ctypes.c_int,
                                                           u8p, i32p, u8p, vp, ctypes.c_int]
            f = getattr(L, "oracle_cn_" + sfx)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_int] * 4 + [i32p, i32p, fp, fp]
            f = getattr(L, "oracle_vn_" + sfx)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_int] * 3 + [i32p] * 4 + [fp, fp, fp, fp, u8p]
        L.oracle_bec.restype = ctypes.c_int
        L.oracle_bec.argtypes = [ctypes.c_int] * 3 + [i32p] * 4 + [ctypes.c_int, u8p, ctypes.c_int, ctypes.c_int,
                                                                  u8p, i32p, u8p, ctypes.c_int]
        L.oracle_llr_bsc.restype = None
        L.oracle_llr_bsc.argtypes = [ctypes.c_double, ctypes.c_size_t, u8p,
                                     np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")]
        L.oracle_llr_biawgn.restype = None
        L.oracle_llr_biawgn.argtypes = [ctypes.c_double, ctypes.c_size_t,
                                        np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS"),
                                        np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")]
        _lib = L
    return _lib


class Graph:
    """Edge tables of H in the reference's edge order (np.where(H), src/bpa.py:12)."""

    def __init__(self, m, n, rows, cols):
        rows = np.asarray(rows, np.int64)
        cols = np.asarray(cols, np.int64)
        order = np.lexsort((cols, rows))            # row-major = check-major, ascending variable
        rows, cols = rows[order], cols[order]
        self.m, self.n, self.E = int(m), int(n), int(rows.size)
        self.rows, self.cols = rows, cols
        self.chk_ptr = np.zeros(self.m + 1, np.int32)
        np.cumsum(np.bincount(rows, minlength=self.m), out=self.chk_ptr[1:])
        self.edge_var = np.ascontiguousarray(cols, np.int32)
        self.var_ptr = np.zeros(self.n + 1, np.int32)
        np.cumsum(np.bincount(cols, minlength=self.n), out=self.var_ptr[1:])
        self.var_edges = np.ascontiguousarray(np.argsort(cols, kind="stable"), np.int32)

    @classmethod
    def from_dense(cls, H):
        H = np.asarray(H)
        rows, cols = np.where(H)
        return cls(H.shape[0], H.shape[1], rows, cols)

    def dense(self, dtype=np.int64):
        H = np.zeros((self.m, self.n), dtype)
        H[self.rows, self.cols] = 1
        return H

    def _tabs(self):
        return (self.n, self.m, self.E, self.chk_ptr, self.edge_var, self.var_ptr, self.var_edges)


def _vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def bp_decode(g, algo, priors, y_hard=None, max_iter=10, iter_cap=0, want_marg=False, nthreads=1):
    """Batch decode.  priors [B,n] float64 or float32 (dtype selects the arithmetic).

    Returns dict(x_hat uint8 [B,n], iters int32 [B], reason uint8 [B], marg [B,n] or None).
    """
    priors = np.ascontiguousarray(priors)
    if priors.dtype not in (np.float64, np.float32):
        raise TypeError("priors must be float64 or float32")
    single = priors.ndim == 1
    P = priors.reshape(-1, g.n)
    B = P.shape[0]
    yh = None
    if y_hard is not None:
        yh = np.ascontiguousarray(np.asarray(y_hard).reshape(B, g.n), np.uint8)
    x_hat = np.zeros((B, g.n), np.uint8)
    iters = np.zeros(B, np.int32)
    reason = np.zeros(B, np.uint8)
    marg = np.zeros((B, g.n), P.dtype) if want_marg else None
    fn = lib().oracle_bp_f64 if P.dtype == np.float64 else lib().oracle_bp_f32
    rc = fn(algo, *g._tabs(), B, P, _vp(yh), int(max_iter), int(iter_cap), x_hat, iters, reason, _vp(marg),
            int(nthreads))
    if rc:
        raise RuntimeError("oracle_bp failed rc=%d" % rc)
    if single:
        return dict(x_hat=x_hat[0], iters=int(iters[0]), reason=int(reason[0]),
                    marg=None if marg is None else marg[0])
    return dict(x_hat=x_hat, iters=iters, reason=reason, marg=marg)


def cn_sweep(g, algo, v2c):
    v2c = np.ascontiguousarray(v2c)
    c2v = np.empty_like(v2c)
    fn = lib().oracle_cn_f64 if v2c.dtype == np.float64 else lib().oracle_cn_f32
    rc = fn(algo, g.n, g.m, g.E, g.chk_ptr, g.edge_var, v2c, c2v)
    if rc:
        raise RuntimeError("oracle_cn failed rc=%d" % rc)
    return c2v


def vn_sweep(g, prior, c2v):
    c2v = np.ascontiguousarray(c2v)
    prior = np.ascontiguousarray(prior, c2v.dtype)
    v2c = np.empty_like(c2v)
    marg = np.empty(g.n, c2v.dtype)
    x_hat = np.empty(g.n, np.uint8)
    fn = lib().oracle_vn_f64 if c2v.dtype == np.float64 else lib().oracle_vn_f32
    fn(g.n, g.m, g.E, g.chk_ptr, g.edge_var, g.var_ptr, g.var_edges, prior, c2v, v2c, marg, x_hat)
    return v2c, marg, x_hat


def bec_decode(g, y, max_iter=10, iter_cap=0, nthreads=1):
    y = np.ascontiguousarray(y, np.uint8)
    single = y.ndim == 1
    Y = y.reshape(-1, g.n)
    B = Y.shape[0]
    x_hat = np.zeros((B, g.n), np.uint8)
    iters = np.zeros(B, np.int32)
    reason = np.zeros(B, np.uint8)
    rc = lib().oracle_bec(*g._tabs(), B, Y, int(max_iter), int(iter_cap), x_hat, iters, reason, int(nthreads))
    if rc:
        raise RuntimeError("oracle_bec failed rc=%d" % rc)
    if single:
        return dict(x_hat=x_hat[0], iters=int(iters[0]), reason=int(reason[0]))
    return dict(x_hat=x_hat, iters=iters, reason=reason)


def noise_var(snr_in_db):
    """src/biawgn.py:10 — same Python float pow as the reference."""
    return 10 ** (-snr_in_db / 10)


def llr_bsc(p, y):
    """src/bsc.py:21,25"""
    y = np.ascontiguousarray(y, np.uint8)
    out = np.empty(y.shape, np.float64)
    lib().oracle_llr_bsc(float(np.log(1 - p) - np.log(p)), y.size, y.reshape(-1), out.reshape(-1))
    return out


def llr_biawgn(snr_in_db, y):
    """src/biawgn.py:28"""
    y = np.ascontiguousarray(y, np.float64)
    out = np.empty(y.shape, np.float64)
    lib().oracle_llr_biawgn(float(noise_var(snr_in_db)), y.size, y.reshape(-1), out.reshape(-1))
    return out
