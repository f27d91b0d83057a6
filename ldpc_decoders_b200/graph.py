"""Edge tables of a parity-check matrix, in the reference's own edge order.

The reference keeps the edge list ``xx, yy = np.where(H)`` (/root/reference/src/bpa.py:12,
src/bec.py:77): row-major, i.e. sorted by check, then by variable.  Edge id = position in that
list.  The CUDA kernels need it check-major (contiguous per check) and variable-major (per
variable, its edge ids in ascending order — the summation order of scipy's coo.sum(axis=0) that
the reference inherits, src/math_utils.py:7)."""
import numpy as np


class Tables:
    """chk_ptr[m+1], edge_var[E], var_ptr[n+1], var_edges[E] (int32, C-contiguous)."""

    def __init__(self, m, n, rows, cols):
        rows = np.asarray(rows, np.int64).ravel()
        cols = np.asarray(cols, np.int64).ravel()
        if rows.size == 0 or rows.size != cols.size:
            raise ValueError("empty or inconsistent edge list")
        if rows.min() < 0 or rows.max() >= m or cols.min() < 0 or cols.max() >= n:
            raise ValueError("edge index out of range")
        order = np.lexsort((cols, rows))
        rows, cols = rows[order], cols[order]
        if ((np.diff(rows) == 0) & (np.diff(cols) == 0)).any():
            raise ValueError("duplicate edge")
        self.m, self.n, self.E = int(m), int(n), int(rows.size)
        self.edge_chk = np.ascontiguousarray(rows, np.int32)
        self.edge_var = np.ascontiguousarray(cols, np.int32)
        self.chk_ptr = np.zeros(self.m + 1, np.int32)
        np.cumsum(np.bincount(rows, minlength=self.m), out=self.chk_ptr[1:])
        self.var_ptr = np.zeros(self.n + 1, np.int32)
        np.cumsum(np.bincount(cols, minlength=self.n), out=self.var_ptr[1:])
        self.var_edges = np.ascontiguousarray(np.argsort(cols, kind="stable"), np.int32)
        self.check_degrees = np.diff(self.chk_ptr)
        self.var_degrees = np.diff(self.var_ptr)

    @classmethod
    def from_dense(cls, parity_mtx):
        H = np.asarray(parity_mtx)
        if H.ndim != 2:
            raise ValueError("parity_mtx must be 2-D")
        rows, cols = np.where(H)
        return cls(H.shape[0], H.shape[1], rows, cols)

    def dense(self, dtype=np.int64):
        H = np.zeros((self.m, self.n), dtype)
        H[self.edge_chk, self.edge_var] = 1
        return H

    def syndrome(self, x):
        """(H @ x) % 2 for x [..., n] of 0/1 ints, through the edge list (no dense H)."""
        x = np.asarray(x)
        s = np.add.reduceat(x[..., self.edge_var].astype(np.int64), self.chk_ptr[:-1].astype(np.int64), axis=-1)
        s[..., self.check_degrees == 0] = 0
        return s % 2
