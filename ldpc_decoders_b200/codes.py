"""Sparse-native code loading and generation.

The reference builds a DENSE int64 H for every code (/root/reference/src/codes.py:93-105) — 16.8 GB at
n = 64800.  Here a code is its edge list (``graph.Tables``); the dense matrix is only materialised on
request for small codes.  File format and index convention are the reference's: one check per line,
variable indices separated by blanks, always shifted by -1 (``mtx[chk-1, var-1]``, codes.py:103 — so a
0-based file wraps variable 0 into the last column, exactly as the reference loader does).
"""
import os

import numpy as np

from .graph import Tables


class Code:
    """Duck-type of the reference's codes.Code for decoder constructors: .parity_mtx, get_n(), get_k()."""

    def __init__(self, tables, name=None):
        self.tables = tables
        self.name = name
        self.gen_mtx = None
        self._dense = None

    @property
    def parity_mtx(self):
        if self._dense is None:
            if self.tables.m * self.tables.n > (1 << 28):
                raise MemoryError("dense H of %dx%d refused; use .tables" % (self.tables.m, self.tables.n))
            self._dense = self.tables.dense()
        return self._dense

    def get_n(self):
        return self.tables.n

    def get_k(self):
        return self.tables.n - self.tables.m


def load_parity_txt(path):
    """Edge tables of a data/codes/*.txt file (reference loader semantics, codes.py:93-105)."""
    rows, cols, max_ind, min_ind, m = [], [], None, None, 0
    with open(path, "r") as fp:
        for line in fp:
            idx = [int(tok) for tok in line.split()]
            if not idx:
                continue
            max_ind = max(idx) if max_ind is None else max(max_ind, max(idx))
            min_ind = min(idx) if min_ind is None else min(min_ind, min(idx))
            rows.extend([m] * len(idx))
            cols.extend(idx)
            m += 1
    if m == 0:
        raise ValueError("no checks in %s" % path)
    if min_ind not in (0, 1):
        raise Exception("Minimum index is not 0 or 1.")
    n = max_ind + (0 if min_ind == 1 else 1)
    cols = (np.asarray(cols, np.int64) - 1) % n          # mtx[chk_num - 1, var_num - 1] with python's negative wrap
    rows = np.asarray(rows, np.int64)
    key = np.unique(rows * n + cols)                     # a repeated index on one line sets the same entry once
    return Tables(m, n, key // n, key % n)


def codes_dir():
    return os.path.abspath(os.environ.get("FILE_CODES_DIR", os.path.join("data", "codes")))


_BUILTIN = {
    # the small textbook matrices the reference hard-codes (codes.py:27-66), as check rows of variable indices
    "4_2_test": (5, [[0, 1], [1, 2, 3], [3, 4]]),
    "6_2_3_ldpc": (6, [[0, 1, 2], [3, 4, 5], [2, 3, 5], [0, 1, 4]]),
    "7_4_hamming": (7, [[3, 4, 5, 6], [1, 2, 5, 6], [0, 2, 4, 6]]),
    "12_3_4_ldpc": (12, [[2, 5, 6, 7], [0, 1, 4, 11], [3, 8, 9, 10], [1, 5, 6, 9], [0, 2, 7, 10],
                         [3, 4, 8, 11], [0, 3, 4, 6], [5, 7, 10, 11], [1, 2, 8, 9]]),
}


def get_code(name, directory=None):
    """Code by name: a built-in, or <directory>/<name>.txt (default $FILE_CODES_DIR or ./data/codes)."""
    if name in _BUILTIN:
        n, checks = _BUILTIN[name]
        rows = np.concatenate([[c] * len(vs) for c, vs in enumerate(checks)])
        cols = np.concatenate(checks)
        return Code(Tables(len(checks), n, rows, cols), name)
    path = os.path.join(directory or codes_dir(), name + ".txt")
    return Code(load_parity_txt(path), name)


def random_regular(n, dv, dc, seed=0):
    """Seeded (dv, dc)-regular code from the configuration model in O(E): permute the variable sockets
    against the check sockets and repair double edges by socket swaps.  (The reference's sampler,
    codes.py:108-120, is O(m n log n) Python on a dense matrix and unseeded; for n = 64800 use this.)"""
    if (n * dv) % dc:
        raise ValueError("n * dv must be a multiple of dc")
    m = n * dv // dc
    E = n * dv
    rng = np.random.default_rng(seed)
    var_sock = np.repeat(np.arange(n, dtype=np.int64), dv)
    chk_sock = np.repeat(np.arange(m, dtype=np.int64), dc)
    perm = rng.permutation(E)
    cols = var_sock[perm]
    for _ in range(1000):
        key = chk_sock * n + cols
        order = np.argsort(key, kind="stable")
        dup = order[1:][key[order][1:] == key[order][:-1]]
        if dup.size == 0:
            break
        # one socket pair at a time: a swap of two sockets keeps the multiset of variable sockets (and so every
        # degree) whatever the indices are, which a vectorised fancy-index swap with repeated indices does not
        for d in dup.tolist():
            o = int(rng.integers(0, E - 1))
            o += o >= d                                   # any socket but d itself
            cols[d], cols[o] = cols[o], cols[d]
    else:
        raise RuntimeError("could not remove double edges")
    if not ((np.bincount(cols, minlength=n) == dv).all() and (np.bincount(chk_sock, minlength=m) == dc).all()):
        raise AssertionError("random_regular lost a socket: the code is not (%d,%d)-regular" % (dv, dc))
    return Code(Tables(m, n, chk_sock, cols), "%d_%d_%d_cfg_seed%d" % (n, dv, dc, seed))


def variable_degree_counts(tables):
    """{degree: number of variables} of a code — the input of random_irregular for "another code like this one"."""
    deg = np.asarray(tables.var_degrees)
    return {int(d): int(c) for d, c in enumerate(np.bincount(deg)) if c}


def random_irregular(degree_counts, dc, seed=0):
    """Seeded irregular code in O(E), the construction of the reference's generator (src/ldpc.py:149-192,
    gen_rand_irg_ldpc) without its dense matrix: one socket per edge end, variables of degree d get d sockets
    (degree_counts = {d: how many}, in ascending-degree order of the variable index like add_sockets, ldpc.py:141-146),
    the variable sockets are shuffled against check sockets dealt round-robin (sockets_chk = range(m) * dc), and
    parallel edges cancel in pairs (parity_mtx += 1 per socket, even entries -> 0, ldpc.py:186-188), so a few
    checks / variables end up two edges short — the shipped 1200_rho_x5 files show exactly that (dc in {4, 6}).
    The degree distribution itself comes from the reference's density-evolution solver (out of scope);
    variable_degree_counts() of an existing code is the usual source."""
    degs = np.concatenate([np.full(int(c), int(d), np.int64) for d, c in sorted(degree_counts.items()) if int(c) > 0 and int(d) > 0]
                          or [np.zeros(0, np.int64)])
    n = int(sum(int(c) for c in degree_counts.values()))
    zero = int(degree_counts.get(0, 0))
    E = int(degs.sum())
    if E == 0 or E % dc:
        raise ValueError("the number of edges (%d) must be a positive multiple of dc" % E)
    m = E // dc
    var_sock = np.repeat(np.arange(zero, zero + degs.size, dtype=np.int64), degs)      # degree-0 variables come first
    chk_sock = np.tile(np.arange(m, dtype=np.int64), dc)
    rng = np.random.default_rng(seed)
    cols = var_sock[rng.permutation(E)]
    key, mult = np.unique(chk_sock * n + cols, return_counts=True)
    key = key[mult % 2 == 1]
    return Code(Tables(m, n, key // n, key % n), "%d_irr_dc%d_cfg_seed%d" % (n, dc, seed))
