"""CPU tier: the C-ABI library builds for sm_100a, loads, and exports every symbol the header declares;
host-side table building and argument validation.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import _golden as G
from ldpc_decoders_b200 import Tables, _lib
from ldpc_decoders_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    B.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "ldpc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ldpc_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.ldpc_abi_version() == 2


def test_library_is_sm100a_only():
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % _lib.LIB_PATH).read()
    if out.strip():
        assert "sm_100a" in out and "sm_90" not in out


def test_tables_follow_np_where_order():
    for name in ("4_2_test", "7_4_hamming", "1200_3_6_rand_ldpc_1", "1200_rho_x5_rand_ldpc_10", "margulis"):
        H = G.dense_H(name)
        t = Tables.from_dense(H)
        xx, yy = np.where(H)
        assert (t.edge_chk == xx).all() and (t.edge_var == yy).all()
        assert (t.check_degrees == H.sum(1)).all() and (t.var_degrees == H.sum(0)).all()
        for v in range(0, t.n, max(1, t.n // 7)):
            e = t.var_edges[t.var_ptr[v]:t.var_ptr[v + 1]]
            assert (np.diff(e) > 0).all() and (yy[e] == v).all()
        x = np.random.RandomState(0).randint(0, 2, size=(5, t.n))
        assert (t.syndrome(x) == (x @ H.T) % 2).all()
        m, n, r, c = G.code_tables(name)
        t2 = Tables(m, n, r[::-1], c[::-1])                # any edge order in, canonical order out
        assert (t2.edge_var == t.edge_var).all() and (t2.var_edges == t.var_edges).all()


def test_tables_reject_bad_input():
    with pytest.raises(ValueError):
        Tables(2, 3, [0, 0], [1, 1])
    with pytest.raises(ValueError):
        Tables(2, 3, [0, 2], [1, 1])
    with pytest.raises(ValueError):
        Tables(2, 3, [], [])


def test_create_validates_tables_and_has_no_cpu_path(lib):
    t = Tables.from_dense(G.dense_H("7_4_hamming"))
    h = ctypes.c_void_p()
    bad = t.edge_var.copy()
    bad[0], bad[1] = bad[1], bad[0]
    rc = lib.ldpc_create(ctypes.byref(h), 0, t.n, t.m, t.E, t.chk_ptr.ctypes.data, bad.ctypes.data,
                         t.var_ptr.ctypes.data, t.var_edges.ctypes.data)
    assert rc == -1 and b"ascending" in lib.ldpc_last_error(None)
    import torch
    if not torch.cuda.is_available():
        rc = lib.ldpc_create(ctypes.byref(h), 0, t.n, t.m, t.E, t.chk_ptr.ctypes.data, t.edge_var.ctypes.data,
                             t.var_ptr.ctypes.data, t.var_edges.ctypes.data)
        assert rc == -2 and b"no CUDA device" in lib.ldpc_last_error(None)
        from ldpc_decoders_b200 import LdpcError, bsc
        with pytest.raises(LdpcError):
            bsc.MSA(.1, t, max_iter=10)                    # the product path fails loudly without a GPU


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ldpc_decoders_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(from|import)\s+oracle|libldpc_oracle|oracle[/.]\w", src), f


def test_bit_packed_row_helpers_round_trip(lib):
    """Host-side packing for LDPC_IN_PACKED / LDPC_OUT_PACKED: the layout ldpc_packed_row_bytes promises (rows padded to
    16 bytes, bit v = bit (v & 7) of byte (v >> 3)), value plane then erasure plane for BEC symbols."""
    from ldpc_decoders_b200 import engine as E
    rng = np.random.RandomState(0)
    for n in (1, 7, 8, 9, 127, 128, 129, 1200, 2640):
        assert E.packed_row_bytes(n) == lib.ldpc_packed_row_bytes(n) and E.packed_row_bytes(n) % 16 == 0
        assert E.packed_row_bytes(n) * 8 >= n
        Y = rng.randint(0, 2, size=(5, n)).astype(np.uint8)
        P = E.pack_bits(Y)
        assert P.shape == (5, E.packed_row_bytes(n)) and (E.unpack_bits(P, n) == Y).all()
        for v in (0, n // 2, n - 1):
            assert (((P[:, v >> 3] >> (v & 7)) & 1) == Y[:, v]).all()
        S = rng.randint(0, 3, size=(5, n)).astype(np.uint8)
        Q = E.pack_symbols(S)
        s = E.packed_row_bytes(n)
        assert Q.shape == (5, 2 * s) and (E.unpack_symbols(Q, n) == S).all()
        assert (E.unpack_bits(Q[:, :s], n) == (S == 1)).all() and (E.unpack_bits(Q[:, s:], n) == (S == 2)).all()
    assert lib.ldpc_packed_row_bytes(0) == 0


def test_monte_carlo_entry_points_validate_arguments(lib):
    assert lib.ldpc_mc_scratch_bytes(None, 2, 0, 0, 128) == 0
    assert lib.ldpc_count_accumulate(None, None, None, None, 1, None, None, 0, None) == -1
    assert lib.ldpc_mc_round(None, 2, 0, 0, 1.0, 1.0, None, 0, 0, 1, 10, 0, None, 0, None, 0, 0, None) == -1
