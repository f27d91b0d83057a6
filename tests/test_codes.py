"""CPU tier: the code loader (SURVEY a-0) and the seeded samplers (f-4).

The golden edge lists in tests/golden/codes.npz were written by make_golden.py from the REFERENCE's own loader
(/root/reference/src/codes.py:93-105) for all 27 shipped files and its 4 built-in matrices.  Here the product's
loader is held to them: a text file in the shipped format is re-created from each golden edge list (1-based, and
0-based for `margulis`, the one shipped file that starts at 0 and so exercises the reference's `var - 1` wrap of
variable 0 into the last column), loaded with codes.load_parity_txt and compared edge by edge.  When the reference
tree is present (the build container) the shipped files themselves are loaded too.
"""
import os

import numpy as np
import pytest

import _golden as G
from ldpc_decoders_b200 import codes
from ldpc_decoders_b200.graph import Tables

REF_CODES = "/root/reference/data/codes"
NAMES = sorted({k.split("__")[0] for k in np.load(os.path.join(G.GOLD, "codes.npz")).files})
BUILTIN = ("4_2_test", "6_2_3_ldpc", "7_4_hamming", "12_3_4_ldpc")
ZERO_BASED = ("margulis",)


def write_txt(path, m, n, rows, cols, zero_based):
    """One check per line, variable numbers separated by blanks — the format of data/codes/*.txt."""
    with open(path, "w") as fp:
        for c in range(m):
            v = cols[rows == c]
            # the loader stores file number k in column (k - 1) mod n (codes.py:103), so column j was written as
            # j + 1 in a 1-based file and as (j + 1) mod n in a 0-based one
            nums = (v + 1) % n if zero_based else v + 1
            fp.write("   ".join(str(int(k)) for k in nums) + "\n")


def same_edges(tab, m, n, rows, cols):
    return (tab.m, tab.n) == (m, n) and (tab.edge_chk == rows).all() and (tab.edge_var == cols).all()


def test_all_31_codes_are_in_the_fixture():
    assert len(NAMES) == 31 and set(BUILTIN) <= set(NAMES)


@pytest.mark.parametrize("name", [n for n in NAMES if n not in BUILTIN])
def test_loader_reproduces_reference_edge_lists(name, tmp_path):
    m, n, rows, cols = G.code_tables(name)
    zero = name in ZERO_BASED
    p = tmp_path / (name + ".txt")
    write_txt(str(p), m, n, rows, cols, zero)
    if zero:
        assert " 0 " in " " + open(str(p)).read().replace("\n", " ") + " "       # the file really is 0-based
    tab = codes.load_parity_txt(str(p))
    assert same_edges(tab, m, n, rows, cols)
    got = codes.get_code(name, str(tmp_path))
    assert got.get_n() == n and got.get_k() == n - m and same_edges(got.tables, m, n, rows, cols)


@pytest.mark.parametrize("name", BUILTIN)
def test_builtin_matrices_match_reference(name):
    m, n, rows, cols = G.code_tables(name)
    c = codes.get_code(name)
    assert same_edges(c.tables, m, n, rows, cols)
    H = c.parity_mtx
    xx, yy = np.where(H)
    assert (xx == rows).all() and (yy == cols).all()


@pytest.mark.skipif(not os.path.isdir(REF_CODES), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", [n for n in NAMES if n not in BUILTIN])
def test_loader_on_the_shipped_files(name):
    m, n, rows, cols = G.code_tables(name)
    assert same_edges(codes.load_parity_txt(os.path.join(REF_CODES, name + ".txt")), m, n, rows, cols)


def test_margulis_wrap_quirk():
    """0-based file: variable 0 lands in the LAST column (python's negative index), everything else shifts down."""
    m, n, rows, cols = G.code_tables("margulis")
    assert n == 2640 and (cols == n - 1).sum() == 3           # the three edges of file-variable 0


def test_loader_rejects_bad_index_base(tmp_path):
    p = tmp_path / "bad.txt"
    p.write_text("2 3 4\n3 4 5\n")
    with pytest.raises(Exception):
        codes.load_parity_txt(str(p))


@pytest.mark.parametrize("n,dv,dc", [(12, 3, 6), (20, 3, 4), (24, 3, 6), (48, 3, 6), (96, 3, 6), (8, 2, 4), (1200, 3, 6)])
def test_random_regular_is_regular_and_simple(n, dv, dc):
    """a vectorised double-edge repair dropped sockets (306 of 1020 small draws were irregular)."""
    for seed in range(100 if n < 200 else 8):
        t = codes.random_regular(n, dv, dc, seed).tables
        assert t.E == n * dv and t.m == n * dv // dc
        assert (t.var_degrees == dv).all() and (t.check_degrees == dc).all()
        key = t.edge_chk.astype(np.int64) * n + t.edge_var
        assert np.unique(key).size == t.E
        again = codes.random_regular(n, dv, dc, seed).tables
        assert (again.edge_var == t.edge_var).all()           # seeded: reproducible


def test_random_regular_long_code():
    t = codes.random_regular(64800, 3, 6, seed=0).tables
    assert (t.var_degrees == 3).all() and (t.check_degrees == 6).all() and t.E == 194400


def test_random_irregular_follows_the_degree_profile():
    tab = Tables(*G.code_tables("1200_rho_x5_rand_ldpc_1"))
    prof = codes.variable_degree_counts(tab)
    E0 = sum(d * c for d, c in prof.items())
    prof2 = dict(prof)
    if E0 % 6:                                                  # the sampler wants E to be a multiple of dc
        prof2[1] = prof2.get(1, 0) + (6 - E0 % 6)
    t = codes.random_irregular(prof2, 6, seed=1).tables
    assert t.n == sum(prof2.values())
    # parallel edges cancel in pairs, so a few variables end up two edges short (as in the shipped files)
    short = sum(d * c for d, c in prof2.items()) - t.E
    assert 0 <= short <= 40 and short % 2 == 0
    assert set(np.unique(t.check_degrees)) <= {2, 4, 6}
