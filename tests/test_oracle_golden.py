"""Pin the CPU oracle (oracle/ldpc_oracle.c) against fixtures made from the unmodified reference.

MSA and BEC: bit-exact words, iteration counts, exit reasons (and marginals).
SPA (float64): identical words / iteration counts; messages within the conditioning-aware
tolerance below (glibc vs numpy transcendental kernels differ in the last ulp and
2*atanh(q) amplifies that by ~exp(|LLR|)).
"""
import numpy as np
import pytest

import _golden as G
from oracle import oracle as O


def graph(name):
    return O.Graph(*G.code_tables(name))


def spa_tol(ref):
    a = np.abs(ref)
    return 1e-12 * np.maximum(1.0, a) + 1e-15 * np.exp(np.minimum(a, 40.0))


def test_graph_tables_match_np_where():
    for name in ("7_4_hamming", "1200_3_6_rand_ldpc_1", "1200_rho_x5_rand_ldpc_10"):
        H = G.dense_H(name)
        g = O.Graph.from_dense(H)
        xx, yy = np.where(H)
        assert (g.rows == xx).all() and (g.cols == yy).all()
        assert (np.diff(g.chk_ptr) == H.sum(1)).all() and (np.diff(g.var_ptr) == H.sum(0)).all()
        for v in (0, g.n // 2, g.n - 1):
            e = g.var_edges[g.var_ptr[v]:g.var_ptr[v + 1]]
            assert (np.diff(e) > 0).all() and (yy[e] == v).all()


@pytest.mark.parametrize("k", G.kats(), ids=lambda k: "%s-%s-%s" % (k["channel"], k["code"], k["decoder"]))
def test_kat(k):
    g = graph(k["code"])
    y = np.array(k["y"])
    if k["channel"] == "bec":
        r = O.bec_decode(g, y.astype(np.uint8), max_iter=k["max_iter"])
        assert r["x_hat"].tolist() == k["x_hat"] and r["iters"] == k["iters"] and r["reason"] == k["reason"]
        return
    if k["channel"] == "bsc":
        pri, yh = O.llr_bsc(k["param"], y.astype(np.uint8)), y.astype(np.uint8)
    else:
        pri, yh = O.llr_biawgn(k["param"], y), None
    algo = O.MSA if k["decoder"] == "MSA" else O.SPA
    r = O.bp_decode(g, algo, pri, y_hard=yh, max_iter=k["max_iter"], want_marg=True)
    assert r["x_hat"].tolist() == k["x_hat"] and r["iters"] == k["iters"]
    ref = np.array(k["marg"])
    if algo == O.MSA:
        assert (r["marg"] == ref).all()
    else:
        assert (np.abs(r["marg"] - ref) <= spa_tol(ref)).all()
    assert k["passed"] and r["x_hat"].tolist() == k["x"]


@pytest.mark.parametrize("rec", G.runs(), ids=lambda r: r["key"])
def test_seeded_run(rec):
    g = graph(rec["code"])
    x, Y = G.run_inputs(rec)
    gold = G.run_arrays(rec)
    if rec["channel"] == "bec":
        r = O.bec_decode(g, Y.astype(np.uint8), max_iter=rec["max_iter"], nthreads=4)
        assert (r["x_hat"] == gold["x_hat"]).all()
        assert (r["iters"] == gold["iters"]).all()
        assert (r["reason"] == gold["reason"]).all()
        return
    dt = np.float64 if rec["dtype"] == "f64" else np.float32
    if rec["channel"] == "bsc":
        pri, yh = O.llr_bsc(rec["param"], Y.astype(np.uint8)).astype(dt), Y.astype(np.uint8)
    else:
        pri, yh = O.llr_biawgn(rec["param"], Y).astype(dt), None
    algo = O.MSA if rec["decoder"] == "MSA" else O.SPA
    r = O.bp_decode(g, algo, pri, y_hard=yh, max_iter=rec["max_iter"], want_marg=True, nthreads=4)
    if algo == O.MSA:
        assert (r["iters"] == gold["iters"]).all()
        assert (r["x_hat"] == gold["x_hat"]).all()
        assert (r["reason"] == gold["reason"]).all()
        m4 = gold["marg4"]
        assert r["marg"].dtype == m4.dtype and (r["marg"][:len(m4)] == m4).all()
    else:
        # Frames whose reference marginals went inf/NaN (|q| == 1 knife edges in 2*atanh(q), src/math_utils.py:56-60)
        # flood with NaN at an iteration that depends on the last ulp of tanh/log/exp: words must still agree,
        # iteration counts are only compared on the finite frames.
        nf = gold["nonfinite"]
        words = (r["x_hat"] == gold["x_hat"]).all(axis=1)
        its = r["iters"] == gold["iters"]
        assert (words & its)[~nf].all(), "SPA words / iteration counts differ on finite frames %s" % (
            np.flatnonzero(~(words & its) & ~nf)[:8])
        assert words[nf].mean() >= 0.9 if nf.any() else True
        m4 = gold["marg4"]
        fin = np.isfinite(m4) & ~nf[:len(m4), None]
        with np.errstate(invalid="ignore"):
            d = np.abs(r["marg"][:len(m4)] - m4)
        assert (d[fin] <= 50 * spa_tol(m4[fin])).all()   # marginals accumulate a few messages over <=100 iterations


@pytest.mark.parametrize("case", G.spa_tf(), ids=lambda c: c[0]["slot"])
def test_spa_teacher_forced(case):
    rec, v2c, c2v_ref = case
    g = graph(rec["code"])
    c2v = O.cn_sweep(g, O.SPA, v2c)
    fin = np.isfinite(c2v_ref)
    assert (np.isfinite(c2v) == fin).all() or (np.abs(c2v_ref[np.isfinite(c2v) != fin]) > 30).all()
    both = fin & np.isfinite(c2v)
    assert (np.abs(c2v[both] - c2v_ref[both]) <= spa_tol(c2v_ref[both])).all()
    assert (np.sign(c2v[both]) == np.sign(c2v_ref[both])).all()


def test_llr_front_ends_match_numpy_expressions():
    rng = np.random.RandomState(5)
    y = rng.normal(size=1000)
    for snr in (0.1, 1.0, 2.5):
        nv = 10 ** (-snr / 10)
        assert (O.llr_biawgn(snr, y) == -2 * y / nv).all()
    yb = (rng.random_sample(1000) < .3).astype(np.uint8)
    for p in (1 / 3, .051, .1):
        llr = np.log(1 - p) - np.log(p)
        assert (O.llr_bsc(p, yb) == llr * (1 - 2 * yb.astype(np.int64))).all()


def test_unlimited_iterations_need_cap_and_flag_it():
    g = graph("1200_3_6_rand_ldpc_1")
    x = np.ones(g.n, np.int64)
    Y = G.channel_send("biawgn", 0.5, np.tile(x, (4, 1)), 7)
    r = O.bp_decode(g, O.MSA, O.llr_biawgn(0.5, Y), max_iter=0, iter_cap=25)
    assert (r["iters"] == 25).all() and (r["reason"] == 4).all()
