"""CPU pre-flight of the product's node arithmetic (csrc/ldpc_math.cuh).

tests/host_emu/emu.cpp compiles the very __host__ __device__ functions the CUDA kernels call and
drives them with plain loops (same book-keeping as the kernels); here they are compared with the
pinned oracle.  This is a development aid that runs without a GPU — the GPU tests in
test_gpu_parity.py remain the parity tests proper.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import _golden as G
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "host_emu")


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU_DIR, "libhost_emu.so")
    src = os.path.join(EMU_DIR, "emu.cpp")
    hdr = os.path.join(HERE, "..", "ldpc_decoders_b200", "csrc", "ldpc_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src],
                       check=True, cwd=EMU_DIR)
    return ctypes.CDLL(so)


def graph(name):
    return O.Graph(*G.code_tables(name))


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def emu_bp(emu, g, algo, priors, y_hard, max_iter):
    B = priors.shape[0]
    x_hat = np.zeros((B, g.n), np.uint8)
    iters = np.zeros(B, np.int32)
    marg = np.zeros_like(priors)
    fn = emu.emu_bp_f64 if priors.dtype == np.float64 else emu.emu_bp_f32
    fn(algo, g.n, g.m, g.E, ptr(g.chk_ptr), ptr(g.edge_var), ptr(g.var_ptr), ptr(g.var_edges), B,
       ptr(priors), ptr(y_hard), max_iter, ptr(x_hat), ptr(iters), ptr(marg))
    return x_hat, iters, marg


BP_RUNS = [r for r in G.runs() if r["channel"] != "bec" and r["max_iter"] <= 40]


@pytest.mark.parametrize("rec", BP_RUNS, ids=lambda r: r["key"])
def test_bp_math_matches_oracle(emu, rec):
    g = graph(rec["code"])
    x, Y = G.run_inputs(rec)
    Y = Y[:64]
    dt = np.float64 if rec["dtype"] == "f64" else np.float32
    if rec["channel"] == "bsc":
        pri, yh = O.llr_bsc(rec["param"], Y.astype(np.uint8)).astype(dt), np.ascontiguousarray(Y, np.uint8)
    else:
        pri, yh = O.llr_biawgn(rec["param"], Y).astype(dt), None
    algo = O.MSA if rec["decoder"] == "MSA" else O.SPA
    ref = O.bp_decode(g, algo, pri, y_hard=yh, max_iter=rec["max_iter"], want_marg=True)
    x_hat, iters, marg = emu_bp(emu, g, algo, np.ascontiguousarray(pri), yh, rec["max_iter"])
    if algo == O.MSA or dt == np.float64:
        # MSA: exact in any IEEE type.  SPA f64 mirror: same formula, same libm on the host.
        assert (iters == ref["iters"]).all()
        assert (x_hat == ref["x_hat"]).all()
        same = (marg == ref["marg"]) | (np.isnan(marg) & np.isnan(ref["marg"]))
        assert same.all()


@pytest.mark.parametrize("rec", [r for r in BP_RUNS if r["decoder"] == "MSA" and r["dtype"] == "f32"], ids=lambda r: r["key"])
def test_msa_bit_pattern_variant_is_identical(emu, rec):
    """cn_msa_bits (integer min/max on float32 bit patterns, used by the resident kernel) == cn_msa<float>."""
    g = graph(rec["code"])
    x, Y = G.run_inputs(rec)
    Y = Y[:96]
    if rec["channel"] == "bsc":
        pri, yh = O.llr_bsc(rec["param"], Y.astype(np.uint8)).astype(np.float32), np.ascontiguousarray(Y, np.uint8)
    else:
        pri, yh = O.llr_biawgn(rec["param"], Y).astype(np.float32), None
    ref = O.bp_decode(g, O.MSA, pri, y_hard=yh, max_iter=rec["max_iter"], want_marg=True)
    x_hat, iters, marg = emu_bp(emu, g, 2, np.ascontiguousarray(pri), yh, rec["max_iter"])
    assert (iters == ref["iters"]).all() and (x_hat == ref["x_hat"]).all() and (marg == ref["marg"]).all()
    rng = np.random.RandomState(0)                        # zeros, negative zeros, ties, infinities
    v = rng.choice(np.array([0.0, -0.0, 1.5, -1.5, 2.0, np.inf, -np.inf, 1e-38, -3.25], np.float32), size=g.E)
    a = O.cn_sweep(g, O.MSA, v)
    b = np.zeros_like(v)
    emu.emu_cn_msa_bits(g.n, g.m, g.E, ptr(g.chk_ptr), ptr(g.edge_var), ptr(v), ptr(b))
    assert (a.view(np.uint32) == b.view(np.uint32)).all()


def test_spa_phi_f32_teacher_forced(emu):
    """float32 phi-domain check node vs the float64 reference formula on the same (f32-rounded) inputs:
    |d| <= 1e-4 * max(1,|ref|) for |ref| < 20 (north_star tolerance); sign + large magnitude beyond."""
    worst = 0.0
    for rec, v2c, _ in G.spa_tf():
        g = graph(rec["code"])
        v32 = np.ascontiguousarray(v2c, np.float32)
        ref = O.cn_sweep(g, O.SPA, v32.astype(np.float64))
        out = np.zeros_like(v32)
        emu.emu_cn_phi(g.n, g.m, g.E, ptr(g.chk_ptr), ptr(g.edge_var), ptr(v32), ptr(out))
        a = np.abs(ref)
        m = a < 20
        err = np.abs(out[m] - ref[m]) / np.maximum(1, a[m])
        worst = max(worst, float(err.max()))
        assert (err <= 1e-4).all()
        big = ~m & np.isfinite(ref)
        assert (np.sign(out[big]) == np.sign(ref[big])).all() and (np.abs(out[big]) >= 15).all()
    assert worst < 5e-6        # measured ~6e-7; the stated tolerance is 1e-4


def test_spa_phi_degenerate_inputs(emu):
    g = graph("7_4_hamming")
    v = np.array([0.0, 1.0, -2.0, 3.0,   50.0, 60.0, -70.0, 0.5,   1e-6, -1e-6, 30., 2.], np.float32)
    out = np.zeros_like(v)
    emu.emu_cn_phi(g.n, g.m, g.E, ptr(g.chk_ptr), ptr(g.edge_var), ptr(v), ptr(out))
    ref = O.cn_sweep(g, O.SPA, v.astype(np.float64))
    assert np.isnan(out[0]) and np.isnan(ref[0])            # v == 0: 0/0 on its own edge (bpa.py:74)
    assert (out[1:4] == 0).all() and (ref[1:4] == 0).all()  # and 0 on the others
    # all OTHER inputs saturated: the float64 formula overflows to -inf (tanh == 1 beyond |v| = 38),
    # the phi form returns the true value ~ -50; both count as "certain" under the stated metric
    assert ref[7] == -np.inf and out[7] < -15
    keep = [4, 5, 6, 8, 9, 10, 11]
    np.testing.assert_allclose(out[keep], ref[keep], rtol=1e-5, atol=1e-12)


BEC_RUNS = [r for r in G.runs() if r["channel"] == "bec"]


@pytest.mark.parametrize("rec", BEC_RUNS, ids=lambda r: r["key"])
def test_bec_bitplanes_match_oracle(emu, rec):
    g = graph(rec["code"])
    x, Y = G.run_inputs(rec)
    Y = np.ascontiguousarray(Y[:96], np.uint8)
    ref = O.bec_decode(g, Y, max_iter=rec["max_iter"])
    B = Y.shape[0]
    nb = 5 if np.diff(g.var_ptr).max() <= 14 else 8
    x_hat, iters, reason = np.zeros((B, g.n), np.uint8), np.zeros(B, np.int32), np.zeros(B, np.uint8)
    emu.emu_bec(g.n, g.m, g.E, ptr(g.chk_ptr), ptr(g.edge_var), ptr(g.var_ptr), ptr(g.var_edges), B, ptr(Y),
                rec["max_iter"], nb, ptr(x_hat), ptr(iters), ptr(reason))
    assert (iters == ref["iters"]).all()
    assert (reason == ref["reason"]).all()
    assert (x_hat == ref["x_hat"]).all()


@pytest.mark.parametrize("nb", [5, 8])
def test_bec_bitplanes_arbitrary_symbols(emu, nb):
    """Inputs that are NOT a codeword plus erasures (conflicting votes): the literal integer message
    passing must still be reproduced bit for bit."""
    rng = np.random.RandomState(11)
    for name in ("7_4_hamming", "12_3_4_ldpc", "1200_rho_x5_rand_ldpc_10"):
        g = graph(name)
        B = 70
        Y = rng.choice(3, size=(B, g.n), p=[.35, .35, .3]).astype(np.uint8)
        for mi in (1, 3, 10, 0):
            ref = O.bec_decode(g, Y, max_iter=mi)
            x_hat, iters, reason = np.zeros((B, g.n), np.uint8), np.zeros(B, np.int32), np.zeros(B, np.uint8)
            emu.emu_bec(g.n, g.m, g.E, ptr(g.chk_ptr), ptr(g.edge_var), ptr(g.var_ptr), ptr(g.var_edges), B, ptr(Y),
                        mi, nb, ptr(x_hat), ptr(iters), ptr(reason))
            assert (iters == ref["iters"]).all() and (reason == ref["reason"]).all() and (x_hat == ref["x_hat"]).all()
