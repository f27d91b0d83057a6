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
    hdrs = [os.path.join(HERE, "..", "ldpc_decoders_b200", "csrc", h) for h in ("ldpc_math.cuh", "res_layout.h", "channel_gen.cuh")]
    hdrs.append(os.path.join(EMU_DIR, "spa_phi_alt.h"))
    if not os.path.exists(so) or os.path.getmtime(so) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
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


def cn_sc(emu, g, v32, sat=38.1230):
    out = np.zeros_like(v32)
    emu.emu_cn_sc(g.n, g.m, g.E, ptr(g.chk_ptr), ptr(g.edge_var), ptr(v32), ptr(out), ctypes.c_float(sat))
    return out


@pytest.mark.parametrize("rule", ["sc", "phi"])
def test_spa_f32_teacher_forced(emu, rule):
    """float32 check node (the kernels' hyperbolic-pair rule, and the phi-domain form it replaced) vs the float64
    reference formula on the same (f32-rounded) inputs:
    |d| <= 1e-4 * max(1,|ref|) for |ref| < 20 (north_star tolerance); sign + large magnitude beyond."""
    worst = 0.0
    for rec, v2c, _ in G.spa_tf():
        g = graph(rec["code"])
        v32 = np.ascontiguousarray(v2c, np.float32)
        ref = O.cn_sweep(g, O.SPA, v32.astype(np.float64))
        if rule == "sc":
            out = cn_sc(emu, g, v32)
        else:
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


def test_spa_f32_degenerate_inputs(emu):
    g = graph("7_4_hamming")
    v = np.array([0.0, 1.0, -2.0, 3.0,   50.0, 60.0, -70.0, 0.5,   1e-6, -1e-6, 30., 2.], np.float32)
    out = cn_sc(emu, g, v, sat=np.inf)                      # saturation emulation off
    ref = O.cn_sweep(g, O.SPA, v.astype(np.float64))
    assert np.isnan(out[0]) and np.isnan(ref[0])            # v == 0: 0/0 on its own edge (bpa.py:74)
    assert (out[1:4] == 0).all() and (ref[1:4] == 0).all()  # and 0 on the others
    # all OTHER inputs saturated: the float64 formula overflows to -inf (tanh == 1 beyond |v| = 38); without the
    # saturation emulation the float32 rule returns the true value ~ -50, with it (the default) -inf like the reference
    assert ref[7] == -np.inf and -51 < out[7] < -49
    assert cn_sc(emu, g, v)[7] == -np.inf
    keep = [4, 5, 6, 8, 9, 10, 11]
    np.testing.assert_allclose(out[keep], ref[keep], rtol=1e-5, atol=1e-7)     # stated tolerance: 1e-4 * max(1, |ref|)
    # degree-6 checks take the step-only tree: a zero input still zeroes the others exactly and NaNs its own edge
    g6 = graph("1200_3_6_rand_ldpc_1")
    v6 = np.random.RandomState(5).normal(size=g6.E).astype(np.float32) * 4
    zeros = np.arange(0, g6.E, 7)
    v6[zeros] = 0.0
    o6 = cn_sc(emu, g6, v6)
    r6 = O.cn_sweep(g6, O.SPA, v6.astype(np.float64))
    assert (np.isnan(o6) == np.isnan(r6)).all() and np.isnan(o6[zeros]).all()
    assert ((o6 == 0) == (r6 == 0)).all()
    ok = np.isfinite(r6)
    assert (np.abs(o6[ok] - r6[ok]) <= 1e-5 * np.maximum(1, np.abs(r6[ok]))).all()


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


# --------------------------------------------------------------------------------------------- on-chip formulation
@pytest.mark.parametrize("channel,code,param,cw", [("biawgn", "1200_3_6_rand_ldpc_1", 2.0, 1),
                                                   ("biawgn", "1200_3_6_rand_ldpc_1", 2.6, 0),
                                                   ("bsc", "1200_rho_x5_rand_ldpc_10", .045, 0),
                                                   ("bsc", "1200_3_6_rand_ldpc_1", .05, 1),
                                                   ("biawgn", "7_4_hamming", 1.0, 1),
                                                   ("biawgn", "512_3_6_rand_ldpc_1", 2.0, 0)])
@pytest.mark.parametrize("variant", ["check_major", "variable_plane"])
def test_resident_formulation_matches_oracle(emu, channel, code, param, cw, variant):
    """The on-chip kernel's restructured decode (marginal gathers, v2c = marg - c2v_old never stored, syndrome from the
    sign bits of marg, lean 3-input-min min-sum, -0.0 priors folded, graph positions and per-check edge order from
    res_layout.h) gives the oracle's words, iteration counts and exit reasons bit for bit."""
    g = graph(code)
    frames = 300
    x = np.zeros(g.n, np.int64) + cw
    Y = G.channel_send(channel, param, np.tile(x, (frames, 1)), 2024)
    if channel == "bsc":
        yh = np.ascontiguousarray(Y, np.uint8)
        pri = O.llr_bsc(param, yh).astype(np.float32)
        yh[0] = cw                                   # a clean word: iteration-0 exit
        pri[0] = O.llr_bsc(param, yh[:1]).astype(np.float32)[0]
    else:
        yh = None
        pri = O.llr_biawgn(param, Y).astype(np.float32)
        pri[1, :5] = [0.0, -0.0, 0.0, -0.0, 1.0]     # zero and negative-zero priors
    ref = O.bp_decode(g, O.MSA, pri, y_hard=yh, max_iter=10, nthreads=4)
    x_hat = np.zeros((frames, g.n), np.uint8)
    iters = np.zeros(frames, np.int32)
    dec = np.zeros(frames, np.uint8)
    # variable_plane = csrc/resident_vp.cuh: messages stored by (edge rank at the variable, variable position), the
    # ordered sum starts from c0 instead of 0 + c0, the word is the sign bits of marg
    fn = emu.emu_resident_msa if variant == "check_major" else emu.emu_resident_vp_msa
    fn(g.n, g.m, g.E, ptr(g.chk_ptr), ptr(g.edge_var), ptr(g.var_ptr), ptr(g.var_edges),
       ctypes.c_double(0.05), frames, ptr(pri), ptr(yh), 10, ptr(x_hat), ptr(iters), ptr(dec))
    assert (iters == ref["iters"]).all()
    assert (x_hat == ref["x_hat"]).all()
    assert ((dec == 1) == (ref["reason"] == 0)).all()
    if channel == "bsc":
        assert iters[0] == 0


@pytest.mark.parametrize("channel,code,param", [("bsc", "1200_rho_x5_rand_ldpc_10", .045), ("biawgn", "1200_rho_x5_rand_ldpc_7", 2.0),
                                                ("biawgn", "7_4_hamming", 1.0), ("bsc", "12_3_4_ldpc", .08),
                                                ("biawgn", "4_2_test", 1.0), ("biawgn", "6_2_3_ldpc", 1.0),
                                                ("biawgn", "1200_3_6_rand_ldpc_2", 2.0)])
def test_irregular_variable_plane_tables_match_oracle(emu, channel, code, param):
    """resident_vp.cuh with IRR = true, on the tables build_vx_tables hands the kernel: degree-sorted positions, planes
    that are prefixes, checks padded to six edges (+inf reads, scratch writes), lean min-sum on the padded check.
    Words, iteration counts and exit reasons are the oracle's, bit for bit; the planes take E cells plus a few."""
    g = graph(code)
    frames = 300
    Y = G.channel_send(channel, param, np.zeros((frames, g.n), np.int64), 2027)
    if channel == "bsc":
        yh = np.ascontiguousarray(Y, np.uint8)
        pri = O.llr_bsc(param, yh).astype(np.float32)
        yh[0] = 0                                    # a clean word: iteration-0 exit
        pri[0] = O.llr_bsc(param, yh[:1]).astype(np.float32)[0]
    else:
        yh = None
        pri = O.llr_biawgn(param, Y).astype(np.float32)
        pri[1, :4] = [0.0, -0.0, 0.0, 1.0]
    ref = O.bp_decode(g, O.MSA, pri, y_hard=yh, max_iter=10, nthreads=4)
    x_hat = np.zeros((frames, g.n), np.uint8)
    iters = np.zeros(frames, np.int32)
    dec = np.zeros(frames, np.uint8)
    info = np.zeros(9, np.int32)
    rc = emu.emu_resident_vx_msa(g.n, g.m, g.E, ptr(g.chk_ptr), ptr(g.edge_var), ptr(g.var_ptr), ptr(g.var_edges),
                                 ctypes.c_double(0.05), frames, ptr(pri), ptr(yh), 10, ptr(x_hat), ptr(iters), ptr(dec), ptr(info))
    assert rc == 0
    assert (iters == ref["iters"]).all()
    assert (x_hat == ref["x_hat"]).all()
    assert ((dec == 1) == (ref["reason"] == 0)).all()
    assert g.E <= info[0] <= 1.08 * g.E + 64 and (np.diff(info[1:]) <= 0).all() and (info[1:] % 8 == 0).all()


@pytest.mark.parametrize("code", ["1200_3_6_rand_ldpc_1", "1200_rho_x5_rand_ldpc_10", "7_4_hamming", "12_3_4_ldpc"])
def test_shared_memory_placement(emu, code):
    """res_layout.h: positions are permutations into a padded range, every check's edge planes are a permutation of
    0..dc-1, and the predicted gather wavefronts only go down (to near the ideal for the (3,6) code's check phase)."""
    g = graph(code)
    stats = (ctypes.c_long * 9)()
    cpos, vpos, eord = np.zeros(g.m, np.int32), np.zeros(g.n, np.int32), np.zeros(g.E, np.uint8)
    emu.emu_plan(g.n, g.m, g.E, ptr(g.chk_ptr), ptr(g.edge_var), ptr(g.var_ptr), ptr(g.var_edges), 8,
                 ctypes.c_double(0.4), stats, ptr(cpos), ptr(vpos), ptr(eord))
    cn_ideal, cn_file, cn_nat, cn_plan, vn_ideal, vn_file, vn_plan, mp, npos = list(stats)
    assert mp % 8 == 0 and npos % 8 == 0 and mp >= g.m and npos >= g.n and mp < g.m + 8 and npos < g.n + 8
    assert len(set(cpos.tolist())) == g.m and cpos.min() >= 0 and cpos.max() < mp
    assert len(set(vpos.tolist())) == g.n and vpos.min() >= 0 and vpos.max() < npos
    for c in range(g.m):
        e0, e1 = g.chk_ptr[c], g.chk_ptr[c + 1]
        assert sorted(eord[e0:e1].tolist()) == list(range(e1 - e0))
    assert cn_ideal <= cn_plan <= cn_file and vn_ideal <= vn_plan <= vn_file and cn_plan <= cn_nat
    if code == "1200_3_6_rand_ldpc_1":
        assert cn_file > 2.3 * cn_ideal and vn_file > 2.3 * vn_ideal        # file order: random gathers
        assert cn_plan < 1.35 * cn_ideal and vn_plan < 2.0 * vn_ideal


@pytest.mark.parametrize("natural", [False, True])
def test_shared_memory_placement_variable_plane(emu, natural):
    """res_layout.h with vn_contiguous (resident_vp.cuh): the variable phase is contiguous whatever the placement, only
    the check gathers (and the scatters, same bank groups) count.  Min-sum may order the edges of a check: the balanced
    colouring + matching decomposition is conflict-free or within a few wavefronts of it.  Sum-product keeps the natural
    order (8-cliques against 8 colours): well below file order, not ideal."""
    g = graph("1200_3_6_rand_ldpc_1")
    stats = (ctypes.c_long * 9)()
    cpos, vpos, eord = np.zeros(g.m, np.int32), np.zeros(g.n, np.int32), np.zeros(g.E, np.uint8)
    fn = emu.emu_plan_vp_natural if natural else emu.emu_plan_vp
    fn(g.n, g.m, g.E, ptr(g.chk_ptr), ptr(g.edge_var), ptr(g.var_ptr), ptr(g.var_edges), 8,
       ctypes.c_double(0.4), stats, ptr(cpos), ptr(vpos), ptr(eord))
    cn_ideal, cn_file, cn_nat, cn_plan, vn_ideal, vn_file, vn_plan, mp, npos = list(stats)
    assert sorted(cpos.tolist()) == list(range(g.m)) and sorted(vpos.tolist()) == list(range(g.n))
    for c in range(g.m):
        e0, e1 = g.chk_ptr[c], g.chk_ptr[c + 1]
        assert sorted(eord[e0:e1].tolist()) == list(range(e1 - e0))
    assert vn_plan == vn_ideal and cn_ideal <= cn_plan <= cn_nat <= cn_file
    if natural:
        assert (eord == np.concatenate([np.arange(g.chk_ptr[c + 1] - g.chk_ptr[c]) for c in range(g.m)])).all()
        assert cn_plan == cn_nat < 0.6 * cn_file
    else:
        assert cn_plan <= 1.03 * cn_ideal


def test_biawgn_llr_without_division_is_exact(emu):
    """llr_biawgn_f32 == float32((-2 y) / noise_var) for every input, the division being evaluated for ~2^-26 of them."""
    rng = np.random.RandomState(5)
    total_slow = 0
    for snr in (0.01, 1.0, 2.0, 2.3, 3.0, 6.0):
        nv = 10 ** (-snr / 10)
        y = (1 + np.sqrt(nv) * rng.standard_normal(2_000_000)).astype(np.float32).astype(np.float64)
        y[:8] = [0.0, -0.0, 1e-42, -1e-40, 1e30, -3e38, np.inf, 1.0]
        y[8:200008] = rng.standard_normal(200000) * 10.0 ** rng.randint(-40, 38, 200000)
        fast, exact = np.zeros(y.size, np.float32), np.zeros(y.size, np.float32)
        slow = ctypes.c_int64()
        emu.emu_llr_biawgn(ctypes.c_size_t(y.size), ptr(y), ctypes.c_double(nv), ptr(fast), ptr(exact), ctypes.byref(slow))
        with np.errstate(over="ignore"):
            ref = (-2 * y / nv).astype(np.float32)
        assert (fast.view(np.uint32) == ref.view(np.uint32)).all()
        assert (exact.view(np.uint32) == ref.view(np.uint32)).all()
        total_slow += slow.value
    assert total_slow < 120000          # the out-of-range test values, plus a handful of boundary cases


# --------------------------------------------------------------------------------------------- on-device channel noise
def test_philox_known_answers(emu):
    """Philox4x32-10 against the Random123 known-answer vectors (kat_vectors)."""
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        c, k, o = np.array(ctr, np.uint32), np.array(key, np.uint32), np.zeros(4, np.uint32)
        emu.emu_philox(ptr(c), ptr(k), ptr(o))
        assert tuple(int(v) for v in o) == want


def test_box_muller_normals(emu):
    """The BIAWGN noise generator: standard normal moments, tails and independence between frames / variables."""
    frames, n = 400, 1200
    z = np.zeros((frames, n), np.float32)
    emu.emu_normals(ctypes.c_ulonglong(7), ctypes.c_ulonglong(10 ** 12), frames, n, ptr(z))
    N = z.size
    assert abs(z.mean()) < 4 / np.sqrt(N) and abs(z.var() - 1) < 4 * np.sqrt(2 / N)
    assert abs((z ** 3).mean()) < 4 * np.sqrt(15 / N) and abs((z ** 4).mean() - 3) < 4 * np.sqrt(96 / N)
    for t, p in ((1.0, 0.31731), (2.0, 0.0455), (3.0, 0.0027)):
        frac = (np.abs(z) > t).mean()
        assert abs(frac - p) < 4 * np.sqrt(p / N)
    assert abs(np.corrcoef(z[:-1].ravel(), z[1:].ravel())[0, 1]) < 4 / np.sqrt(N)       # frame f vs f+1
    assert abs(np.corrcoef(z[:, :-1].ravel(), z[:, 1:].ravel())[0, 1]) < 4 / np.sqrt(N)  # variable v vs v+1
    z2 = np.zeros((10, n), np.float32)
    emu.emu_normals(ctypes.c_ulonglong(7), ctypes.c_ulonglong(10 ** 12 + 5), 10, n, ptr(z2))
    assert (z2 == z[5:15]).all()                      # keyed by the global frame index only


def test_bec_degree3_variable_rule_is_exhaustively_the_integer_form(emu):
    """ldpc::bec_vn3 (boolean leave-one-out form of bec.py:115-119 for degree-3 variables) == the bit-sliced integer
    form on all 3^4 ternary inputs, 32 random assignments of them to bit lanes at a time."""
    import itertools
    combos = list(itertools.product((-1, 0, 1), repeat=4))
    rng = np.random.RandomState(0)
    for rep in range(8):
        pick = [combos[i] for i in rng.permutation(len(combos))[:32]] if rep else combos[:32]
        if rep == 1:
            pick = combos[32:64]
        if rep == 2:
            pick = combos[49:81]
        nz = np.zeros(4, np.uint32); pos = np.zeros(4, np.uint32)
        for lane, c in enumerate(pick):
            for i, t in enumerate(c):
                if t != 0:
                    nz[i] |= np.uint32(1 << lane)
                if t > 0:
                    pos[i] |= np.uint32(1 << lane)
        fast = np.zeros(8, np.uint32); ref = np.zeros(8, np.uint32)
        emu.emu_bec_vn3(ptr(nz), ptr(pos), ptr(fast), ptr(ref))
        assert (fast == ref).all(), (rep, fast, ref)
        for lane, c in enumerate(pick):                    # and against plain integers
            S = sum(c)
            assert ((int(fast[6]) >> lane) & 1, (int(fast[7]) >> lane) & 1) == (int(S != 0), int(S > 0))
            for e in range(3):
                Se = S - c[e + 1]
                assert ((int(fast[2 * e]) >> lane) & 1, (int(fast[2 * e + 1]) >> lane) & 1) == (int(Se != 0), int(Se > 0))


def test_bec_degree6_check_tree_equals_sequential_accumulator(emu):
    """ldpc::BecCn6 (erasure count / parity as reduction trees) == BecCnAccT (bec.py:100-112) on random ternary inputs,
    including every erasure count 0..6 in some lane."""
    rng = np.random.RandomState(1)
    for rep in range(200):
        pe = rng.choice([0.0, 0.1, 0.3, 0.6, 1.0])
        er = rng.rand(6, 32) < pe
        if rep < 7:                                         # lanes with exactly rep erasures
            er[:] = False
            er[:rep] = True
        posb = (rng.rand(6, 32) < .5) & ~er
        nz = np.array([sum((0 if er[k, l] else 1) << l for l in range(32)) for k in range(6)], np.uint32)
        pos = np.array([sum(int(posb[k, l]) << l for l in range(32)) for k in range(6)], np.uint32)
        fast = np.zeros(12, np.uint32); ref = np.zeros(12, np.uint32)
        emu.emu_bec_cn6(ptr(nz), ptr(pos), ptr(fast), ptr(ref))
        assert (fast == ref).all(), rep


def test_bec_conflict_free_variable_rules_equal_the_literal_ones(emu):
    """bec_vn3_or / BecVnOr (ORs of the other inputs) == the bit-sliced integer rule of bec.py:115-119 on every lane
    whose inputs carry no conflicting votes, for every degree 0..8; lanes WITH conflicts are flagged by the conflict
    word (the kernel then takes the literal rule).  Exhaustive for degree 3."""
    import itertools
    emu.emu_bec_vn_or.restype = ctypes.c_uint32
    rng = np.random.RandomState(2)
    for d in range(0, 9):
        combos = list(itertools.product((-1, 0, 1), repeat=d + 1)) if d <= 3 else None
        reps = (len(combos) + 31) // 32 if combos else 40
        for rep in range(reps):
            if combos:
                pick = combos[32 * rep:32 * rep + 32]
            else:
                pick = [tuple(rng.choice([-1, 0, 1], p=[.3, .4, .3]) if rng.rand() < .5 else abs(rng.choice([-1, 0, 1])) * rng.choice([-1, 1])
                              for _ in range(d + 1)) for _ in range(32)]
                pick = [tuple(int(abs(t)) * (1 if rng.rand() < .5 else -1) for t in c) if l % 4 == 0 else
                        tuple(int(abs(t)) * s0 for t in c) for l, (c, s0) in enumerate(zip(pick, rng.choice([-1, 1], size=32)))]
            nz = np.zeros(9, np.uint32); pos = np.zeros(9, np.uint32)
            for lane, c in enumerate(pick):
                for i, t in enumerate(c):
                    if t != 0:
                        nz[i] |= np.uint32(1 << lane)
                    if t > 0:
                        pos[i] |= np.uint32(1 << lane)
            f3 = np.zeros(20, np.uint32); fg = np.zeros(20, np.uint32); ref = np.zeros(20, np.uint32)
            conflict = emu.emu_bec_vn_or(d, ptr(nz), ptr(pos), ptr(f3), ptr(fg), ptr(ref))
            assert conflict != 0xdeadbeef
            for lane, c in enumerate(pick):
                has_conflict = any(t > 0 for t in c) and any(t < 0 for t in c)
                assert ((conflict >> lane) & 1) == int(has_conflict)
            ok = np.uint32(~np.uint32(conflict) & np.uint32((1 << len(pick)) - 1 if len(pick) < 32 else 0xffffffff))
            assert ((fg[:2 * d + 2] ^ ref[:2 * d + 2]) & ok).max() == 0, (d, rep)
            if d == 3:
                assert ((f3[:8] ^ ref[:8]) & ok).max() == 0, rep
