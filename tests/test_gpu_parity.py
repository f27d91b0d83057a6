"""GPU tier (-m gpu): the CUDA path, called through the C ABI, against the oracle and the golden fixtures.

Parity bars:
  BEC          x_hat, iteration count and exit reason bit-exact
  MSA          x_hat, iteration count (and marginals) bit-exact against the oracle AT THE SAME DTYPE
  SPA float64  formula mirror: identical words / iteration counts on frames whose reference marginals stay
               finite; messages within 1e-12*max(1,|ref|) + 1e-15*exp(|ref|) (2*atanh conditioning, H3)
  SPA float32  hyperbolic-pair rule (ldpc_math.cuh cn_spa_sc): |d| <= 1e-4*max(1,|ref|) for |ref| < 20 against the float64 formula on the same
               inputs, sign agreement + |LLR| >= 15 beyond; identical words on frames away from a decision
               boundary
"""
import numpy as np
import pytest

import _golden as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    from ldpc_decoders_b200 import Tables, _lib, bec, biawgn, bpa, bsc, engine, models
    torch.cuda.set_device(0)
    return dict(torch=torch, Tables=Tables, lib=_lib, bec=bec, biawgn=biawgn, bpa=bpa, bsc=bsc,
                engine=engine, models=models)


_tabs = {}


def tables(mods, name):
    if name not in _tabs:
        _tabs[name] = mods["Tables"](*G.code_tables(name))
    return _tabs[name]


def ograph(name):
    return O.Graph(*G.code_tables(name))


def spa_tol(ref):
    a = np.abs(ref)
    return 1e-12 * np.maximum(1.0, a) + 1e-15 * np.exp(np.minimum(a, 40.0))


class Code:
    """What the reference hands to a decoder constructor: an object with a dense .parity_mtx (codes.Code)."""

    def __init__(self, name):
        self.parity_mtx = G.dense_H(name)


# --------------------------------------------------------------------------------------------- KATs
@pytest.mark.parametrize("k", G.kats(), ids=lambda k: "%s-%s-%s" % (k["channel"], k["code"], k["decoder"]))
def test_kat_through_reference_protocol(mods, k):
    """The six Test.sample vectors (src/bec.py:132-139, src/bsc.py:82-89, src/biawgn.py:85-92), decoded the
    way utils.TestCase.sample does it: decoder(param, code, **kwargs).decode(y)  (src/utils.py:84)."""
    model = mods["models"].models[k["channel"]]
    dec = getattr(model, k["decoder"])(k["param"], Code(k["code"]), max_iter=k["max_iter"], mu=3., eps=1e-5)
    y = np.array(k["y"]) if k["channel"] == "biawgn" else np.array(k["y"]).astype(int)
    est = dec.decode(y)
    assert (np.asarray(est) == np.array(k["x"])).all()
    assert np.asarray(est).tolist() == k["x_hat"]
    st = dec.stats()
    assert st["iter"][k["iters"]] == 1 and sum(st["iter"]) == 1


# --------------------------------------------------------------------------------------------- golden runs
@pytest.mark.parametrize("rec", G.runs(), ids=lambda r: r["key"])
def test_golden_run(mods, rec):
    x, Y = G.run_inputs(rec)
    gold = G.run_arrays(rec)
    tab = tables(mods, rec["code"])
    model = mods["models"].models[rec["channel"]]
    if rec["channel"] == "bec":
        dec = model.SPA(rec["param"], tab, max_iter=rec["max_iter"])
        x_hat, iters, reason = dec.decode_batch(Y, return_reason=True)
        assert (iters == gold["iters"]).all() and (reason == gold["reason"]).all() and (x_hat == gold["x_hat"]).all()
        return
    dt = np.float64 if rec["dtype"] == "f64" else np.float32
    dec = getattr(model, rec["decoder"])(rec["param"], tab, max_iter=rec["max_iter"], dtype=dt)
    x_hat, iters, reason = dec.decode_batch(Y.astype(np.uint8) if rec["channel"] == "bsc" else Y, return_reason=True)
    if rec["decoder"] == "MSA":
        assert (iters == gold["iters"]).all()
        assert (x_hat == gold["x_hat"]).all()
        assert (reason == gold["reason"]).all()
    else:
        nf = gold["nonfinite"]
        ok = (iters == gold["iters"]) & (x_hat == gold["x_hat"]).all(axis=1)
        assert ok[~nf].all(), "SPA f64 differs on finite frames %s" % np.flatnonzero(~ok & ~nf)[:8]
        if nf.any():
            assert (x_hat == gold["x_hat"]).all(axis=1)[nf].mean() >= 0.9


@pytest.mark.parametrize("rec", [r for r in G.runs() if r["channel"] != "bec"], ids=lambda r: r["key"])
def test_golden_marginals(mods, rec):
    """Last marginal (src/bpa.py:35) of the first frames: bit-exact for MSA, conditioning-aware tolerance for SPA."""
    torch = mods["torch"]
    x, Y = G.run_inputs(rec)
    gold = G.run_arrays(rec)
    m4 = gold["marg4"]
    k = len(m4)
    og = ograph(rec["code"])
    dt = np.float64 if rec["dtype"] == "f64" else np.float32
    if rec["channel"] == "bsc":
        pri, yh = O.llr_bsc(rec["param"], Y[:k].astype(np.uint8)).astype(dt), np.ascontiguousarray(Y[:k], np.uint8)
    else:
        pri, yh = O.llr_biawgn(rec["param"], Y[:k]).astype(dt), None
    eng = mods["engine"].engine_for(tables(mods, rec["code"]))
    algo = mods["lib"].MSA if rec["decoder"] == "MSA" else mods["lib"].SPA
    out = eng.decode_device(algo, torch.from_numpy(pri).cuda(), None if yh is None else torch.from_numpy(yh).cuda(),
                            max_iter=rec["max_iter"], want_marg=True)
    marg = out["marg"].cpu().numpy()
    if rec["decoder"] == "MSA":
        assert marg.dtype == m4.dtype and (marg == m4).all()
    else:
        fin = np.isfinite(m4) & ~gold["nonfinite"][:k, None]
        with np.errstate(invalid="ignore"):
            d = np.abs(marg - m4)
        assert (d[fin] <= 50 * spa_tol(m4[fin])).all()


# --------------------------------------------------------------------------------------------- MSA / BEC at scale
@pytest.mark.parametrize("channel,code,param,cw,dt", [
    ("biawgn", "1200_3_6_rand_ldpc_1", 2.0, 1, np.float32),
    ("biawgn", "1200_3_6_rand_ldpc_1", 2.6, 1, np.float64),
    ("bsc", "1200_3_6_rand_ldpc_1", .05, 1, np.float32),
    ("bsc", "1200_3_6_rand_ldpc_1", .041, 1, np.float64),
    ("bsc", "1200_rho_x5_rand_ldpc_10", .045, 0, np.float32),
    ("biawgn", "1200_rho_x5_rand_ldpc_3", 1.5, 0, np.float32),
])
def test_msa_bit_exact_1e4_frames(mods, channel, code, param, cw, dt):
    frames = 10000
    tab, og = tables(mods, code), ograph(code)
    x = np.zeros(tab.n, np.int64) + cw
    Y = G.channel_send(channel, param, np.tile(x, (frames, 1)), 4242)
    dec = mods["models"].models[channel].MSA(param, tab, max_iter=10, dtype=dt)
    if channel == "bsc":
        Yh = Y.astype(np.uint8)
        ref = O.bp_decode(og, O.MSA, O.llr_bsc(param, Yh).astype(dt), y_hard=Yh, max_iter=10, nthreads=8)
        x_hat, iters = dec.decode_batch(Yh)
    else:
        ref = O.bp_decode(og, O.MSA, O.llr_biawgn(param, Y).astype(dt), max_iter=10, nthreads=8)
        x_hat, iters = dec.decode_batch(Y)
    assert (iters == ref["iters"]).all()
    assert (x_hat == ref["x_hat"]).all()


@pytest.mark.parametrize("code,p,mi", [("1200_3_6_rand_ldpc_1", .42, 10), ("1200_3_6_rand_ldpc_1", .38, 100),
                                        ("1200_rho_x5_rand_ldpc_10", .42, 0), ("margulis", .41, 100)])
def test_bec_bit_exact_many_frames(mods, code, p, mi):
    frames = 20000
    tab, og = tables(mods, code), ograph(code)
    Y = G.channel_send("bec", p, np.zeros((frames, tab.n), np.int64), 99).astype(np.uint8)
    ref = O.bec_decode(og, Y, max_iter=mi, nthreads=8)
    x_hat, iters, reason = mods["bec"].SPA(p, tab, max_iter=mi).decode_batch(Y, return_reason=True)
    assert (iters == ref["iters"]).all() and (reason == ref["reason"]).all() and (x_hat == ref["x_hat"]).all()


def test_bec_wide_and_narrow_sweeps_agree(mods):
    """Above ~129 k frames of an n = 1200 code the erasure sweeps take four flag words (128 frames) per thread
    (stream_bec.cuh, W4); below, one.  The same frames decoded in one large batch and in four small ones must give
    identical words, iteration counts and exit reasons (and the small batches are held to the oracle elsewhere)."""
    torch, lib = mods["torch"], mods["lib"]
    tab = tables(mods, "1200_3_6_rand_ldpc_1")
    eng = mods["engine"].engine_for(tab)
    B = 4 * 33000 + 17
    g = torch.Generator(device="cuda").manual_seed(11)
    r = torch.rand((B, tab.n), generator=g, device="cuda")
    y = torch.where(r < 0.41, 2, 0).to(torch.uint8)
    y[5] = 0
    y[B - 1, ::2] = 2
    big = eng.decode_device_channel(lib.CH_BEC, lib.BEC, lib.F32, 0.0, y, max_iter=100)
    big = {k: v.clone() for k, v in big.items() if v is not None}
    for lo in range(0, B, 33000):
        part = eng.decode_device_channel(lib.CH_BEC, lib.BEC, lib.F32, 0.0, y[lo:lo + 33000].contiguous(), max_iter=100)
        for k in ("x_hat", "iters", "reason"):
            assert bool((part[k] == big[k][lo:lo + 33000]).all()), k
    og = ograph("1200_3_6_rand_ldpc_1")
    ref = O.bec_decode(og, y[:3000].cpu().numpy(), max_iter=100, nthreads=8)
    assert (big["iters"][:3000].cpu().numpy() == ref["iters"]).all() and (big["x_hat"][:3000].cpu().numpy() == ref["x_hat"]).all()
    assert (big["reason"][:3000].cpu().numpy() == ref["reason"]).all()


def test_bec_arbitrary_symbols(mods):
    rng = np.random.RandomState(3)
    for code in ("7_4_hamming", "12_3_4_ldpc", "1200_rho_x5_rand_ldpc_10"):
        tab, og = tables(mods, code), ograph(code)
        Y = rng.choice(3, size=(333, tab.n), p=[.35, .35, .3]).astype(np.uint8)
        for mi in (1, 4, 10, 0):
            ref = O.bec_decode(og, Y, max_iter=mi)
            x_hat, iters, reason = mods["bec"].SPA(.3, tab, max_iter=mi).decode_batch(Y, return_reason=True)
            assert (iters == ref["iters"]).all() and (reason == ref["reason"]).all() and (x_hat == ref["x_hat"]).all()


def _code_names():
    import os
    z = np.load(os.path.join(G.GOLD, "codes.npz"))
    return sorted({k.rsplit("__", 1)[0] for k in z.files})


@pytest.mark.parametrize("code", _code_names())
def test_every_shipped_code_bit_exact(mods, code):
    """All 31 code files of data/codes (regular ensemble, irregular ensemble, n = 512, Margulis, toy codes): every code
    gets its own shared-memory placement (res_layout.h) and its own kernel instance, so each one is held to the oracle —
    min-sum float32 on BIAWGN and on BSC (hard input: iteration-0 exit), erasure decoding on BEC."""
    frames = 384
    tab, og = tables(mods, code), ograph(code)
    zeros = np.zeros((frames, tab.n), np.int64)
    small = tab.n < 100
    snr, p, pe = (1.0, .08, .3) if small else (2.0, .04, .4)
    Y = G.channel_send("biawgn", snr, zeros, 2024)
    ref = O.bp_decode(og, O.MSA, O.llr_biawgn(snr, Y).astype(np.float32), max_iter=10, nthreads=8)
    x_hat, iters, reason = mods["biawgn"].MSA(snr, tab, max_iter=10, dtype=np.float32).decode_batch(Y, return_reason=True)
    assert (iters == ref["iters"]).all() and (reason == ref["reason"]).all() and (x_hat == ref["x_hat"]).all()
    Yh = G.channel_send("bsc", p, zeros, 2025).astype(np.uint8)
    ref = O.bp_decode(og, O.MSA, O.llr_bsc(p, Yh).astype(np.float32), y_hard=Yh, max_iter=10, nthreads=8)
    x_hat, iters, reason = mods["bsc"].MSA(p, tab, max_iter=10, dtype=np.float32).decode_batch(Yh, return_reason=True)
    assert (iters == ref["iters"]).all() and (reason == ref["reason"]).all() and (x_hat == ref["x_hat"]).all()
    Ye = G.channel_send("bec", pe, zeros, 2026).astype(np.uint8)
    ref = O.bec_decode(og, Ye, max_iter=100, nthreads=8)
    x_hat, iters, reason = mods["bec"].SPA(pe, tab, max_iter=100).decode_batch(Ye, return_reason=True)
    assert (iters == ref["iters"]).all() and (reason == ref["reason"]).all() and (x_hat == ref["x_hat"]).all()


@pytest.mark.parametrize("mi", [1, 2, 3, 6, 10, 40, 100])
def test_max_iter_sweep_irregular_bsc(mods, mi):
    """BASELINE config 4 (src/simulations.py:77, IREG_ENS max_iter sweep on BSC): min-sum float32 bit-exact, float64
    sum-product identical words / iteration counts on frames whose reference marginals stay finite."""
    code, p, frames = "1200_rho_x5_rand_ldpc_4", .07, 512
    tab, og = tables(mods, code), ograph(code)
    Yh = G.channel_send("bsc", p, np.zeros((frames, tab.n), np.int64), 77 + mi).astype(np.uint8)
    pri = O.llr_bsc(p, Yh)
    ref = O.bp_decode(og, O.MSA, pri.astype(np.float32), y_hard=Yh, max_iter=mi, nthreads=8)
    x_hat, iters, reason = mods["bsc"].MSA(p, tab, max_iter=mi, dtype=np.float32).decode_batch(Yh, return_reason=True)
    assert (iters == ref["iters"]).all() and (reason == ref["reason"]).all() and (x_hat == ref["x_hat"]).all()
    ref = O.bp_decode(og, O.SPA, pri, y_hard=Yh, max_iter=mi, want_marg=True, nthreads=8)
    x_hat, iters = mods["bsc"].SPA(p, tab, max_iter=mi, dtype=np.float64).decode_batch(Yh)
    fin = np.isfinite(ref["marg"]).all(axis=1)
    assert fin.sum() > frames // 2
    assert (iters[fin] == ref["iters"][fin]).all() and (x_hat[fin] == ref["x_hat"][fin]).all()


@pytest.mark.parametrize("n,dv,dc", [(600, 3, 6), (1000, 3, 6), (96, 2, 4), (300, 3, 4), (640, 4, 8), (250, 3, 5), (1280, 3, 6), (2000, 3, 6), (1296, 3, 6), (2688, 3, 6), (2848, 3, 6), (3008, 3, 6)])
def test_random_regular_codes_fuzz(mods, n, dv, dc):
    """Seeded regular codes of several degree pairs and lengths: (3,6) runs on resident_vp (two CTAs per SM up to
    n = 1280, one up to n = 2848), check degrees up to 6 on its irregular instance, (4,8) on resident_bp.  Min-sum against the
    oracle at float32 (and float64 where the on-chip float64 kernel applies), sum-product on chip against streaming."""
    torch, lib = mods["torch"], mods["lib"]
    from ldpc_decoders_b200 import codes
    tab = codes.random_regular(n, dv, dc, seed=n + dc).tables
    og = O.Graph(tab.m, tab.n, tab.edge_chk.astype(np.int64), tab.edge_var.astype(np.int64))
    eng = mods["engine"].engine_for(tab)
    assert eng.resident_kernel == ("" if n > 2848 else "resident_bp" if dc > 6 else "resident_vp")     # 3008: streaming only
    B = 260
    Y = G.channel_send("biawgn", 3.0, np.zeros((B, tab.n), np.int64), 7000 + n)
    nv = 10 ** (-3.0 / 10)
    dtypes = [(np.float32, lib.F32)] + ([(np.float64, lib.F64)] if dc <= 6 and n <= 1280 else [])
    for dt, ldt in dtypes:
        ref = O.bp_decode(og, O.MSA, O.llr_biawgn(3.0, Y).astype(dt), max_iter=15, nthreads=4)
        out = eng.decode_device_channel(lib.CH_BIAWGN, lib.MSA, ldt, nv, torch.from_numpy(Y).cuda(), max_iter=15)      # AUTO: on chip where possible
        assert (out["iters"].cpu().numpy() == ref["iters"]).all() and (out["x_hat"].cpu().numpy() == ref["x_hat"]).all()
    y32 = torch.from_numpy(Y.astype(np.float32)).cuda()
    a = eng.decode_device_channel(lib.CH_BIAWGN, lib.SPA, lib.F32, nv, y32, max_iter=15, flags=lib.PATH_STREAMING)
    a = {k: (v.clone() if v is not None else None) for k, v in a.items()}
    b = eng.decode_device_channel(lib.CH_BIAWGN, lib.SPA, lib.F32, nv, y32, max_iter=15)
    assert bool((a["iters"] == b["iters"]).all()) and bool((a["x_hat"] == b["x_hat"]).all())


def test_long_irregular_code_streams_bit_exact(mods):
    """An irregular code too long for shared memory (n = 5000, variable degrees 1..8, a few short checks): the streaming
    sweeps with non-uniform degree tables, min-sum float32 / float64 and erasure decoding against the oracle."""
    torch, lib = mods["torch"], mods["lib"]
    from ldpc_decoders_b200 import codes
    prof = {2: 2760, 3: 1000, 4: 400, 7: 380, 8: 200, 1: 60, 5: 100, 6: 100}
    short = (-sum(d * c for d, c in prof.items())) % 6
    if short:
        prof[2] -= 1
        prof[2 + short] = prof.get(2 + short, 0) + 1
    tab = codes.random_irregular(prof, 6, seed=21).tables
    og = O.Graph(tab.m, tab.n, tab.edge_chk.astype(np.int64), tab.edge_var.astype(np.int64))
    eng = mods["engine"].engine_for(tab)
    assert eng.resident_kernel == "" and tab.n == 5000
    B = 200
    zeros = np.zeros((B, tab.n), np.int64)
    Y = G.channel_send("biawgn", 2.5, zeros, 8101)
    nv = 10 ** (-2.5 / 10)
    for dt, ldt in ((np.float32, lib.F32), (np.float64, lib.F64)):
        ref = O.bp_decode(og, O.MSA, O.llr_biawgn(2.5, Y).astype(dt), max_iter=12, nthreads=8)
        out = eng.decode_device_channel(lib.CH_BIAWGN, lib.MSA, ldt, nv, torch.from_numpy(Y).cuda(), max_iter=12)
        assert (out["iters"].cpu().numpy() == ref["iters"]).all() and (out["x_hat"].cpu().numpy() == ref["x_hat"]).all()
        assert (out["reason"].cpu().numpy() == ref["reason"]).all()
    Ye = G.channel_send("bec", .33, zeros, 8102).astype(np.uint8)
    ref = O.bec_decode(og, Ye, max_iter=60, nthreads=8)
    out = eng.decode_device_channel(lib.CH_BEC, lib.BEC, lib.F32, 0.0, torch.from_numpy(Ye).cuda(), max_iter=60)
    assert (out["iters"].cpu().numpy() == ref["iters"]).all() and (out["x_hat"].cpu().numpy() == ref["x_hat"]).all()
    assert (out["reason"].cpu().numpy() == ref["reason"]).all()


@pytest.mark.parametrize("seed", list(range(8)))
def test_random_irregular_codes_fuzz(mods, seed):
    """Randomly drawn irregular codes (codes.random_irregular: lengths that are not multiples of 4 or 8, variables of
    degree 0..8, checks that lost edges to cancellation) through every kernel family: on-chip float32 / float64
    min-sum and erasure decoding against the oracle, on-chip against streaming for sum-product."""
    torch, lib = mods["torch"], mods["lib"]
    from ldpc_decoders_b200 import codes
    rng = np.random.RandomState(1000 + seed)
    n = int(rng.randint(30, 900))
    prof = {}
    left = n
    for d, frac in ((2, .55), (3, .2), (4, .08), (7, .06), (8, .04), (1, .02), (0, .01), (5, .02), (6, .02)):
        c = min(left, int(round(frac * n * rng.uniform(.6, 1.4))))
        prof[d] = c
        left -= c
    prof[2] += left
    short = (-sum(d * c for d, c in prof.items())) % 6
    if short:
        prof[2] -= 1
        prof[2 + short] = prof.get(2 + short, 0) + 1
    tab = codes.random_irregular(prof, 6, seed=seed).tables
    if tab.check_degrees.min() < 2:
        pytest.skip("a check lost all but one of its edges: not a code the on-chip path takes")
    og = O.Graph(tab.m, tab.n, tab.edge_chk.astype(np.int64), tab.edge_var.astype(np.int64))
    eng = mods["engine"].engine_for(tab)
    assert eng.resident_kernel == "resident_vp"
    B = 300
    zeros = np.zeros((B, tab.n), np.int64)
    Y = G.channel_send("biawgn", 3.0, zeros, 4000 + seed)
    nv = 10 ** (-3.0 / 10)
    for dt, ldt in ((np.float32, lib.F32), (np.float64, lib.F64)):
        ref = O.bp_decode(og, O.MSA, O.llr_biawgn(3.0, Y).astype(dt), max_iter=20, nthreads=4)
        out = eng.decode_device_channel(lib.CH_BIAWGN, lib.MSA, ldt, nv, torch.from_numpy(Y).cuda(), max_iter=20, flags=lib.PATH_RESIDENT)
        assert (out["iters"].cpu().numpy() == ref["iters"]).all() and (out["x_hat"].cpu().numpy() == ref["x_hat"]).all()
        assert (out["reason"].cpu().numpy() == ref["reason"]).all()
    Yb = G.channel_send("bsc", .03, zeros, 5000 + seed).astype(np.uint8)
    Yb[3] = 0
    L = float(np.log(1 - .03) - np.log(.03))
    ref = O.bp_decode(og, O.MSA, O.llr_bsc(.03, Yb).astype(np.float32), y_hard=Yb, max_iter=20, nthreads=4)
    out = eng.decode_device_channel(lib.CH_BSC, lib.MSA, lib.F32, L, torch.from_numpy(Yb).cuda(), max_iter=20, flags=lib.PATH_RESIDENT)
    assert (out["iters"].cpu().numpy() == ref["iters"]).all() and (out["x_hat"].cpu().numpy() == ref["x_hat"]).all()
    for ch, prm, y in ((lib.CH_BIAWGN, nv, torch.from_numpy(Y.astype(np.float32)).cuda()), (lib.CH_BSC, L, torch.from_numpy(Yb).cuda())):
        a = eng.decode_device_channel(ch, lib.SPA, lib.F32, prm, y, max_iter=20, flags=lib.PATH_STREAMING)
        a = {k: (v.clone() if v is not None else None) for k, v in a.items()}
        b = eng.decode_device_channel(ch, lib.SPA, lib.F32, prm, y, max_iter=20, flags=lib.PATH_RESIDENT)
        assert bool((a["iters"] == b["iters"]).all()) and bool((a["x_hat"] == b["x_hat"]).all()) and bool((a["reason"] == b["reason"]).all())
    # tiny batches, one iteration, unlimited iterations with a cap, and the host entry point, on the same kernels
    yd = torch.from_numpy(Y).cuda()
    for ldt, dt in ((lib.F32, np.float32), (lib.F64, np.float64)):
        full = eng.decode_device_channel(lib.CH_BIAWGN, lib.MSA, ldt, nv, yd, max_iter=20, flags=lib.PATH_RESIDENT)
        full = {k: v.clone() for k, v in full.items() if v is not None}
        for b in (1, 5):
            part = eng.decode_device_channel(lib.CH_BIAWGN, lib.MSA, ldt, nv, yd[:b].contiguous(), max_iter=20, flags=lib.PATH_RESIDENT)
            assert bool((part["iters"] == full["iters"][:b]).all()) and bool((part["x_hat"] == full["x_hat"][:b]).all())
        for mi, cap in ((1, 0), (0, 30)):
            ref = O.bp_decode(og, O.MSA, O.llr_biawgn(3.0, Y[:64]).astype(dt), max_iter=mi, iter_cap=cap, nthreads=4)
            out = eng.decode_device_channel(lib.CH_BIAWGN, lib.MSA, ldt, nv, yd[:64].contiguous(), max_iter=mi, iter_cap=cap, flags=lib.PATH_RESIDENT)
            assert (out["iters"].cpu().numpy() == ref["iters"]).all() and (out["x_hat"].cpu().numpy() == ref["x_hat"]).all()
            assert (out["reason"].cpu().numpy() == ref["reason"]).all()
        xh, it, rs = eng.decode_host(lib.CH_BIAWGN, lib.MSA, ldt, nv, Y, max_iter=20, chunk=128)
        assert (it == full["iters"].cpu().numpy()).all() and (xh == full["x_hat"].cpu().numpy()).all()
    Ye = G.channel_send("bec", .35, zeros, 6000 + seed).astype(np.uint8)
    ref = O.bec_decode(og, Ye, max_iter=50, nthreads=4)
    out = eng.decode_device_channel(lib.CH_BEC, lib.BEC, lib.F32, 0.0, torch.from_numpy(Ye).cuda(), max_iter=50)
    assert (out["iters"].cpu().numpy() == ref["iters"]).all() and (out["x_hat"].cpu().numpy() == ref["x_hat"]).all()
    assert (out["reason"].cpu().numpy() == ref["reason"]).all()


# --------------------------------------------------------------------------------------------- SPA
@pytest.mark.parametrize("case", G.spa_tf(), ids=lambda c: c[0]["slot"])
def test_spa_teacher_forced(mods, case):
    """One check-node sweep from a shared input state (reference v2c snapshots), both SPA kernels."""
    torch, lib = mods["torch"], mods["lib"]
    rec, v2c, c2v_ref = case
    eng = mods["engine"].engine_for(tables(mods, rec["code"]))
    # float64 mirror against the reference's own c2v
    out, _ = eng.debug_step(lib.SPA, 0, torch.from_numpy(np.tile(v2c, (3, 1))).cuda())
    out = out.cpu().numpy()
    assert (out[0] == out[2]).all() or np.isnan(out[0]).any()
    fin = np.isfinite(c2v_ref) & np.isfinite(out[0])
    assert (np.isfinite(c2v_ref) == np.isfinite(out[0])).all() or (np.abs(c2v_ref[np.isfinite(c2v_ref) != np.isfinite(out[0])]) > 30).all()
    assert (np.abs(out[0][fin] - c2v_ref[fin]) <= spa_tol(c2v_ref[fin])).all()
    # float32 rule against the float64 formula evaluated on the same float32 inputs
    v32 = np.ascontiguousarray(v2c, np.float32)
    ref = O.cn_sweep(ograph(rec["code"]), O.SPA, v32.astype(np.float64))
    out32, _ = eng.debug_step(lib.SPA, 0, torch.from_numpy(v32[None, :]).cuda())
    out32 = out32.cpu().numpy()[0].astype(np.float64)
    a = np.abs(ref)
    m = a < 20
    assert (np.abs(out32[m] - ref[m]) <= 1e-4 * np.maximum(1, a[m])).all()
    big = ~m & np.isfinite(ref)
    assert (np.sign(out32[big]) == np.sign(ref[big])).all() and (np.abs(out32[big]) >= 15).all()


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_msa_and_vn_sweeps_bit_exact(mods, dt):
    torch, lib = mods["torch"], mods["lib"]
    for code in ("1200_3_6_rand_ldpc_1", "1200_rho_x5_rand_ldpc_10", "7_4_hamming"):
        og = ograph(code)
        eng = mods["engine"].engine_for(tables(mods, code))
        rng = np.random.RandomState(1)
        B = 5
        v2c = (rng.normal(size=(B, og.E)) * 3).astype(dt)
        v2c[0, :7] = [0.0, -0.0, 1.5, -1.5, 1.5, 2.0, -2.0]          # ties, zeros, negative zero
        prior = rng.normal(size=(B, og.n)).astype(dt)
        c2v, _ = eng.debug_step(lib.MSA, 0, torch.from_numpy(v2c).cuda())
        c2v = c2v.cpu().numpy()
        for b in range(B):
            ref = O.cn_sweep(og, O.MSA, v2c[b])
            assert (c2v[b] == ref).all() and (np.signbit(c2v[b]) == np.signbit(ref)).all()
        nxt, marg = eng.debug_step(lib.MSA, 1, torch.from_numpy(c2v).cuda(), prior=torch.from_numpy(prior).cuda())
        nxt, marg = nxt.cpu().numpy(), marg.cpu().numpy()
        for b in range(B):
            rv, rm, _ = O.vn_sweep(og, prior[b], c2v[b])
            assert (nxt[b] == rv).all() and (marg[b] == rm).all()


@pytest.mark.parametrize("channel,code,param,mi", [("biawgn", "1200_3_6_rand_ldpc_1", 2.0, 10),
                                                    ("biawgn", "1200_3_6_rand_ldpc_1", 2.75, 10),
                                                    ("bsc", "1200_rho_x5_rand_ldpc_1", .06, 40)])
def test_spa_f32_words_agree_away_from_boundaries(mods, channel, code, param, mi):
    """End to end: float32 SPA (hyperbolic-pair rule) vs the float64 reference formula on the same float32 priors."""
    frames = 1500
    tab, og = tables(mods, code), ograph(code)
    Y = G.channel_send(channel, param, np.zeros((frames, tab.n), np.int64), 31337)
    if channel == "bsc":
        Yh = Y.astype(np.uint8)
        pri32 = O.llr_bsc(param, Yh).astype(np.float32)
        ref = O.bp_decode(og, O.SPA, pri32.astype(np.float64), y_hard=Yh, max_iter=mi, want_marg=True, nthreads=8)
        x_hat, iters = mods["bsc"].SPA(param, tab, max_iter=mi, dtype=np.float32).decode_batch(Yh)
    else:
        pri32 = O.llr_biawgn(param, Y).astype(np.float32)
        ref = O.bp_decode(og, O.SPA, pri32.astype(np.float64), max_iter=mi, want_marg=True, nthreads=8)
        x_hat, iters = mods["biawgn"].SPA(param, tab, max_iter=mi, dtype=np.float32).decode_batch(Y)
    finite = np.isfinite(ref["marg"]).all(axis=1)
    clear = finite & (np.abs(ref["marg"]).min(axis=1) > 1e-3) & (ref["iters"] < mi)      # converged, no bit near 0
    assert clear.sum() > frames // 4
    # Loopy BP amplifies a 1e-7 perturbation by a constant factor per iteration, so frames that need many iterations
    # legitimately part ways between float32 and float64 (either arithmetic, any rounding: the float64 reference differs
    # from itself as much when its inputs are perturbed in the last bit).  Strict agreement is asked where the
    # trajectory is short, statistical agreement everywhere.
    early = clear & (ref["iters"] <= 10)
    assert early.sum() > frames // 4
    assert (x_hat[early] == ref["x_hat"][early]).all()
    assert (np.abs(iters[early] - ref["iters"][early]) <= 1).all()
    assert (iters[early] == ref["iters"][early]).mean() > 0.99
    assert (x_hat[clear] == ref["x_hat"][clear]).all(axis=1).mean() > 0.99
    assert (iters[clear] == ref["iters"][clear]).mean() > 0.97
    # error rates of the two arithmetic types agree statistically (same frames)
    wer32 = (x_hat != 0).any(axis=1).mean()
    wer64 = (ref["x_hat"] != 0).any(axis=1).mean()
    assert abs(wer32 - wer64) <= 3 * np.sqrt(max(wer64 * (1 - wer64), 1e-4) / frames) + 2e-3


# --------------------------------------------------------------------------------------------- edges of the batch API
@pytest.mark.parametrize("B", [1, 2, 31, 33, 127, 129, 257])
def test_ragged_batch_sizes(mods, B):
    tab, og = tables(mods, "512_3_6_rand_ldpc_1"), ograph("512_3_6_rand_ldpc_1")
    Y = G.channel_send("biawgn", 2.0, np.ones((B, tab.n), np.int64), 5)
    for dt in (np.float32, np.float64):
        ref = O.bp_decode(og, O.MSA, O.llr_biawgn(2.0, Y).astype(dt), max_iter=10)
        x_hat, iters = mods["biawgn"].MSA(2.0, tab, max_iter=10, dtype=dt).decode_batch(Y)
        assert (iters == ref["iters"]).all() and (x_hat == ref["x_hat"]).all()
    Yb = G.channel_send("bec", .4, np.ones((B, tab.n), np.int64), 6).astype(np.uint8)
    ref = O.bec_decode(og, Yb, max_iter=10)
    x_hat, iters, reason = mods["bec"].SPA(.4, tab, max_iter=10).decode_batch(Yb, return_reason=True)
    assert (iters == ref["iters"]).all() and (x_hat == ref["x_hat"]).all() and (reason == ref["reason"]).all()


def test_zero_iteration_exits(mods):
    """BSC: a received word that already satisfies H returns y itself after 0 iterations (bpa.py:20,29);
    BIAWGN: integer-valued y passes the reference's (H @ y) % 2 test too; BEC without erasures: 'decoded' at 0."""
    code = Code("1200_3_6_rand_ldpc_1")
    n = code.parity_mtx.shape[1]
    y = np.ones(n, int)
    dec = mods["bsc"].MSA(.05, code, max_iter=10)
    assert dec.decode(y) is y
    yf = np.ones(n)                       # noise-free BIAWGN, all-ones codeword: H @ y is even everywhere
    assert mods["biawgn"].MSA(2.0, code, max_iter=10).decode(yf) is yf
    yb = np.zeros(n, int)
    assert mods["bec"].SPA(.3, code, max_iter=10).decode(yb) is yb


def test_unlimited_iterations_are_capped_and_flagged(mods):
    tab, og = tables(mods, "1200_3_6_rand_ldpc_1"), ograph("1200_3_6_rand_ldpc_1")
    Y = G.channel_send("biawgn", 0.5, np.ones((40, tab.n), np.int64), 7)
    dec = mods["biawgn"].MSA(0.5, tab, max_iter=0, iter_cap=25, dtype=np.float32)
    x_hat, iters, reason = dec.decode_batch(Y, return_reason=True)
    ref = O.bp_decode(og, O.MSA, O.llr_biawgn(0.5, Y).astype(np.float32), max_iter=0, iter_cap=25)
    assert (iters == ref["iters"]).all() and (reason == ref["reason"]).all() and (x_hat == ref["x_hat"]).all()
    assert (reason == 4).any()


def test_long_run_with_polling_matches(mods):
    """max_iter = 100 takes the early-out polling path of ldpc_decode; results must not change."""
    tab, og = tables(mods, "1200_3_6_rand_ldpc_1"), ograph("1200_3_6_rand_ldpc_1")
    Y = G.channel_send("biawgn", 2.2, np.ones((300, tab.n), np.int64), 8)
    ref = O.bp_decode(og, O.MSA, O.llr_biawgn(2.2, Y), max_iter=100, nthreads=8)
    x_hat, iters = mods["biawgn"].MSA(2.2, tab, max_iter=100).decode_batch(Y)
    assert (iters == ref["iters"]).all() and (x_hat == ref["x_hat"]).all()


def test_device_and_host_entry_points_agree(mods):
    torch, lib = mods["torch"], mods["lib"]
    tab = tables(mods, "1200_3_6_rand_ldpc_1")
    eng = mods["engine"].engine_for(tab)
    Y = G.channel_send("biawgn", 2.0, np.ones((5000, tab.n), np.int64), 9).astype(np.float32)
    nv = 10 ** (-2.0 / 10)
    xh, it, rs = eng.decode_host(lib.CH_BIAWGN, lib.MSA, lib.F32, nv, Y, max_iter=10, chunk=1024)
    pri = eng.llr_biawgn(nv, torch.from_numpy(Y).cuda(), lib.F32)
    out = eng.decode_device(lib.MSA, pri, max_iter=10)
    assert (out["x_hat"].cpu().numpy() == xh).all() and (out["iters"].cpu().numpy() == it).all()
    assert (out["reason"].cpu().numpy() == rs).all()
    ref = (-2.0 * Y.astype(np.float64) / nv).astype(np.float32)
    assert (pri.cpu().numpy() == ref).all()


def test_host_entry_point_as_a_stream_of_batches(mods):
    """LDPC_HOST_ASYNC: several batches submitted back to back and completed by one ldpc_host_sync give the blocking
    call's results, each in its own output buffers (stream order keeps the handle's staging buffers consistent)."""
    lib, E = mods["lib"], mods["engine"]
    tab = tables(mods, "1200_3_6_rand_ldpc_1")
    eng = E.engine_for(tab)
    nv = 10 ** (-2.0 / 10)
    sets = []
    for k, B in enumerate((5000, 777, 9000)):
        Y = E.pinned_empty((B, tab.n), np.float32)
        Y[...] = G.channel_send("biawgn", 2.0, np.ones((B, tab.n), np.int64), 60 + k).astype(np.float32)
        sets.append((Y, E.pinned_empty((B, tab.n), np.uint8), E.pinned_empty((B,), np.int32), E.pinned_empty((B,), np.uint8)))
    ref = [tuple(a.copy() for a in eng.decode_host(lib.CH_BIAWGN, lib.MSA, lib.F32, nv, Y, max_iter=10, chunk=1024)) for Y, _, _, _ in sets]
    for Y, xh, it, rs in sets:
        xh[...] = 7
        eng.decode_host(lib.CH_BIAWGN, lib.MSA, lib.F32, nv, Y, max_iter=10, chunk=1024, x_hat=xh, iters=it, reason=rs, wait=False)
    eng.host_sync()
    for (Y, xh, it, rs), (rx, ri, rr) in zip(sets, ref):
        assert (xh == rx).all() and (it == ri).all() and (rs == rr).all()
    with pytest.raises(ValueError):
        eng.decode_host(lib.CH_BIAWGN, lib.MSA, lib.F32, nv, sets[0][0], max_iter=10, wait=False)


@pytest.mark.parametrize("code", ["1200_3_6_rand_ldpc_1", "1200_rho_x5_rand_ldpc_10", "7_4_hamming",
                                  "512_3_6_rand_ldpc_1", "margulis", "12_3_4_ldpc"])
def test_resident_and_streaming_paths_agree(mods, code):
    """The on-chip path (resident_bp, what LDPC_PATH_AUTO picks for short codes in float32) and the streaming
    sweeps run the same arithmetic: identical words, iteration counts and exit reasons, for every front end."""
    torch, lib = mods["torch"], mods["lib"]
    tab = tables(mods, code)
    eng = mods["engine"].engine_for(tab)
    if code == "margulis":                  # n = 2640: one variable-plane CTA per SM (resident_vp, 672 threads)
        assert eng.resident_kernel == "resident_vp"
    assert eng.resident_frames in (4, 8)
    for B in (1, 13, 700):
        Yg = G.channel_send("biawgn", 2.0, np.zeros((B, tab.n), np.int64), 77)
        Yb = G.channel_send("bsc", .05, np.zeros((B, tab.n), np.int64), 78).astype(np.uint8)
        for algo in (lib.MSA, lib.SPA):
            for mi in (10, 3):
                cases = [(lib.CH_BIAWGN, 10 ** (-2.0 / 10), torch.from_numpy(Yg.astype(np.float32)).cuda()),
                         (lib.CH_BIAWGN, 10 ** (-2.0 / 10), torch.from_numpy(Yg).cuda()),
                         (lib.CH_BSC, float(np.log(1 - .05) - np.log(.05)), torch.from_numpy(Yb).cuda())]
                for ch, prm, y in cases:
                    a = eng.decode_device_channel(ch, algo, lib.F32, prm, y, max_iter=mi, flags=lib.PATH_STREAMING)
                    a = {k: (v.clone() if v is not None else None) for k, v in a.items()}
                    n0 = eng.launch_count
                    b = eng.decode_device_channel(ch, algo, lib.F32, prm, y, max_iter=mi, flags=lib.PATH_RESIDENT)
                    assert eng.launch_count - n0 == 1                     # one kernel for the whole decode
                    assert bool((a["iters"] == b["iters"]).all()) and bool((a["reason"] == b["reason"]).all())
                    assert bool((a["x_hat"] == b["x_hat"]).all())
    # priors + hard input through ldpc_decode, and unlimited iterations with a cap
    pri = torch.from_numpy(O.llr_bsc(.05, Yb).astype(np.float32)).cuda()
    a = eng.decode_device(lib.MSA, pri, y_hard=torch.from_numpy(Yb).cuda(), max_iter=0, iter_cap=17, flags=lib.PATH_STREAMING)
    a = {k: (v.clone() if v is not None else None) for k, v in a.items()}
    b = eng.decode_device(lib.MSA, pri, y_hard=torch.from_numpy(Yb).cuda(), max_iter=0, iter_cap=17, flags=lib.PATH_RESIDENT)
    assert bool((a["iters"] == b["iters"]).all()) and bool((a["reason"] == b["reason"]).all()) and bool((a["x_hat"] == b["x_hat"]).all())


@pytest.mark.parametrize("code", ["1200_3_6_rand_ldpc_1", "1200_rho_x5_rand_ldpc_3", "12_3_4_ldpc"])
def test_spa_degenerate_messages_agree_between_paths(mods, code):
    """Saturation (|v2c| beyond the float64 tanh saturation point -> +-inf -> the NaN flood of bpa.py:37), exact
    zeros (0/0 = NaN on the edge's own output) and NaN priors take the rarely executed fix-up branches of cn_spa_sc;
    the on-chip kernel additionally reads hard decisions off sign bits.  Both paths must still agree bit for bit."""
    torch, lib = mods["torch"], mods["lib"]
    tab = tables(mods, code)
    eng = mods["engine"].engine_for(tab)
    B = 600
    Y = G.channel_send("biawgn", 1.0, np.zeros((B, tab.n), np.int64), 4242)
    pri = O.llr_biawgn(1.0, Y).astype(np.float32)
    pri[100:300] *= 12.0                                   # confidently wrong bits: saturates within a few iterations
    pri[300:400, ::7] = 0.0                                # exact zeros -> NaN messages
    pri[400:420, 3] = np.nan
    pri[420:440, 5] = -0.0
    d = torch.from_numpy(pri).cuda()
    for mi in (10, 40):
        a = eng.decode_device(lib.SPA, d, max_iter=mi, flags=lib.PATH_STREAMING)
        a = {k: (v.clone() if v is not None else None) for k, v in a.items()}
        b = eng.decode_device(lib.SPA, d, max_iter=mi, flags=lib.PATH_RESIDENT)
        assert bool((a["iters"] == b["iters"]).all()) and bool((a["reason"] == b["reason"]).all())
        assert bool((a["x_hat"] == b["x_hat"]).all())
    if tab.n >= 1200:
        xs = a["x_hat"].cpu().numpy()[100:300]             # a flooded frame (NaN marginals) decodes to the all-zero word
        assert (xs == 0).all(axis=1).mean() > 0.5 and (a["iters"].cpu().numpy()[100:300] < 10).mean() > 0.5


def test_resident_path_is_refused_where_it_cannot_run(mods):
    torch, lib = mods["torch"], mods["lib"]
    from ldpc_decoders_b200 import LdpcError, codes
    long_eng = mods["engine"].engine_for(codes.random_regular(6000, 3, 6, seed=1).tables)    # 4 frames need > 227 KB
    assert long_eng.resident_frames == 0 and long_eng.resident_kernel == ""
    with pytest.raises(LdpcError):
        long_eng.decode_device(lib.MSA, torch.zeros((4, 6000), dtype=torch.float32, device="cuda"), max_iter=3, flags=lib.PATH_RESIDENT)
    eng = mods["engine"].engine_for(tables(mods, "1200_3_6_rand_ldpc_1"))
    pri = torch.zeros((4, 1200), dtype=torch.float64, device="cuda")
    with pytest.raises(LdpcError):
        eng.decode_device(lib.SPA, pri, max_iter=3, flags=lib.PATH_RESIDENT)      # float64 sum-product has no resident kernel
    mar = mods["engine"].engine_for(tables(mods, "margulis"))
    with pytest.raises(LdpcError):                                               # float64 min-sum: not in the one-CTA-per-SM geometry
        mar.decode_device(lib.MSA, torch.zeros((4, 2640), dtype=torch.float64, device="cuda"), max_iter=3, flags=lib.PATH_RESIDENT)


@pytest.mark.parametrize("code", ["1200_3_6_rand_ldpc_2", "512_3_6_rand_ldpc_3", "1200_3_6_ldpc", "1200_rho_x5_rand_ldpc_5",
                                  "1200_rho_x5_rand_ldpc_7", "7_4_hamming", "12_3_4_ldpc"])
def test_float64_on_chip_min_sum(mods, code):
    """resident_vd: float64 min-sum on chip (the reference's own arithmetic).  Same words, iteration counts and exit
    reasons as the float64 streaming sweeps for every front end, and as the float64 oracle."""
    torch, lib = mods["torch"], mods["lib"]
    tab, og = tables(mods, code), ograph(code)
    eng = mods["engine"].engine_for(tab)
    B = 777
    cw = 1 if "3_6" in code else 0                                               # all-ones is a codeword of the (3,6) codes only
    Yg = G.channel_send("biawgn", 2.0, np.zeros((B, tab.n), np.int64) + cw, 515)
    Yb = G.channel_send("bsc", .05, np.zeros((B, tab.n), np.int64) + cw, 516).astype(np.uint8)
    Yb[5] = cw                                                                   # a clean word: iteration-0 exit
    nv = 10 ** (-2.0 / 10)
    for mi in (10, 3, 40):
        cases = [(lib.CH_BIAWGN, nv, torch.from_numpy(Yg).cuda()),
                 (lib.CH_BIAWGN, nv, torch.from_numpy(Yg.astype(np.float32)).cuda()),
                 (lib.CH_BSC, float(np.log(1 - .05) - np.log(.05)), torch.from_numpy(Yb).cuda())]
        for ch, prm, y in cases:
            a = eng.decode_device_channel(ch, lib.MSA, lib.F64, prm, y, max_iter=mi, flags=lib.PATH_STREAMING)
            a = {k: (v.clone() if v is not None else None) for k, v in a.items()}
            n0 = eng.launch_count
            b = eng.decode_device_channel(ch, lib.MSA, lib.F64, prm, y, max_iter=mi, flags=lib.PATH_RESIDENT)
            assert eng.launch_count - n0 == 1
            assert bool((a["iters"] == b["iters"]).all()) and bool((a["reason"] == b["reason"]).all())
            assert bool((a["x_hat"] == b["x_hat"]).all())
        ref = O.bp_decode(og, O.MSA, O.llr_biawgn(2.0, Yg), max_iter=mi, nthreads=8)
        c = eng.decode_device_channel(lib.CH_BIAWGN, lib.MSA, lib.F64, nv, torch.from_numpy(Yg).cuda(), max_iter=mi)   # AUTO
        assert (c["iters"].cpu().numpy() == ref["iters"]).all() and (c["x_hat"].cpu().numpy() == ref["x_hat"]).all()
        assert (c["reason"].cpu().numpy() == ref["reason"]).all()
        refb = O.bp_decode(og, O.MSA, O.llr_bsc(.05, Yb), y_hard=Yb, max_iter=mi, nthreads=8)
        pri = torch.from_numpy(O.llr_bsc(.05, Yb)).cuda()
        d = eng.decode_device(lib.MSA, pri, y_hard=torch.from_numpy(Yb).cuda(), max_iter=mi, flags=lib.PATH_RESIDENT)
        assert (d["iters"].cpu().numpy() == refb["iters"]).all() and (d["x_hat"].cpu().numpy() == refb["x_hat"]).all()
        assert d["iters"][5].item() == 0


def test_register_and_bulk_async_check_node_sweeps_agree(mods):
    """Both stagings of the check-node sweep (cn_sweep_tma, default; cn_sweep, LDPC_CN_REGISTER) run the same
    arithmetic: identical words, iteration counts and marginals, all decoders and dtypes."""
    torch, lib = mods["torch"], mods["lib"]
    for code in ("1200_3_6_rand_ldpc_1", "1200_rho_x5_rand_ldpc_10", "7_4_hamming"):
        tab = tables(mods, code)
        eng = mods["engine"].engine_for(tab)
        Y = G.channel_send("biawgn", 2.0, np.zeros((700, tab.n), np.int64), 21)
        for algo in (lib.MSA, lib.SPA):
            for dt in (np.float32, np.float64):
                pri = torch.from_numpy(O.llr_biawgn(2.0, Y).astype(dt)).cuda()
                a = eng.decode_device(algo, pri, max_iter=10, want_marg=True, flags=lib.PATH_STREAMING)
                a = {k: v.clone() for k, v in a.items()}
                b = eng.decode_device(algo, pri, max_iter=10, want_marg=True, flags=lib.PATH_STREAMING | lib.CN_REGISTER)
                assert bool((a["iters"] == b["iters"]).all()) and bool((a["x_hat"] == b["x_hat"]).all())
                assert bool(((a["marg"] == b["marg"]) | (a["marg"].isnan() & b["marg"].isnan())).all())


# --------------------------------------------------------------------------------------------- size-independent properties
def test_sign_symmetry_at_full_batch(mods):
    """All-ones is a codeword of a (3,6) code, so flipping the sign of every prior must flip every decoded
    bit and leave every iteration count unchanged (min-sum is odd in its inputs).  Checked at a batch
    larger than anything the oracle is run on."""
    torch, lib = mods["torch"], mods["lib"]
    tab = tables(mods, "1200_3_6_rand_ldpc_1")
    eng = mods["engine"].engine_for(tab)
    B = 32768
    g = torch.Generator(device="cuda").manual_seed(1)
    nv = 10 ** (-2.0 / 10)
    y = 1.0 + nv ** .5 * torch.randn((B, tab.n), generator=g, device="cuda", dtype=torch.float32)
    pri = eng.llr_biawgn(nv, y, lib.F32)
    a = eng.decode_device(lib.MSA, pri, max_iter=10)
    xa, ia = a["x_hat"].clone(), a["iters"].clone()
    b = eng.decode_device(lib.MSA, -pri, max_iter=10)
    assert bool((ia == b["iters"]).all())
    assert bool(((xa ^ 1) == b["x_hat"]).all())
    # decoded frames satisfy every check; frames at max_iter are exactly the undecoded ones
    H = torch.from_numpy(tab.dense(np.float32)).cuda()
    syn = (xa.float() @ H.T) % 2
    ok = (syn == 0).all(dim=1)
    assert bool(ok[a["reason"] == 0].all())
    assert bool((ia[a["reason"] == 1] == 10).all())
    assert 0.1 < float(ok.float().mean()) < 0.6            # WER ~0.7 at 2 dB


def test_bec_round_trip_property(mods):
    """Erase, decode: every symbol the decoder resolves equals the transmitted bit, unresolved ones stay 2,
    and a frame reported 'decoded' has no erasures left."""
    tab = tables(mods, "1200_3_6_rand_ldpc_1")
    B = 50000
    rng = np.random.RandomState(12)
    Y = np.where(rng.random_sample((B, tab.n)) < .36, 2, 0).astype(np.uint8)
    x_hat, iters, reason = mods["bec"].SPA(.36, tab, max_iter=100).decode_batch(Y, return_reason=True)
    assert ((x_hat == 0) | (x_hat == 2)).all()
    assert ((x_hat == 2) <= (Y == 2)).all()
    assert ((x_hat == 2).sum(axis=1)[reason == 0] == 0).all()
    assert ((x_hat == 2).sum(axis=1)[reason == 2] > 0).all()
    assert (reason == 0).mean() > 0.8


# --------------------------------------------------------------------------------------------- config 5: n = 64800
_long = {}


def long_code(mods):
    """The synthetic (3,6) code of BASELINE.json config 5 (seeded configuration model, O(E))."""
    if "code" not in _long:
        from ldpc_decoders_b200 import codes
        _long["code"] = codes.random_regular(64800, 3, 6, seed=0)
    return _long["code"]


@pytest.mark.parametrize("snr,dt", [(1.0, np.float32), (2.5, np.float32), (2.5, np.float64)])
def test_long_code_msa_bit_exact(mods, snr, dt):
    """n = 64800 (streaming path, 194400 edges): words, iteration counts and exit reasons of min-sum are
    bit-identical with the oracle at the same dtype; 1.0 dB never converges (fixed work), 2.5 dB exits early."""
    tab = long_code(mods).tables
    assert (tab.n, tab.m, tab.E) == (64800, 32400, 194400)
    og = O.Graph(tab.m, tab.n, tab.edge_chk, tab.edge_var)
    B = 24
    rng = np.random.RandomState(64800)
    Y = 1.0 + np.sqrt(O.noise_var(snr)) * rng.standard_normal((B, tab.n))
    dec = mods["biawgn"].MSA(snr, tab, max_iter=10, dtype=dt)
    x_hat, iters = dec.decode_batch(Y)
    ref = O.bp_decode(og, O.MSA, O.llr_biawgn(snr, Y).astype(dt), max_iter=10, nthreads=8)
    assert (iters == ref["iters"]).all()
    assert (x_hat == ref["x_hat"]).all()
    if snr == 1.0:
        assert (iters == 10).all()
    else:
        assert iters.min() < 10 and (x_hat == 1).all()


def test_long_code_full_batch_properties(mods):
    """BASELINE-size batch of the long code (2048 frames, 1.6 GB of messages): sign symmetry (all-ones is a
    codeword of a code with even check degree), every frame reported decoded satisfies every check, and the
    result does not depend on where in the batch a frame sits."""
    torch, lib = mods["torch"], mods["lib"]
    tab = long_code(mods).tables
    eng = mods["engine"].engine_for(tab)
    B = 2048
    g = torch.Generator(device="cuda").manual_seed(5)
    nv = 10 ** (-2.5 / 10)
    y = 1.0 + nv ** .5 * torch.randn((B, tab.n), generator=g, device="cuda", dtype=torch.float32)
    pri = eng.llr_biawgn(nv, y, lib.F32)
    a = eng.decode_device(lib.MSA, pri, max_iter=10)
    xa, ia, ra = a["x_hat"].clone(), a["iters"].clone(), a["reason"].clone()
    b = eng.decode_device(lib.MSA, -pri, max_iter=10)
    assert bool((ia == b["iters"]).all()) and bool(((xa ^ 1) == b["x_hat"]).all())
    # syndrome of every word through the edge list (no dense H at this length)
    rows = torch.from_numpy(np.asarray(tab.edge_chk, np.int64)).cuda()
    cols = torch.from_numpy(np.asarray(tab.edge_var, np.int64)).cuda()
    syn = torch.zeros((B, tab.m), dtype=torch.int32, device="cuda")
    syn.index_add_(1, rows, xa[:, cols].to(torch.int32))
    ok = ((syn & 1) == 0).all(dim=1)
    assert bool(ok[ra == 0].all())
    assert bool((ia[ra == 1] == 10).all()) and bool((ia[ra == 0] < 10).any())
    # a permuted batch decodes to the permuted result
    perm = torch.randperm(B, generator=torch.Generator(device="cuda").manual_seed(6), device="cuda")
    c = eng.decode_device(lib.MSA, pri[perm].contiguous(), max_iter=10)
    assert bool((c["iters"] == ia[perm]).all()) and bool((c["x_hat"] == xa[perm]).all())
