"""Randomised soak of the on-chip kernels against the oracle at batch sizes that force frame hand-overs.

The fuzz tests of test_gpu_parity.py use batches that fit the kernels' slots (no frame ever follows another one into a
slot); the 1e4-frame tests exercise the hand-over on two codes at fixed parameters.  Here every case draws a code, a
channel, an operating point, an iteration bound, an arithmetic type and a batch size of a few thousand frames
(several frames per slot, ragged last tile), and the words, iteration counts and exit reasons must equal the oracle's.
LDPC_SOAK_CASES (default 6) sets the number of cases, LDPC_SOAK_SEED which ones; profiles/r2/soak.log holds runs
with 1000 and 2500 cases.
"""
import os

import numpy as np
import pytest

import _golden as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu

CASES = int(os.environ.get("LDPC_SOAK_CASES", "6"))
SEED = int(os.environ.get("LDPC_SOAK_SEED", "31000"))           # a different value draws a different set of cases
CODES = ["1200_3_6_rand_ldpc_1", "1200_3_6_rand_ldpc_7", "1200_rho_x5_rand_ldpc_2", "1200_rho_x5_rand_ldpc_9",
         "512_3_6_rand_ldpc_1", "margulis", "12_3_4_ldpc", "7_4_hamming"]


@pytest.fixture(scope="module")
def mods():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ldpc_decoders_b200 as pkg
    from ldpc_decoders_b200 import _lib, bec, biawgn, bsc, engine
    return dict(torch=torch, pkg=pkg, lib=_lib, bec=bec, biawgn=biawgn, bsc=bsc, engine=engine)


@pytest.mark.parametrize("case", list(range(CASES)))
def test_soak_case(mods, case):
    rng = np.random.RandomState(SEED + case)
    code = CODES[rng.randint(len(CODES))]
    m, n, rows, cols = G.code_tables(code)
    tab = mods["pkg"].Tables(m, n, rows, cols)
    og = O.Graph(m, n, rows, cols)
    small = n < 100
    # several frames per slot: 296 CTAs x 4 frames (148 x 4 for the one-CTA geometry, 64-frame tiles for erasures)
    B = int(rng.randint(2500, 7000)) if not small else int(rng.randint(40000, 90000))
    mi = int(rng.choice([1, 2, 5, 10, 10, 10, 25, 100]))
    cw = int(rng.randint(2)) if code.startswith("1200_3_6") else 0        # all-ones is a codeword of the (3,6) codes (even checks)
    x = np.zeros((B, n), np.int64) + cw
    kind = rng.choice(["biawgn", "bsc", "bec"])
    dt = np.float32 if rng.rand() < .6 else np.float64
    if kind == "bec":
        p = float(rng.uniform(.3, .5)) if not small else float(rng.uniform(.1, .4))
        Y = G.channel_send("bec", p, x, SEED % 1000 + 100 + case).astype(np.uint8)
        ref = O.bec_decode(og, Y, max_iter=mi, nthreads=8)
        x_hat, iters, reason = mods["bec"].SPA(p, tab, max_iter=mi).decode_batch(Y, return_reason=True)
    elif kind == "bsc":
        p = float(rng.uniform(.02, .09))
        Yh = G.channel_send("bsc", p, x, SEED % 1000 + 200 + case).astype(np.uint8)
        ref = O.bp_decode(og, O.MSA, O.llr_bsc(p, Yh).astype(dt), y_hard=Yh, max_iter=mi, nthreads=8)
        x_hat, iters, reason = mods["bsc"].MSA(p, tab, max_iter=mi, dtype=dt).decode_batch(Yh, return_reason=True)
    else:
        snr = float(rng.uniform(.5, 3.5))
        Y = G.channel_send("biawgn", snr, x, SEED % 1000 + 300 + case)
        ref = O.bp_decode(og, O.MSA, O.llr_biawgn(snr, Y).astype(dt), max_iter=mi, nthreads=8)
        x_hat, iters, reason = mods["biawgn"].MSA(snr, tab, max_iter=mi, dtype=dt).decode_batch(Y, return_reason=True)
    what = "%s %s %s B=%d max_iter=%d cw=%d" % (code, kind, dt.__name__, B, mi, cw)
    assert (iters == ref["iters"]).all(), what
    assert (reason == ref["reason"]).all(), what
    assert (x_hat == ref["x_hat"]).all(), what
    print("soak ok:", what, "mean iters %.2f" % iters.mean())


@pytest.mark.parametrize("case", list(range(max(2, CASES // 3))))
def test_soak_sum_product_on_chip_equals_streaming(mods, case):
    """Sum-product float32 has no bit-exact CPU counterpart (tolerance tests hold it to the reference); the on-chip kernel
    and the streaming sweeps run the same arithmetic, so at hand-over batch sizes they must agree bit for bit."""
    torch, lib = mods["torch"], mods["lib"]
    rng = np.random.RandomState(SEED + 16000 + case)
    code = CODES[rng.randint(len(CODES) - 2)]                             # the 1200 / 512 / Margulis codes
    m, n, rows, cols = G.code_tables(code)
    eng = mods["engine"].engine_for(mods["pkg"].Tables(m, n, rows, cols))
    B = int(rng.randint(2500, 6000))
    mi = int(rng.choice([2, 5, 10, 10, 25, 100]))
    x = np.zeros((B, n), np.int64)
    if rng.rand() < .5:
        snr = float(rng.uniform(.5, 3.5))
        ch, prm = lib.CH_BIAWGN, 10 ** (-snr / 10)
        y = torch.from_numpy(G.channel_send("biawgn", snr, x, 400 + case).astype(np.float32)).cuda()
    else:
        p = float(rng.uniform(.02, .09))
        ch, prm = lib.CH_BSC, float(np.log(1 - p) - np.log(p))
        y = torch.from_numpy(G.channel_send("bsc", p, x, 500 + case).astype(np.uint8)).cuda()
    a = eng.decode_device_channel(ch, lib.SPA, lib.F32, prm, y, max_iter=mi, flags=lib.PATH_STREAMING)
    a = {k: (v.clone() if v is not None else None) for k, v in a.items()}
    n0 = eng.launch_count
    b = eng.decode_device_channel(ch, lib.SPA, lib.F32, prm, y, max_iter=mi)
    assert eng.launch_count - n0 == 1                                     # the on-chip kernel: one launch
    what = "%s SPA f32 B=%d max_iter=%d" % (code, B, mi)
    for k in ("iters", "reason", "x_hat"):
        assert bool((a[k] == b[k]).all()), (what, k)
    print("soak ok:", what, "mean iters %.2f" % float(b["iters"].float().mean()))
