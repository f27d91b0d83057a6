"""GPU tier: the batched driver (sim.main) and the reference's own per-frame loop over the drop-in models."""
import json
import os

import numpy as np
import pytest

import _golden as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def write_code_txt(directory, name):
    m, n, rows, cols = G.code_tables(name)
    with open(os.path.join(directory, name + ".txt"), "w") as fp:
        for c in range(m):
            fp.write(" ".join(str(v + 1) for v in cols[rows == c]) + "\n")     # 1-based, src/codes.py:131-136


@pytest.mark.parametrize("channel,decoder,param,cw,dtype", [("biawgn", "MSA", 1.75, 1, "f32"),
                                                             ("bsc", "MSA", .06, 1, "f64"),
                                                             ("bec", "SPA", .42, 0, "f64")])
def test_sim_main_counters_match_sequential_reference_loop(tmp_path, channel, decoder, param, cw, dtype):
    from ldpc_decoders_b200 import dist, sim
    name = "512_3_6_rand_ldpc_1"
    write_code_txt(str(tmp_path), name)
    sim.main([channel, name, decoder, "--codeword", str(cw), "--min-wec", "6", "--params", str(param),
              "--max-iter", "10", "--batch", "32", "--dtype", dtype, "--seed", "5", "--console",
              "--data_dir", str(tmp_path), "--codes-dir", str(tmp_path)])
    out = json.load(open(os.path.join(str(tmp_path), "%s-%s-%s-%d-6-10.json" % (channel, name, decoder, cw))))
    g = O.Graph(*G.code_tables(name))
    x = np.zeros(g.n, np.int64) + cw
    dt = np.float32 if dtype == "f32" else np.float64

    def decode_batch(Y):                                   # the oracle, frame by frame semantics
        if channel == "bec":
            r = O.bec_decode(g, Y.astype(np.uint8), max_iter=10)
        elif channel == "bsc":
            r = O.bp_decode(g, O.MSA, O.llr_bsc(param, Y.astype(np.uint8)).astype(dt), y_hard=Y.astype(np.uint8), max_iter=10)
        else:
            r = O.bp_decode(g, O.MSA, O.llr_biawgn(param, Y).astype(dt), max_iter=10)
        return r["x_hat"], r["iters"]

    np.random.seed(5)
    ref = sim.run_param(decode_batch, lambda X: _send(channel, param, X), x, dist.Comm(), 1, 6)
    key = str(float(param))
    assert (out["tot"][key], out["wec"][key], out["bec"][key]) == (ref["tot"], ref["wec"], ref["bec"])
    assert out["dec"][key]["iter"] == ref["dec"]["iter"]
    assert list(out)[:6] == ["channel", "code", "decoder", "codeword", "min_wec", "max_iter"]


def _send(channel, param, x):
    if channel == "bec":
        return np.clip(x + (np.random.random(x.shape) < param).astype(int) * 10, 0, 2)
    if channel == "bsc":
        return (x + (np.random.random(x.shape) < param)) % 2
    return (2 * x - 1) + np.random.normal(0, np.sqrt(10 ** (-param / 10)), x.shape)


def test_reference_frame_loop_runs_on_dropin_models():
    """main.test's loop (src/main.py:22-45) with `models` swapped for ours: per-frame decode(y), same counters."""
    from ldpc_decoders_b200.models import models

    class Code:                                            # codes.Code stand-in: dense parity_mtx + get_n
        parity_mtx = G.dense_H("512_3_6_rand_ldpc_1")
        def get_n(self): return self.parity_mtx.shape[1]

    args = dict(channel="bsc", code="512_3_6_rand_ldpc_1", decoder="MSA", codeword=1, min_wec=4, params=[.07],
                max_iter=10, mu=3., eps=1e-5, allow_pseudo=False, layers=[100, 100], train=False, apprx=-1, log_freq=5.)
    model = models[args["channel"]]
    dec_fac = getattr(model, args["decoder"])
    assert ['channel', 'code', 'decoder', 'codeword', 'min_wec'] + dec_fac.id_keys == \
        ['channel', 'code', 'decoder', 'codeword', 'min_wec', 'max_iter']
    code = Code()
    x = code.parity_mtx[0] * 0 + args["codeword"]
    g = O.Graph.from_dense(code.parity_mtx)
    for param in args["params"]:
        channel = model.Channel(param)
        decoder = dec_fac(param, code, **args)
        np.random.seed(9)
        tot = wec = bec = 0
        ys = []
        while wec < args["min_wec"]:
            y = channel.send(x)
            x_hat = decoder.decode(y)
            errors = (~(x == x_hat)).sum()
            wec += errors > 0; bec += errors; tot += 1
            ys.append(y)
        Y = np.array(ys).astype(np.uint8)
        ref = O.bp_decode(g, O.MSA, O.llr_bsc(param, Y), y_hard=Y, max_iter=10)
        e = (ref["x_hat"] != x[None, :]).sum(1)
        assert bec == e.sum() and wec == (e > 0).sum() and (e[-1] > 0)
        assert hasattr(decoder, "stats") and sum(decoder.stats()["iter"]) == tot


# --------------------------------------------------------------------------------------------- on-device channel
def _engine(name="1200_3_6_rand_ldpc_1"):
    import torch
    from ldpc_decoders_b200 import Tables, _lib, engine
    torch.cuda.set_device(0)
    tab = Tables(*G.code_tables(name))
    return torch, _lib, tab, engine.engine_for(tab)


def test_device_channels_have_the_reference_statistics():
    """ldpc_channel_generate against the definitions of bec.py:15-18, bsc.py:15-16, biawgn.py:17-18."""
    torch, lib, tab, eng = _engine()
    B = 4096
    x = torch.from_numpy((np.arange(tab.n) % 3 == 0).astype(np.uint8)).cuda()
    N = B * tab.n
    nv = 10 ** (-2.0 / 10)
    y = eng.channel_generate(lib.CH_BIAWGN, nv, B, seed=11, x=x).cpu().numpy().astype(np.float64)
    z = (y - (2.0 * x.cpu().numpy()[None, :] - 1)) / np.sqrt(nv)
    assert abs(z.mean()) < 4 / np.sqrt(N) and abs(z.var() - 1) < 4 * np.sqrt(2 / N)
    assert abs((np.abs(z) > 2).mean() - 0.0455) < 4 * np.sqrt(0.0455 / N)
    for ch, p in ((lib.CH_BSC, .06), (lib.CH_BEC, .4)):
        y = eng.channel_generate(ch, p, B, seed=12, x=x).cpu().numpy()
        hit = (y != x.cpu().numpy()[None, :])
        assert abs(hit.mean() - p) < 4 * np.sqrt(p * (1 - p) / N)
        assert set(np.unique(y).tolist()) <= ({0, 1} if ch == lib.CH_BSC else {0, 1, 2})
        if ch == lib.CH_BEC:
            assert (y[hit] == 2).all()
        assert abs(hit.mean(axis=0).std() - np.sqrt(p * (1 - p) / B)) < 0.3 * np.sqrt(p * (1 - p) / B)   # no per-variable bias
    # all-zero word when x is NULL, and bit errors counted like main.py:41
    y0 = eng.channel_generate(lib.CH_BSC, .1, 64, seed=3)
    errs = eng.count_errors(y0).cpu().numpy()
    assert (errs == y0.cpu().numpy().sum(axis=1)).all()
    errs1 = eng.count_errors(y0, x).cpu().numpy()
    assert (errs1 == (y0.cpu().numpy() != x.cpu().numpy()[None, :]).sum(axis=1)).all()


@pytest.mark.parametrize("channel,algo,param", [("biawgn", "MSA", 2.0), ("bsc", "SPA", .05), ("bec", "SPA", .4)])
def test_device_noise_is_keyed_by_global_frame_index(channel, algo, param):
    """Same (seed, frame index) -> same frame, however the run is cut into batches (and so over GPUs), and whichever
    path decodes it."""
    torch, lib, tab, eng = _engine()
    ch = dict(biawgn=lib.CH_BIAWGN, bsc=lib.CH_BSC, bec=lib.CH_BEC)[channel]
    al = lib.MSA if algo == "MSA" else lib.SPA
    prm = 10 ** (-param / 10) if channel == "biawgn" else param
    whole = eng.simulate(ch, al, lib.F32, prm, 1000, seed=77, frame0=5000)
    e, i = whole["bit_errs"].cpu().numpy(), whole["iters"].cpu().numpy()
    parts_e, parts_i = [], []
    for f0, b in ((5000, 300), (5300, 1), (5301, 699)):
        r = eng.simulate(ch, al, lib.F32, prm, b, seed=77, frame0=f0)
        parts_e.append(r["bit_errs"].cpu().numpy()); parts_i.append(r["iters"].cpu().numpy())
    assert (np.concatenate(parts_e) == e).all() and (np.concatenate(parts_i) == i).all()
    if channel != "bec":
        s = eng.simulate(ch, al, lib.F32, prm, 1000, seed=77, frame0=5000, flags=lib.PATH_STREAMING)
        assert (s["bit_errs"].cpu().numpy() == e).all() and (s["iters"].cpu().numpy() == i).all()
    other = eng.simulate(ch, al, lib.F32, prm, 1000, seed=78, frame0=5000)
    assert (other["bit_errs"].cpu().numpy() != e).any()
    assert 0 < (e > 0).mean() < 1


def test_device_noise_error_rates_agree_with_numpy_noise():
    """WER / BER of a device-noise run lie within the Monte-Carlo confidence interval of a numpy-noise run
    (BIAWGN 2.2 dB min-sum, the reference's operating region)."""
    from ldpc_decoders_b200 import biawgn
    torch, lib, tab, eng = _engine()
    frames, snr = 40000, 2.2
    x = np.ones(tab.n, np.int64)
    dec = biawgn.MSA(snr, tab, max_iter=10, dtype=np.float32)
    errs_d, iters_d = dec.simulate_batch(x, frames, seed=1)
    Y = G.channel_send("biawgn", snr, np.tile(x, (frames, 1)), 99).astype(np.float32)
    xh, iters_h = dec.decode_batch(Y)
    errs_h = (xh != 1).sum(axis=1)
    wd, wh = (errs_d > 0).mean(), (errs_h > 0).mean()
    assert abs(wd - wh) < 4 * np.sqrt(2 * wh * (1 - wh) / frames)
    bd, bh = errs_d.mean(), errs_h.mean()
    assert abs(bd - bh) < 4 * np.sqrt((errs_d.var() + errs_h.var()) / frames)
    assert abs(iters_d.mean() - iters_h.mean()) < 4 * np.sqrt(2 * iters_h.var() / frames)


def test_sim_main_with_device_noise_is_independent_of_batch(tmp_path):
    from ldpc_decoders_b200 import sim
    name = "512_3_6_rand_ldpc_1"
    write_code_txt(str(tmp_path), name)
    res = []
    for batch, sub in ((64, "a"), (1000, "b")):
        d = os.path.join(str(tmp_path), sub)
        os.makedirs(d)
        sim.main(["biawgn", name, "MSA", "--codeword", "1", "--min-wec", "40", "--params", "2.0", "--max-iter", "10",
                  "--batch", str(batch), "--dtype", "f32", "--seed", "5", "--noise", "device", "--console",
                  "--data_dir", d, "--codes-dir", str(tmp_path)])
        res.append(json.load(open(os.path.join(d, "biawgn-%s-MSA-1-40-10.json" % name))))
    a, b = res
    assert a["tot"] == b["tot"] and a["wec"] == b["wec"] and a["bec"] == b["bec"] and a["dec"] == b["dec"]
    assert a["wec"]["2.0"] == 40
