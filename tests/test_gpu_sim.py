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
