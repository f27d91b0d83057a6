"""Helpers shared by the tests: load the committed golden fixtures and re-create their inputs.

The fixtures were produced by tests/golden/make_golden.py from the unmodified
reference; inputs are re-drawn here from the stored seed with numpy's legacy
global RNG exactly as the reference's Channel.send does (src/bec.py:15-18,
src/bsc.py:15-16, src/biawgn.py:17-18) and verified against the stored sha256.
"""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

_codes = None


def code_tables(name):
    """(m, n, rows, cols) of a golden code in np.where(H) order."""
    global _codes
    if _codes is None:
        _codes = np.load(os.path.join(GOLD, "codes.npz"))
    m, n = (int(v) for v in _codes[name + "__shape"])
    return m, n, _codes[name + "__rows"].astype(np.int64), _codes[name + "__cols"].astype(np.int64)


def dense_H(name):
    m, n, r, c = code_tables(name)
    H = np.zeros((m, n), np.int64)
    H[r, c] = 1
    return H


def kats():
    with open(os.path.join(GOLD, "kat.json")) as fp:
        return json.load(fp)


_runs = None


def runs():
    """List of run records; arrays via run_arrays(rec)."""
    global _runs
    if _runs is None:
        z = np.load(os.path.join(GOLD, "runs.npz"))
        _runs = (z, json.loads(bytes(z["index_json"]).decode()))
    return _runs[1]


def run_arrays(rec):
    z = runs() and _runs[0]
    k = rec["slot"]
    xh = z[k + "_xhat"]
    if rec["channel"] != "bec":
        xh = np.unpackbits(xh, axis=1)[:, :rec["n"]]
    out = dict(x_hat=xh.astype(np.uint8), iters=z[k + "_iters"].astype(np.int32), reason=z[k + "_reason"])
    if k + "_marg4" in z.files:
        out["marg4"] = z[k + "_marg4"]
        out["nonfinite"] = z[k + "_nonfinite"].astype(bool)
    return out


def spa_tf():
    z = np.load(os.path.join(GOLD, "spa_tf.npz"))
    idx = json.loads(bytes(z["index_json"]).decode())
    return [(r, z[r["slot"] + "_v2c"], z[r["slot"] + "_c2v"]) for r in idx]


def channel_send(channel, param, x, seed):
    """Re-create the received block of a golden run (same RNG call sequence as the reference)."""
    np.random.seed(seed)
    if channel == "bec":        # src/bec.py:15-18
        tt = (np.random.random(x.shape) < param).astype(int)
        return np.clip(x + tt * 10, 0, 2)
    if channel == "bsc":        # src/bsc.py:15-16
        return (x + (np.random.random(x.shape) < param)) % 2
    if channel == "biawgn":     # src/biawgn.py:10,14,17-18
        std = np.sqrt(10 ** (-param / 10))
        return (2 * x - 1) + np.random.normal(0, std, x.shape)
    raise KeyError(channel)


def run_inputs(rec):
    x = np.zeros(rec["n"], np.int64) + rec["codeword"]
    Y = channel_send(rec["channel"], rec["param"], np.tile(x, (rec["frames"], 1)), rec["seed"])
    digest = hashlib.sha256(np.ascontiguousarray(Y).tobytes()).hexdigest()
    assert digest == rec["y_sha256"], "numpy legacy RNG stream changed: golden inputs cannot be re-created"
    return x, Y
