#!/usr/bin/env python
"""tests/golden/make_golden.py — freeze golden vectors from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/{codes,runs,spa_tf}.npz + kat.json

The reference is imported in place (nothing is copied or edited).  It is
instrumented from the outside through instance attributes only (SURVEY.md
Appendix B): ``decode_`` (one call per iteration, src/bpa.py:32), ``sum_cols``
(src/bpa.py:15,35) and, for BEC, the ``symbols`` lookup table (src/bec.py:75,119).

Fixtures:
  codes.npz   edge lists (np.where(H)) of every code the tests/bench use
  kat.json    the six Test.sample known-answer tests (src/bec.py:132-139,
              src/bsc.py:82-89, src/biawgn.py:85-92) with words, iteration counts, marginals
  runs.npz    seeded multi-frame runs: decoded words, iteration counts, exit reasons, and a
              per-frame flag for frames whose last marginal went non-finite in the reference
              (inputs are re-created from the seed by the tests; a sha256 of the inputs is stored)
  spa_tf.npz  teacher-forced SPA check-node sweeps (v2c in, c2v out) at float64
"""
import hashlib
import json
import os
import sys

import numpy as np

REF = os.environ.get("LDPC_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

os.environ.setdefault("FILE_CODES_DIR", os.path.join(REF, "data", "codes"))
sys.path.insert(0, os.path.join(REF, "src"))
np.int = int            # removed numpy aliases the reference still uses (math_utils.py:25, bec.py:34)
np.NINF = -np.inf

import bec      # noqa: E402
import biawgn   # noqa: E402
import bsc      # noqa: E402
import codes    # noqa: E402

MODELS = {"bsc": bsc, "bec": bec, "biawgn": biawgn}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# --------------------------------------------------------------------------- instrumentation
def run_bp(dec_adapter, y, dtype):
    """Decode one frame with a bsc/biawgn adapter; returns x_hat, iters, last marginal (pre-scrub)."""
    core = dec_adapter.dec
    calls, last_sum = [0], [None]
    orig_decode_, orig_sum_cols = core.decode_, core.sum_cols

    def decode_(*a):
        calls[0] += 1
        return orig_decode_(*a)

    def sum_cols(d):
        r = orig_sum_cols(d)
        last_sum[0] = r
        return r

    core.decode_, core.sum_cols = decode_, sum_cols
    try:
        if isinstance(dec_adapter, (bsc.SPA, bsc.MSA)):
            priors = dec_adapter.llr * (1 - 2 * y)                 # src/bsc.py:25
        else:
            priors = -2 * y / dec_adapter.noise_var                # src/biawgn.py:28
        priors = priors.astype(dtype)
        with np.errstate(all="ignore"):
            x_hat = core.decode(y, priors)
        marg = priors.copy() if last_sum[0] is None else priors + last_sum[0]
    finally:
        core.decode_, core.sum_cols = orig_decode_, orig_sum_cols
    return np.asarray(x_hat), calls[0], np.asarray(marg)


class _Rec(np.ndarray):
    """ndarray that logs every fancy lookup made through it (BEC: x_new = symbols[sign(marginal)])."""
    log = None

    def __getitem__(self, idx):
        r = np.asarray(super().__getitem__(idx))
        if isinstance(idx, np.ndarray) and self.log is not None:
            self.log.append(r.copy())
        return r


def run_bec(dec, y):
    """Decode one BEC frame; returns x_hat, iters, reason (0 decoded, 1 maximum, 2 stopping)."""
    orig = dec.symbols
    rec = np.asarray(orig).view(_Rec)
    rec.log = []
    dec.symbols = rec
    try:
        x_hat = np.asarray(dec.decode(y))
    finally:
        dec.symbols = orig
    rounds = len(rec.log)
    prev = rec.log[-2] if rounds >= 2 else y
    stopping = rounds >= 1 and bool((rec.log[-1] == prev).all())
    iters = rounds - 1 if stopping else rounds
    if stopping:
        reason = 2
    elif 0 < dec.max_iter <= iters:
        reason = 1
    else:
        reason = 0
    return x_hat, iters, reason


# --------------------------------------------------------------------------- fixtures
CODE_NAMES = (["4_2_test", "6_2_3_ldpc", "7_4_hamming", "12_3_4_ldpc"]
              + sorted(codes.get_file_code_map().keys()))


def make_codes():
    out = {}
    for name in CODE_NAMES:
        H = codes.get_code(name).parity_mtx
        rows, cols = np.where(H)
        out[name + "__shape"] = np.array(H.shape, np.int32)
        out[name + "__rows"] = rows.astype(np.uint16)
        out[name + "__cols"] = cols.astype(np.uint16)
    np.savez_compressed(os.path.join(HERE, "codes.npz"), **out)
    print("codes.npz:", len(CODE_NAMES), "codes")


KATS = [  # (channel, code, param, x, y)
    ("bec", "4_2_test", 1 / 3, [1, 1, 0, 1, 1], [1, 2, 0, 1, 2]),
    ("bec", "7_4_hamming", .1, [1, 0, 0, 1, 1, 0, 0], [2, 0, 2, 1, 1, 0, 2]),
    ("bsc", "4_2_test", 1 / 3, [1, 1, 0, 1, 1], [1, 0, 0, 1, 1]),
    ("bsc", "7_4_hamming", .1, [1, 0, 0, 1, 1, 0, 0], [1, 0, 1, 1, 1, 0, 0]),
    ("biawgn", "4_2_test", 1, [1, 1, 0, 1, 1], [1, 1, 1.6, .9, 1]),
    ("biawgn", "7_4_hamming", .1, [1, 0, 0, 1, 1, 0, 0], [1, -1, 1.1, 1, 1, -1, -1]),
]


def make_kats():
    out = []
    for channel, code_name, param, x, y in KATS:
        code = codes.get_code(code_name)
        y_ = np.array(y)
        for dec_name in (["SPA"] if channel == "bec" else ["SPA", "MSA"]):
            dec = getattr(MODELS[channel], dec_name)(param, code, max_iter=100)
            rec = dict(channel=channel, code=code_name, param=param, decoder=dec_name, max_iter=100,
                       x=list(map(int, x)), y=[float(v) for v in y])
            if channel == "bec":
                x_hat, iters, reason = run_bec(dec, y_)
                rec.update(x_hat=x_hat.astype(int).tolist(), iters=iters, reason=reason)
            else:
                x_hat, iters, marg = run_bp(dec, y_, np.float64)
                rec.update(x_hat=np.asarray(x_hat).astype(int).tolist(), iters=iters,
                           marg=[float(v) for v in marg])
            rec["passed"] = bool((np.asarray(rec["x_hat"]) == np.asarray(x)).all())
            out.append(rec)
    with open(os.path.join(HERE, "kat.json"), "w") as fp:
        json.dump(out, fp, indent=1)
    print("kat.json:", len(out), "cases; all passed =", all(r["passed"] for r in out))


RUNS = [  # (channel, code, decoder, param, codeword, max_iter, seed, frames, dtype)
    ("bsc", "1200_3_6_rand_ldpc_1", "MSA", .051, 1, 10, 101, 200, "f64"),
    ("bsc", "1200_3_6_rand_ldpc_1", "MSA", .051, 1, 10, 101, 200, "f32"),
    ("bsc", "1200_3_6_rand_ldpc_1", "MSA", .031, 1, 10, 102, 200, "f64"),
    ("bsc", "1200_3_6_rand_ldpc_1", "MSA", .031, 1, 10, 102, 200, "f32"),
    ("bsc", "1200_3_6_rand_ldpc_1", "MSA", .05, 1, 10, 103, 200, "f64"),
    ("bsc", "1200_3_6_rand_ldpc_1", "MSA", .05, 1, 10, 103, 200, "f32"),
    ("bsc", "1200_3_6_rand_ldpc_2", "MSA", .0451, 0, 40, 104, 100, "f64"),
    ("bsc", "1200_3_6_rand_ldpc_1", "SPA", .05, 0, 10, 105, 200, "f64"),
    ("bsc", "1200_3_6_rand_ldpc_1", "SPA", .06, 0, 10, 106, 200, "f64"),
    ("bsc", "1200_rho_x5_rand_ldpc_1", "SPA", .06, 0, 100, 107, 150, "f64"),
    ("bsc", "1200_rho_x5_rand_ldpc_10", "SPA", .05, 0, 10, 108, 150, "f64"),
    ("bsc", "1200_rho_x5_rand_ldpc_10", "MSA", .04, 1, 10, 109, 150, "f64"),
    ("bsc", "1200_rho_x5_rand_ldpc_10", "MSA", .04, 1, 10, 109, 150, "f32"),
    ("biawgn", "1200_3_6_rand_ldpc_1", "MSA", 1.0, 1, 10, 201, 150, "f64"),
    ("biawgn", "1200_3_6_rand_ldpc_1", "MSA", 2.0, 1, 10, 202, 200, "f64"),
    ("biawgn", "1200_3_6_rand_ldpc_1", "MSA", 2.0, 1, 10, 202, 200, "f32"),
    ("biawgn", "1200_3_6_rand_ldpc_1", "MSA", 3.0, 1, 10, 203, 200, "f64"),
    ("biawgn", "1200_3_6_rand_ldpc_1", "MSA", 3.0, 1, 10, 203, 200, "f32"),
    ("biawgn", "1200_3_6_rand_ldpc_1", "MSA", 2.0, 1, 100, 204, 80, "f64"),
    ("biawgn", "1200_3_6_rand_ldpc_3", "MSA", 2.5, 0, 10, 205, 150, "f64"),
    ("biawgn", "1200_3_6_rand_ldpc_1", "SPA", 2.0, 0, 10, 206, 200, "f64"),
    ("biawgn", "1200_3_6_rand_ldpc_1", "SPA", 3.0, 0, 10, 207, 200, "f64"),
    ("biawgn", "1200_3_6_rand_ldpc_1", "SPA", 1.0, 0, 10, 208, 100, "f64"),
    ("biawgn", "1200_rho_x5_rand_ldpc_1", "SPA", 2.0, 0, 40, 209, 100, "f64"),
    ("biawgn", "margulis", "MSA", 2.0, 1, 10, 210, 40, "f64"),
    ("biawgn", "margulis", "SPA", 2.0, 0, 10, 211, 40, "f64"),
    ("biawgn", "7_4_hamming", "SPA", 2.0, 1, 10, 212, 500, "f64"),
    ("biawgn", "7_4_hamming", "MSA", 4.0, 1, 10, 213, 500, "f64"),
    ("bsc", "7_4_hamming", "MSA", .1, 1, 10, 214, 500, "f64"),
    ("bsc", "7_4_hamming", "SPA", .1, 1, 10, 215, 500, "f64"),
    ("bsc", "12_3_4_ldpc", "SPA", .08, 0, 10, 216, 300, "f64"),
    ("biawgn", "512_3_6_rand_ldpc_1", "MSA", 2.0, 1, 10, 217, 150, "f32"),
    ("bec", "1200_3_6_rand_ldpc_1", "SPA", .5, 0, 10, 301, 150, "i"),
    ("bec", "1200_3_6_rand_ldpc_1", "SPA", .425, 0, 10, 302, 200, "i"),
    ("bec", "1200_3_6_rand_ldpc_1", "SPA", .4, 0, 10, 303, 300, "i"),
    ("bec", "1200_3_6_rand_ldpc_1", "SPA", .4, 0, 100, 304, 300, "i"),
    ("bec", "1200_3_6_rand_ldpc_1", "SPA", .4, 1, 0, 305, 200, "i"),
    ("bec", "1200_3_6_rand_ldpc_1", "SPA", .35, 0, 10, 306, 300, "i"),
    ("bec", "1200_3_6_rand_ldpc_1", "SPA", .3, 1, 10, 307, 300, "i"),
    ("bec", "1200_rho_x5_rand_ldpc_10", "SPA", .4, 0, 100, 308, 200, "i"),
    ("bec", "1200_rho_x5_rand_ldpc_1", "SPA", .45, 1, 10, 309, 200, "i"),
    ("bec", "7_4_hamming", "SPA", .3, 1, 10, 310, 500, "i"),
    ("bec", "margulis", "SPA", .4, 0, 100, 311, 60, "i"),
]


def run_key(r):
    ch, code, dec, param, cw, mi, seed, frames, dt = r
    return "%s|%s|%s|%g|%d|%d|%d|%d|%s" % (ch, code, dec, param, cw, mi, seed, frames, dt)


def make_runs():
    arrays, index = {}, []
    for i, r in enumerate(RUNS):
        ch, code_name, dec_name, param, cw, mi, seed, frames, dt = r
        model = MODELS[ch]
        code = codes.get_code(code_name)
        n = code.get_n()
        x = code.parity_mtx[0] * 0 + cw                       # src/main.py:18
        np.random.seed(seed)
        Y = model.Channel(param).send(np.tile(x, (frames, 1)))   # batched draw == sequential draws
        dec = getattr(model, dec_name)(param, code, max_iter=mi)
        xh = np.zeros((frames, n), np.uint8)
        iters = np.zeros(frames, np.int32)
        reason = np.zeros(frames, np.uint8)
        marg4 = []
        nonfinite = np.zeros(frames, np.uint8)
        for b in range(frames):
            if ch == "bec":
                x_hat, it, rs = run_bec(dec, Y[b])
            else:
                x_hat, it, marg = run_bp(dec, Y[b], np.float64 if dt == "f64" else np.float32)
                rs = 1 if 0 < mi <= it else 0
                nonfinite[b] = not np.isfinite(marg).all()
                if b < 4:
                    marg4.append(marg)
                if it == 0:
                    x_hat = (np.asarray(x_hat) != 0)           # 0-iteration exit returns y itself (bpa.py:20)
            xh[b] = np.asarray(x_hat).astype(np.uint8)
            iters[b], reason[b] = it, rs
        k = "r%02d" % i
        arrays[k + "_xhat"] = np.packbits(xh, axis=1) if ch != "bec" else xh
        arrays[k + "_iters"] = iters.astype(np.int16)
        arrays[k + "_reason"] = reason
        if marg4:
            arrays[k + "_marg4"] = np.stack(marg4)
            arrays[k + "_nonfinite"] = nonfinite   # frames whose last reference marginal holds inf/NaN
        errs = (xh != x[None, :]).sum(axis=1)
        index.append(dict(key=run_key(r), slot=k, channel=ch, code=code_name, decoder=dec_name, param=param,
                          codeword=cw, max_iter=mi, seed=seed, frames=frames, dtype=dt, n=int(n),
                          y_sha256=sha(Y), wec=int((errs > 0).sum()), bec=int(errs.sum()),
                          mean_iters=float(iters.mean())))
        print("%-60s wec=%4d mean_it=%.2f reasons=%s" % (run_key(r), index[-1]["wec"], iters.mean(),
                                                         np.bincount(reason, minlength=3).tolist()), flush=True)
    arrays["index_json"] = np.frombuffer(json.dumps(index).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "runs.npz"), **arrays)
    print("runs.npz:", len(index), "runs")


TF = [  # (channel, code, param, codeword, max_iter, seed, iterations to snapshot)
    ("biawgn", "1200_3_6_rand_ldpc_1", 2.0, 0, 10, 401, (0, 3, 7)),
    ("biawgn", "1200_3_6_rand_ldpc_1", 3.0, 0, 10, 402, (0, 2, 4)),
    ("bsc", "1200_rho_x5_rand_ldpc_1", .06, 0, 100, 403, (0, 3, 6)),
]


def make_spa_tf():
    arrays, index = {}, []
    for i, (ch, code_name, param, cw, mi, seed, snaps) in enumerate(TF):
        model = MODELS[ch]
        code = codes.get_code(code_name)
        x = code.parity_mtx[0] * 0 + cw
        np.random.seed(seed)
        y = model.Channel(param).send(x)
        dec = model.SPA(param, code, max_iter=mi)
        core = dec.dec
        log = []
        orig = core.decode_

        def decode_(v2c, xx, c2v, _log=log, _orig=orig):
            vin = np.array(v2c, np.float64)
            _orig(v2c, xx, c2v)
            _log.append((vin, np.array(c2v, np.float64)))

        core.decode_ = decode_
        with np.errstate(all="ignore"):
            dec.decode(y)
        core.decode_ = orig
        for it in snaps:
            if it < len(log):
                arrays["t%d_i%d_v2c" % (i, it)] = log[it][0]
                arrays["t%d_i%d_c2v" % (i, it)] = log[it][1]
                index.append(dict(slot="t%d_i%d" % (i, it), channel=ch, code=code_name, param=param, iteration=it))
        print("spa_tf", ch, code_name, param, "iterations run:", len(log))
    arrays["index_json"] = np.frombuffer(json.dumps(index).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "spa_tf.npz"), **arrays)


if __name__ == "__main__":
    what = sys.argv[1:] or ["codes", "kats", "runs", "spa_tf"]
    if "codes" in what:
        make_codes()
    if "kats" in what:
        make_kats()
    if "spa_tf" in what:
        make_spa_tf()
    if "runs" in what:
        make_runs()
