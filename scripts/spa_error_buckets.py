#!/usr/bin/env python
"""Observed error of the GPU sum-product check node against the REFERENCE's own messages (teacher-forced sweeps,
tests/golden/spa_tf.npz), per bucket of |ref| — the evidence behind the SPA tolerances ("report the observed
max error per |ref| bucket so the bound can be tightened instead of assumed").   python scripts/spa_error_buckets.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import _golden as G
    from ldpc_decoders_b200 import Tables, _lib as lib, engine
    edges = [0, 1, 2, 5, 10, 15, 20, 25, 30, 40, np.inf]
    acc = {dt: [[0.0, 0.0, 0] for _ in edges[:-1]] for dt in ("f64", "f32")}
    for rec, v2c, c2v in G.spa_tf():
        tab = Tables(*G.code_tables(rec["code"]))
        eng = engine.engine_for(tab)
        for dt, tdt in (("f64", torch.float64), ("f32", torch.float32)):
            out, _ = eng.debug_step(lib.SPA, 0, torch.from_numpy(np.ascontiguousarray(v2c)[None, :]).to("cuda", tdt))
            got = out.double().cpu().numpy()[0]
            ok = np.isfinite(c2v) & np.isfinite(got)
            d = np.abs(got - c2v)[ok]
            a = np.abs(c2v)[ok]
            for k in range(len(edges) - 1):
                m = (a >= edges[k]) & (a < edges[k + 1])
                if m.any():
                    acc[dt][k][0] = max(acc[dt][k][0], float(d[m].max()))
                    acc[dt][k][1] = max(acc[dt][k][1], float((d[m] / np.maximum(1.0, a[m])).max()))
                    acc[dt][k][2] += int(m.sum())
    print("GPU sum-product check-node sweep against the reference's c2v snapshots (spa_tf.npz), max |error| per |ref| bucket")
    print("%-12s %12s | %-26s | %-26s" % ("|ref| in", "messages", "float64 (formula mirror)", "float32 (hyperbolic pairs)"))
    print("%-12s %12s | %12s %13s | %12s %13s" % ("", "", "max abs", "max rel", "max abs", "max rel"))
    for k in range(len(edges) - 1):
        a, b = acc["f64"][k], acc["f32"][k]
        print("[%4g, %4g) %12d | %12.3e %13.3e | %12.3e %13.3e" % (edges[k], edges[k + 1], a[2], a[0], a[1], b[0], b[1]))
    print("bounds in the tests: float64 1e-12 * max(1, |ref|) + 1e-15 * exp(|ref|); float32 1e-4 * max(1, |ref|) for |ref| < 20")


if __name__ == "__main__":
    main()
