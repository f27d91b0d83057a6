#!/usr/bin/env python
"""Host-buffer throughput of ldpc_decode_host as a function of the chunk size (pipeline ramp / tail vs launch count),
next to a plain pinned H2D copy of the same bytes.  Measurement aid for DESIGN.md section 6, not the bench."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import _golden as G
    from ldpc_decoders_b200 import Tables, _lib as lib
    from ldpc_decoders_b200 import engine as eng_mod
    tab = Tables(*G.code_tables("1200_3_6_rand_ldpc_1"))
    eng = eng_mod.engine_for(tab)
    B = 32768
    nv = 10 ** (-2.0 / 10)
    rng = np.random.RandomState(1)
    Yp = eng_mod.pinned_empty((B, tab.n), np.float32)
    Yp[:] = 1 + np.sqrt(nv) * rng.standard_normal((B, tab.n)).astype(np.float32)
    xh, it, rs = (eng_mod.pinned_empty((B, tab.n), np.uint8), eng_mod.pinned_empty((B,), np.int32), eng_mod.pinned_empty((B,), np.uint8))
    d = torch.empty((B, tab.n), dtype=torch.float32, device="cuda")
    src = torch.from_numpy(Yp)
    for _ in range(3):
        d.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        d.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print("plain pinned H2D of the block: %.3f ms = %.1f GB/s -> ceiling %.2f M frames/s" % (dt * 1e3, Yp.nbytes / dt / 1e9, B / dt / 1e6))
    for chunk in (0, 1024, 2048, 2731, 4096, 5461, 8192, 16384, 32768):
        for _ in range(3):
            eng.decode_host(lib.CH_BIAWGN, lib.MSA, lib.F32, nv, Yp, max_iter=10, chunk=chunk, x_hat=xh, iters=it, reason=rs)
        t0 = time.perf_counter()
        for _ in range(10):
            eng.decode_host(lib.CH_BIAWGN, lib.MSA, lib.F32, nv, Yp, max_iter=10, chunk=chunk, x_hat=xh, iters=it, reason=rs)
        dt = (time.perf_counter() - t0) / 10
        print("chunk %6d: %.3f ms/step, %.2f M frames/s, H2D %.1f GB/s" % (chunk, dt * 1e3, B / dt / 1e6, Yp.nbytes / dt / 1e9), flush=True)


if __name__ == "__main__":
    main()
