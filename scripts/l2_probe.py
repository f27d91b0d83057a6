#!/usr/bin/env python
"""How fast do the streaming sweeps run when their working set fits in the 126 MB L2?

Runs the HBM-streaming path (LDPC_PATH_STREAMING) of LDPC(1200,3,6) at 1.0 dB (no frame converges: fixed work) for a
range of batch sizes and prints the event-timed bandwidth of the check-node and variable-node sweeps.  Footprint of a
batch = frames x (E + n) x 4 B = frames x 19.2 KB.  A measurement aid for DESIGN.md section 4 (L2-resident tiles for
long codes), not a bench.
"""
import os
import sys

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
    nv = 10 ** (-1.0 / 10)
    steps = 20
    for frames in (512, 1024, 2048, 4096, 6144, 8192, 16384, 32768):
        g = torch.Generator(device="cuda").manual_seed(3)
        y = 1 + nv ** .5 * torch.randn((frames, tab.n), generator=g, device="cuda", dtype=torch.float32)
        res = {}

        def step():
            res["o"] = eng.decode_device_channel(lib.CH_BIAWGN, lib.MSA, lib.F32, nv, y, max_iter=10, out=res.get("o"),
                                                 flags=lib.PATH_STREAMING)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        eng.profile(True)
        eng.profile_read()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            step()
        t1.record()
        torch.cuda.synchronize()
        pr = eng.profile_read()
        eng.profile(False)
        ms = t0.elapsed_time(t1) / steps
        cn_b = 2 * tab.E * 4 * frames + tab.E * frames / 8
        vn_b = (2 * tab.E + tab.n) * 4 * frames + tab.n * frames / 8
        cn_us = pr["cn_ms"] / pr["cn_launches"] * 1e3
        vn_us = pr["vn_ms"] / pr["vn_launches"] * 1e3
        print("frames %6d  footprint %6.1f MB  step %8.3f ms  %6.2f M frames/s | cn %7.1f us %6.0f GB/s | vn %7.1f us %6.0f GB/s"
              % (frames, frames * (tab.E + tab.n) * 4 / 1e6, ms, frames / ms / 1e3, cn_us, cn_b / cn_us / 1e3,
                 vn_us, vn_b / vn_us / 1e3), flush=True)


if __name__ == "__main__":
    main()
