#!/usr/bin/env python
"""Host <-> device copy bandwidth per GPU when N processes copy at the same time (one process per GPU, under torchrun),
next to ldpc_decode_host in the same two situations: separates what the box's PCIe / host-memory topology allows from
what the host-buffer decode path achieves at N GPUs.  One line per rank."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import _golden as G
    from ldpc_decoders_b200 import Tables, _lib as lib, dist as ldist
    from ldpc_decoders_b200 import engine as eng_mod
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    cpus = ldist.bind_near_gpu(local)
    dist.init_process_group("gloo")
    B, n = 32768, 1200
    h_in = torch.empty(B * n * 4, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(B * n, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    d_in = torch.empty_like(h_in, device="cuda")
    d_out = torch.empty_like(h_out, device="cuda")
    s2 = torch.cuda.Stream()

    def bw(mode, reps=20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if mode in ("h2d", "both"):
                d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return (h_in.numel() * reps / dt / 1e9 if mode != "d2h" else 0.0, h_out.numel() * reps / dt / 1e9 if mode != "h2d" else 0.0)

    tab = Tables(*G.code_tables("1200_3_6_rand_ldpc_1"))
    eng = eng_mod.engine_for(tab, local)
    nv = 10 ** (-2.0 / 10)
    Yp = eng_mod.pinned_empty((B, n), np.float32)
    Yp[:] = 1 + np.sqrt(nv) * np.random.RandomState(rank).standard_normal((B, n)).astype(np.float32)
    xh, it, rs = eng_mod.pinned_empty((B, n), np.uint8), eng_mod.pinned_empty((B,), np.int32), eng_mod.pinned_empty((B,), np.uint8)

    def dec(reps=10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.decode_host(lib.CH_BIAWGN, lib.MSA, lib.F32, nv, Yp, max_iter=10, x_hat=xh, iters=it, reason=rs)
        return B * reps / (time.perf_counter() - t0) / 1e6

    for m in ("h2d", "d2h", "both"):
        bw(m, 3)
    dec(2)
    alone = {}
    for r in range(world):
        dist.barrier()
        if r == rank:
            alone = {m: bw(m) for m in ("h2d", "d2h", "both")}
            alone["dec"] = dec()
    tog = {}
    for m in ("h2d", "d2h", "both"):
        dist.barrier()
        tog[m] = bw(m, 40)
    dist.barrier()
    tog["dec"] = dec(20)
    dist.barrier()
    fmt = lambda d: "h2d %.1f | d2h %.1f | both %.1f + %.1f GB/s | decode_host %.2f M frames/s" % (d["h2d"][0], d["d2h"][1], d["both"][0], d["both"][1], d["dec"])
    print("rank %d cpus %s\n   alone:        %s\n   all %d ranks:  %s" % (rank, len(cpus) if cpus else None, fmt(alone), world, fmt(tog)), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
