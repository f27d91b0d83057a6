#!/usr/bin/env python
"""Run ONE decoder configuration on cuda:0 a few times and print its event-timed throughput.

A profiling aid (ncu -k regex:<kernel> python scripts/run_case.py ...), not the bench: bench.py is the number of record.

    python scripts/run_case.py --algo SPA --dtype f32 --snr 2.0 --frames 32768 --cw 0 --steps 5 [--streaming]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", default="1200_3_6_rand_ldpc_1")
    ap.add_argument("--algo", default="MSA", choices=["MSA", "SPA"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--channel", default="biawgn", choices=["biawgn", "bsc", "bec"])
    ap.add_argument("--snr", type=float, default=2.0, help="SNR in dB (biawgn) or crossover probability (bsc)")
    ap.add_argument("--frames", type=int, default=32768)
    ap.add_argument("--cw", type=int, default=1)
    ap.add_argument("--max-iter", type=int, default=10)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--streaming", action="store_true")
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--iters-out", default=None, help="save the per-frame iteration counts (.npy)")
    ap.add_argument("--n", type=int, default=0, help="synthetic (3,6) code of this length instead of --code")
    args = ap.parse_args()

    import torch
    import _golden as G
    from ldpc_decoders_b200 import Tables, _lib as lib, codes
    from ldpc_decoders_b200 import engine as eng_mod

    tab = codes.random_regular(args.n, 3, 6, seed=0).tables if args.n else Tables(*G.code_tables(args.code))
    eng = eng_mod.engine_for(tab)
    algo = lib.MSA if args.algo == "MSA" else lib.SPA
    dt = lib.F32 if args.dtype == "f32" else lib.F64
    flags = args.flags | (lib.PATH_STREAMING if args.streaming else 0)
    g = torch.Generator(device="cuda").manual_seed(3)
    if args.channel == "biawgn":
        nv = 10 ** (-args.snr / 10)
        y = (2 * args.cw - 1) + nv ** .5 * torch.randn((args.frames, tab.n), generator=g, device="cuda", dtype=torch.float32)
        ch, par = lib.CH_BIAWGN, nv
    elif args.channel == "bec":
        er = torch.rand((args.frames, tab.n), generator=g, device="cuda") < args.snr
        y = torch.where(er, 2, args.cw).to(torch.uint8)
        ch, par, algo = lib.CH_BEC, 0.0, lib.BEC
    else:
        flip = torch.rand((args.frames, tab.n), generator=g, device="cuda") < args.snr
        import math
        y = (flip ^ bool(args.cw)).to(torch.uint8)
        ch, par = lib.CH_BSC, math.log(1 - args.snr) - math.log(args.snr)      # the C ABI takes the LLR magnitude
    res = {}

    def step():
        res["o"] = eng.decode_device_channel(ch, algo, dt, par, y, max_iter=args.max_iter, out=res.get("o"), flags=flags)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / args.steps
    iters = res["o"]["iters"].cpu().numpy()
    x = res["o"]["x_hat"]
    if args.iters_out:
        np.save(args.iters_out, iters)
    print("%s %s %s %s=%g frames=%d: %.3f ms/step, %.3f M frames/s, mean iters %.2f, WER %.4f, %.3e edge updates/s"
          % (args.code if not args.n else "synthetic(3,6) n=%d" % args.n, args.algo, args.dtype, args.channel, args.snr,
             args.frames, ms, args.frames / ms / 1e3, iters.mean(), float((x != args.cw).any(dim=1).float().mean()),
             2 * tab.E * iters.sum() / (ms / 1e3)))


if __name__ == "__main__":
    main()
