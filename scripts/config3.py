#!/usr/bin/env python
"""BASELINE config 3 at scale: the regular ensemble `1200_3_6_rand_ldpc_{1..10}` on BIAWGN, min-sum (codeword 1) and
sum-product (codeword 0) over the reference's SNR lists (src/simulations.py:32-33,36), a FIXED number of frames per
(code, SNR) (default 1e6, BASELINE.json) instead of min_wec, noise drawn on the GPU, counters kept on the GPU and
all-reduced once per point.  One process per GPU:

    python scripts/config3.py --frames 1000000 --out profiles/r2/config3.json
    torchrun --nproc-per-node 8 scripts/config3.py ...

Code tables come from tests/golden/codes.npz (the reference loader's edge lists): nothing is read from /root/reference.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000000)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--codes", type=int, default=10)
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import _golden as G
    from ldpc_decoders_b200 import Tables, biawgn, dist, sim, simulations
    comm = dist.Comm()
    torch.cuda.set_device(comm.local_rank)
    dt = np.float32 if args.dtype == "f32" else np.float64
    cases = [("MSA", 1, [float(v) for v in simulations.P_AWGN_MSA.split()]), ("SPA", 0, [float(v) for v in simulations.P_AWGN_SPA.split()])]
    out = {"config": "regular ensemble 1200_3_6_rand_ldpc_{1..%d}, BIAWGN, max_iter 10, %d frames per (code, SNR), messages %s, device noise"
                     % (args.codes, args.frames, args.dtype), "n_gpus": comm.world, "points": []}
    t_all = time.time()
    frames_done = 0
    for k in range(1, args.codes + 1):
        name = "1200_3_6_rand_ldpc_%d" % k
        tab = Tables(*G.code_tables(name))
        for decoder, cw, snrs in cases:
            x = np.zeros(tab.n, np.int64) + cw
            for snr in snrs:
                dec = getattr(biawgn, decoder)(snr, tab, max_iter=10, dtype=dt)
                t0 = time.time()
                r = sim.run_fixed_on_device(dec.simulate_round, dec.dec.engine.new_counters, x, comm, args.batch, args.frames,
                                            10, seed=1000 * k + int(round(snr * 100)))
                el = time.time() - t0
                frames_done += r["tot"]
                out["points"].append({"code": name, "decoder": decoder, "codeword": cw, "snr_db": snr, "tot": r["tot"], "wec": r["wec"],
                                      "wer": r["wer"], "bec": r["bec"], "ber": r["ber"], "mean_iters": r["dec"]["average"],
                                      "frames_per_s": r["tot"] / el})
    out["seconds"] = time.time() - t_all
    out["frames"] = frames_done
    out["frames_per_s_overall"] = frames_done / out["seconds"]
    if comm.rank == 0:
        print("config 3: %d points, %.3g frames in %.1f s (%.3g frames/s incl. Python, decoder construction and the per-point all-reduce)"
              % (len(out["points"]), frames_done, out["seconds"], out["frames_per_s_overall"]))
        # ensemble averages per (decoder, SNR), the curves the reference plots (src/graph.py)
        for decoder, _, snrs in cases:
            for snr in snrs:
                pts = [p for p in out["points"] if p["decoder"] == decoder and p["snr_db"] == snr]
                print("  %s %.2f dB: WER %.3e  BER %.3e  (min/max WER over codes %.3e / %.3e)"
                      % (decoder, snr, np.mean([p["wer"] for p in pts]), np.mean([p["ber"] for p in pts]),
                         min(p["wer"] for p in pts), max(p["wer"] for p in pts)))
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as fp:
                json.dump(out, fp, indent=1)
    comm.close()


if __name__ == "__main__":
    main()
