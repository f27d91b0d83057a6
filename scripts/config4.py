#!/usr/bin/env python
"""BASELINE config 4 at scale: the irregular ensemble `1200_rho_x5_rand_ldpc_{1..10}` on the BSC, sum-product, codeword 0,
p in {.1, .09, ..., .04} (src/simulations.py:35), max_iter in {1, 2, 3, 6, 10, 40, 100} (simulations.py:77), a fixed
number of frames per (code, p, max_iter), noise and counters on the GPU.  One process per GPU.

    python scripts/config4.py --frames 200000 --out profiles/r2/config4.json
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
    ap.add_argument("--frames", type=int, default=200000)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--codes", type=int, default=10)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import _golden as G
    from ldpc_decoders_b200 import Tables, bsc, dist, sim
    comm = dist.Comm()
    torch.cuda.set_device(comm.local_rank)
    ps = [round(.1 - .01 * k, 2) for k in range(7)]
    mis = [1, 2, 3, 6, 10, 40, 100]
    out = {"config": "irregular ensemble 1200_rho_x5_rand_ldpc_{1..%d}, BSC, sum-product float32, codeword 0, %d frames per (code, p, max_iter), device noise"
                     % (args.codes, args.frames), "n_gpus": comm.world, "points": []}
    t_all = time.time()
    frames_done = 0
    for k in range(1, args.codes + 1):
        name = "1200_rho_x5_rand_ldpc_%d" % k
        tab = Tables(*G.code_tables(name))
        x = np.zeros(tab.n, np.int64)
        for mi in mis:
            for p in ps:
                dec = bsc.SPA(p, tab, max_iter=mi, dtype=np.float32)
                t0 = time.time()
                r = sim.run_fixed_on_device(dec.simulate_round, dec.dec.engine.new_counters, x, comm, args.batch, args.frames,
                                            mi, seed=100000 * k + 1000 * mi + int(round(p * 100)))
                frames_done += r["tot"]
                out["points"].append({"code": name, "max_iter": mi, "p": p, "tot": r["tot"], "wec": r["wec"], "wer": r["wer"], "bec": r["bec"],
                                      "ber": r["ber"], "mean_iters": r["dec"]["average"], "frames_per_s": r["tot"] / (time.time() - t0)})
    out["seconds"] = time.time() - t_all
    out["frames"] = frames_done
    out["frames_per_s_overall"] = frames_done / out["seconds"]
    if comm.rank == 0:
        print("config 4: %d points, %.3g frames in %.1f s (%.3g frames/s overall)" % (len(out["points"]), frames_done, out["seconds"], out["frames_per_s_overall"]))
        for mi in mis:
            row = []
            for p in ps:
                pts = [q for q in out["points"] if q["max_iter"] == mi and q["p"] == p]
                row.append("%.2e" % np.mean([q["wer"] for q in pts]))
            print("  max_iter %3d: WER at p = %s: %s" % (mi, ps, " ".join(row)))
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as fp:
                json.dump(out, fp, indent=1)
    comm.close()


if __name__ == "__main__":
    main()
