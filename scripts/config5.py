#!/usr/bin/env python
"""BASELINE config 5: synthetic (3,6) regular LDPC n = 64800 (seeded configuration model, seed 0), BIAWGN, min-sum
float32, max_iter 10, 1e8 frames sharded over the GPUs of one box: each rank decodes its own frame indices (noise
keyed by the global frame index), counters stay on the device, ONE NCCL all-reduce ends the run.

    torchrun --nproc-per-node 8 scripts/config5.py --frames 100000000 --snr 2.5 --out profiles/r2/config5_n8.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=100000000)
    ap.add_argument("--snr", type=float, default=2.5)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--n", type=int, default=64800)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from ldpc_decoders_b200 import biawgn, codes, dist, sim
    comm = dist.Comm()
    torch.cuda.set_device(comm.local_rank)
    tab = codes.random_regular(args.n, 3, 6, seed=0).tables
    dec = biawgn.MSA(args.snr, tab, max_iter=10, dtype=np.float32)
    x = np.ones(tab.n, np.int64)
    eng = dec.dec.engine
    sim.run_fixed_on_device(dec.simulate_round, eng.new_counters, x, comm, args.batch, 2 * args.batch * comm.world, 10, seed=1)   # warm
    torch.cuda.synchronize()
    comm.barrier()
    seen = []
    t0 = time.time()
    r = sim.run_fixed_on_device(dec.simulate_round, eng.new_counters, x, comm, args.batch, args.frames, 10,
                                on_status=lambda tot, wec, bec, its, hist: seen.append((round(time.time() - t0, 1), tot, wec)),
                                log_freq=10., seed=5)
    torch.cuda.synchronize()
    el = time.time() - t0
    if comm.rank == 0:
        rate = r["tot"] / el
        bytes_per_frame_it = 4 * tab.E * 4 + tab.n * 4 + (tab.n + tab.E) / 8
        out = {"config": "synthetic (3,6) n=%d seed 0, BIAWGN %.2f dB, min-sum f32, max_iter 10, cw=1, device noise" % (tab.n, args.snr),
               "n_gpus": comm.world, "frames": r["tot"], "seconds": el, "frames_per_s": rate, "frames_per_s_per_gpu": rate / comm.world,
               "wec": r["wec"], "wer": r["wer"], "bec": r["bec"], "ber": r["ber"], "mean_iters": r["dec"]["average"], "iter_hist": r["dec"]["iter"],
               "edge_updates_per_s": 2 * tab.E * r["dec"]["average"] * rate,
               "hbm_GBps_per_gpu_algorithmic": rate / comm.world * r["dec"]["average"] * bytes_per_frame_it / 1e9,
               "progress_reports": seen[:12], "exchange": "one all_reduce(int64[17]) at the end + one per progress report"}
        print(json.dumps(out, indent=1))
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as fp:
                json.dump(out, fp, indent=1)
    comm.close()


if __name__ == "__main__":
    main()
