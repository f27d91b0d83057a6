#!/usr/bin/env python
"""Time the UNMODIFIED reference decoder (src/biawgn.py MSA -> src/bpa.py:86-102, scipy.sparse, one frame per call) on
this machine's cores: 1 process, then N independent processes (the reference's own parallel model, run_sims.sh:15).
BASELINE.md §4.  The reference is Python and does not travel to the GPU box, so this runs where /root/reference exists
(the build container) and writes profiles/reference_python_r2.json, which bench.py reports as
cpu_baseline.reference_python next to the live C-port baseline; with LDPC_REFERENCE=<path> bench.py runs it live.

    python scripts/time_reference_python.py [--ref /root/reference] [--frames 150] [--out profiles/reference_python_r2.json]
"""
import argparse
import json
import multiprocessing as mp
import os
import platform
import sys
import time

import numpy as np


def one_process(args):
    ref, code_name, snr, frames, seed, algo = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    sys.path.insert(0, os.path.join(ref, "src"))
    os.chdir(ref)
    import warnings
    warnings.simplefilter("ignore")
    if not hasattr(np, "int"):
        np.int = int                                   # math_utils.py:25 (numpy >= 1.24)
    import biawgn
    import codes
    code = codes.get_code(code_name)
    dec = getattr(biawgn, algo)(snr, code, max_iter=10)
    ch = biawgn.Channel(snr)
    x = code.parity_mtx[0] * 0 + (1 if algo == "MSA" else 0)
    np.random.seed(seed)
    for _ in range(10):                                # warm-up frames
        dec.decode(ch.send(x))
    t0 = time.perf_counter()
    errs = 0
    for _ in range(frames):
        errs += int((dec.decode(ch.send(x)) != x).sum() > 0)
    return frames / (time.perf_counter() - t0), errs


def measure(ref, frames, procs, code="1200_3_6_rand_ldpc_1", snr=2.0, algo="MSA"):
    r1, e1 = one_process((ref, code, snr, frames, 1, algo))
    with mp.get_context("spawn").Pool(procs) as pool:
        t0 = time.perf_counter()
        res = pool.map(one_process, [(ref, code, snr, frames, 100 + i, algo) for i in range(procs)])
        wall = time.perf_counter() - t0
    import scipy
    return {"workload": "LDPC(1200,3,6) %s, BIAWGN %.1f dB, %s, max_iter 10 (float64, scipy.sparse, one frame per decode call)" % (code, snr, algo),
            "single_process_frames_per_s": r1, "processes": procs, "aggregate_frames_per_s": float(sum(r for r, _ in res)),
            "aggregate_wall_frames_per_s": frames * procs / wall, "frames_per_process": frames,
            "wer_seen": (e1 + sum(e for _, e in res)) / (frames * (procs + 1)),
            "python": platform.python_version(), "numpy": np.__version__, "scipy": scipy.__version__,
            "cpu": platform.processor() or platform.machine(), "cores": os.cpu_count(), "blas_threads": 1}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("LDPC_REFERENCE", "/root/reference"))
    ap.add_argument("--frames", type=int, default=150)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    here = os.getcwd()
    rec = {"where": "build container (the reference does not travel to the GPU box)", "MSA": measure(a.ref, a.frames, a.procs, algo="MSA"),
           "SPA": measure(a.ref, a.frames, a.procs, algo="SPA")}
    os.chdir(here)
    print(json.dumps(rec, indent=1))
    if a.out:
        with open(a.out, "w") as fp:
            json.dump(rec, fp, indent=1)
