#!/usr/bin/env python
"""bench.py — decoded frames/s of the LDPC hot path on B200 (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json metric, config 3): LDPC(1200,3,6) `1200_3_6_rand_ldpc_1`, BIAWGN at 2.0 dB,
min-sum, float32 messages, max_iter 10, all-ones codeword (src/simulations.py:32-33), synthetic noise.
A step = one pass of the hot path over one batch of FRAMES frames per GPU: LLR front end + transpose,
10 x (check-node sweep, book-keeping, variable-node sweep) with per-frame early exit, hard decisions out.

  value     frames/s with the received block already resident in HBM (CUDA events on the launching stream)
  e2e       the same through the host-buffer entry point (ldpc_decode_host behind decode_batch): pinned
            host y in, x_hat / iteration counts out, copies inside the timed region; the K steps are submitted as a
            stream of batches (LDPC_HOST_ASYNC) and completed by one ldpc_host_sync — `blocking_call_value` is the
            same with K blocking calls
  roofline  the dominant kernel against the bound it really has.  On-chip path (what LDPC_PATH_AUTO runs for this code):
            the SHARED-MEMORY pipe — bytes the formulation moves through shared memory / event-timed kernel time against
            the LDS.128 rate measured on this GPU (tools/smem_peak); `traffic` = real DRAM bytes (ncu).  The old
            "algorithmic HBM bytes / time" figure (> 1 by design) is reported separately as `effective_hbm`.
  roofline_streaming   the HBM-streaming path on the same workload (results asserted identical): the HBM roofline
  msa_f64   the reference's own arithmetic (float64 messages, float64 rows): device rate, roofline, end-to-end rate
  spa       the other half of the metric: float32 sum-product on the same code / SNR / frames
  mc        the Monte-Carlo path at this GPU count: device noise rounds + ONE counter all-reduce (NCCL) per parameter,
            LDPC(1200,3,6) and the synthetic (3,6) n = 64800 code (config 5)
  e2e_variants   end-to-end with fewer PCIe bytes: binary16 rows, bit-packed BSC / BEC symbols (N = 1)
  cpu_baseline / --impl reference: the oracle port (oracle/ldpc_oracle.c, scalar C restatement of src/bpa.py)
            on the host cores — the reference itself is Python and does not travel to the GPU box; its own rate,
            measured where it exists, is attached as cpu_baseline.reference_python (scripts/time_reference_python.py).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CODE = "1200_3_6_rand_ldpc_1"
SNR_DB = 2.0
MAX_ITER = 10
METRIC = "decoded_frames_per_s"
UNIT = "frames/s"


def kernel_label(eng):
    """The on-chip kernel LDPC_PATH_AUTO launches in float32: ldpc_resident_kernel names the LAYOUT family
    ("resident_vp" = variable planes, regular or irregular); its kernel is resident_vq unless LDPC_RESIDENT_VP is set."""
    name = eng.resident_kernel
    if name == "resident_vp" and os.environ.get("LDPC_RESIDENT_VP") is None:
        return "resident_vq"
    return name


def load_code(name=CODE):
    import _golden as G
    return G.code_tables(name)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def smem_peak():
    """Measured LDS.128 rate of this GPU (tools/smem_peak, built by __graft_entry__.build()): GB/s and bytes/clk/SM.
    Falls back to the committed measurement (profiles/smem_peak_r2.json), then to the 128 B/clk/SM of the guides."""
    exe = os.path.join(ROOT, "tools", "smem_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        rec = json.loads(out.stdout.strip().splitlines()[-1])
        if "lds128_GBps" in rec:
            rec["source"] = "measured now (tools/smem_peak)"
            return rec
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "smem_peak_r2.json")) as fp:
            rec = json.load(fp)
        rec["source"] = "committed measurement (profiles/smem_peak_r2.json)"
        return rec
    except Exception:
        return {"lds128_bytes_per_clk_per_sm": 128.0, "source": "asserted (B300_MICROARCH.md)"}


def reference_python_record():
    """The real Python reference's rate: live when LDPC_REFERENCE points at a reference tree (never on the GPU box),
    else the measurement committed from the build container."""
    ref = os.environ.get("LDPC_REFERENCE")
    if ref and os.path.isdir(os.path.join(ref, "src")):
        try:
            out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "time_reference_python.py"), "--ref", ref,
                                  "--frames", "100"], capture_output=True, text=True, timeout=600)
            rec = json.loads(out.stdout)
            rec["where"] = "this machine, live (LDPC_REFERENCE)"
            return rec
        except Exception:
            pass
    try:
        with open(os.path.join(ROOT, "profiles", "reference_python_r2.json")) as fp:
            return json.load(fp)
    except Exception:
        return None


def traffic_per_launch(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fp:
            return json.load(fp).get(kernel)
    except Exception:
        return None


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val == "Active":
                    reasons.add(nm)
        self.tmp.close()
        os.unlink(self.tmp.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


def oracle_rate(tables, snr_db, frames, seconds, threads, dtype=np.float32):
    """frames/s of the oracle port (scalar C min-sum, one frame per thread at a time) on `threads` host threads."""
    from oracle import oracle as O
    m, n, rows, cols = tables
    g = O.Graph(m, n, rows, cols)
    rng = np.random.RandomState(7)
    std = np.sqrt(10 ** (-snr_db / 10))
    Y = 1.0 + rng.normal(0, std, (frames, n))
    pri = O.llr_biawgn(snr_db, Y).astype(dtype)
    O.bp_decode(g, O.MSA, pri[:max(64, threads)], max_iter=MAX_ITER, nthreads=threads)       # warm
    done, t0, its = 0, time.perf_counter(), 0
    while True:
        r = O.bp_decode(g, O.MSA, pri, max_iter=MAX_ITER, nthreads=threads)
        done += frames
        its += int(r["iters"].sum())
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return done / el, done, el, its / done


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    tables = load_code()
    threads = os.cpu_count() or 1
    frames = args.frames                                  # the SAME frames per step as the GPU arm (config identical)
    from oracle import oracle as O
    m, n, rows, cols = tables
    g = O.Graph(m, n, rows, cols)
    rng = np.random.RandomState(7)
    Y = 1.0 + rng.normal(0, np.sqrt(10 ** (-SNR_DB / 10)), (frames, n))
    pri = O.llr_biawgn(SNR_DB, Y).astype(np.float32)
    its = 0
    for _ in range(args.warmup):
        O.bp_decode(g, O.MSA, pri, max_iter=MAX_ITER, nthreads=threads)
    steps, el = 0, 0.0
    t0 = time.perf_counter()
    while steps < args.steps or el < 5.0:                 # at least K steps AND at least 5 s of work: a stable rate
        its += int(O.bp_decode(g, O.MSA, pri, max_iter=MAX_ITER, nthreads=threads)["iters"].sum())
        steps += 1
        el = time.perf_counter() - t0
    val = frames * steps / el
    args.steps = steps
    sample = "%d frames/step x %d steps (%.1f s) of the same workload (float32 min-sum, max_iter %d)" % (frames, steps, el, MAX_ITER)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(frames, world=max(1, args.gpus)),
        "edge_updates_per_s": 2 * len(rows) * its / el,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "reference_python": reference_python_record()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT, flush=True)


def workload_config(frames, world):
    return {"workload": "LDPC(1200,3,6) %s, BIAWGN %.1f dB, min-sum, max_iter %d, codeword=1, %d frames/step/GPU"
                        % (CODE, SNR_DB, MAX_ITER, frames),
            "code": CODE, "n": 1200, "m": 600, "E": 3600, "channel": "biawgn", "snr_db": SNR_DB, "decoder": "MSA",
            "max_iter": MAX_ITER, "frames_per_step_per_gpu": frames, "sharding": "frames x%d" % world,
            "l2": "inputs larger than L2 (message array %.0f MB per GPU vs 126 MB L2)" % (3600 * frames * 4 / 1e6)}


def algorithmic_bytes(E, n, s, iters, max_iter):
    """Per SURVEY.md 8(d): CN sweep 2*E*s + E/8, VN sweep 2*E*s + n*s + n/8 bytes per frame it processes.
    CN(it) runs for frames with iter_count >= it, VN(it) for frames with iter_count > it."""
    iters = np.asarray(iters, np.int64)
    cn_frames = int(np.minimum(iters + 1, max_iter).sum())
    vn_frames = int(iters.sum())
    return cn_frames * (2 * E * s + E / 8.0), vn_frames * (2 * E * s + n * s + n / 8.0)


def timed_steps(torch, fn, steps, warmup, dist):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    if dist is not None:
        dist.barrier()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def extra_workloads(torch, lib, eng_mod, Tables, peak):
    """Short side measurements (3 steps each) of the other §8 configurations; rank 0, N = 1 only."""
    out = []
    from ldpc_decoders_b200 import codes

    def bp_case(label, tab, algo, dtype, snr, frames, cw=1):
        eng = eng_mod.engine_for(tab)
        tdt = torch.float32
        nv = 10 ** (-snr / 10)
        g = torch.Generator(device="cuda").manual_seed(3)
        y = (2 * cw - 1) + nv ** .5 * torch.randn((frames, tab.n), generator=g, device="cuda", dtype=tdt)
        res = {}
        def fn():
            res["o"] = eng.decode_device_channel(lib.CH_BIAWGN, algo, dtype, nv, y, max_iter=MAX_ITER, out=res.get("o"))
        n0 = eng.launch_count
        ms = timed_steps(torch, fn, 3, 2, None)
        on_chip = (eng.launch_count - n0) == 5                   # one launch per decode (2 warm-up + 3 timed): the on-chip path
        iters = res["o"]["iters"].cpu().numpy()
        s = 4 if dtype == lib.F32 else 8
        cn_b, vn_b = algorithmic_bytes(tab.E, tab.n, s, iters, MAX_ITER)
        fps = frames * 3 / (ms / 1e3)
        frac = (cn_b + vn_b) * 3 / (ms / 1e3) / 1e9 / peak
        rec = {"workload": label, "value": fps, "unit": UNIT, "mean_iters": float(iters.mean()),
               "edge_updates_per_s": 2 * tab.E * float(iters.sum()) * 3 / (ms / 1e3),
               "path": ("on-chip (%s)" % ("resident_vq<double> / resident_vd" if dtype == lib.F64 else kernel_label(eng))) if on_chip else "streaming"}
        # streaming: fraction of the measured HBM peak the whole step reaches; on-chip: the same algorithmic bytes never
        # touch HBM, so the figure is an EFFECTIVE one (see roofline.note)
        rec["effective_hbm_frac" if on_chip else "step_hbm_frac"] = frac
        out.append(rec)

    tab = Tables(*load_code())
    # the headline workload in a four times larger batch: the launch's fixed cost (first rows in, last frames out: ~29 us) is
    # 3.5 % of a 32768-frame step and 0.9 % of this one
    bp_case("headline workload at 131072 frames per step: LDPC(1200,3,6) BIAWGN 2.0 dB MSA f32, max_iter 10", tab, lib.MSA, lib.F32, 2.0, 131072)
    bp_case("LDPC(1200,3,6) BIAWGN 2.0 dB MSA f64 (the reference's arithmetic), max_iter 10", tab, lib.MSA, lib.F64, 2.0, 32768)
    bp_case("LDPC(1200,3,6) BIAWGN 2.0 dB SPA f64 (formula mirror), max_iter 10, cw=0", tab, lib.SPA, lib.F64, 2.0, 16384, cw=0)
    big = codes.random_regular(64800, 3, 6, seed=0).tables
    bp_case("synthetic (3,6) n=64800 BIAWGN 1.0 dB MSA f32, max_iter 10 (no convergence)", big, lib.MSA, lib.F32, 1.0, 2048)
    bp_case("synthetic (3,6) n=64800 BIAWGN 2.5 dB MSA f32, max_iter 10", big, lib.MSA, lib.F32, 2.5, 2048)
    # config 4: the irregular n = 1200 ensemble on BSC (src/simulations.py:35,77), on-chip variable-plane kernel
    import math
    irr = Tables(*load_code("1200_rho_x5_rand_ldpc_1"))
    eng = eng_mod.engine_for(irr)
    frames, pflip = 32768, 0.06
    g = torch.Generator(device="cuda").manual_seed(5)
    yh = (torch.rand((frames, irr.n), generator=g, device="cuda") < pflip).to(torch.uint8)
    for algo, nm in ((lib.SPA, "SPA"), (lib.MSA, "MSA")):
        res = {}
        def fi():
            res["o"] = eng.decode_device_channel(lib.CH_BSC, algo, lib.F32, math.log(1 - pflip) - math.log(pflip), yh,
                                                 max_iter=MAX_ITER, out=res.get("o"))
        ms = timed_steps(torch, fi, 3, 2, None)
        iters = res["o"]["iters"].cpu().numpy()
        out.append({"workload": "irregular LDPC n=1200 (1200_rho_x5_rand_ldpc_1) BSC p=0.06 %s f32, max_iter 10, cw=0 (%s)"
                                % (nm, kernel_label(eng) or "streaming"),
                    "value": frames * 3 / (ms / 1e3), "unit": UNIT, "mean_iters": float(iters.mean()),
                    "edge_updates_per_s": 2 * irr.E * float(iters.sum()) * 3 / (ms / 1e3),
                    "path": ("on-chip (%s)" % kernel_label(eng)) if eng.resident_kernel else "streaming"})
    # Monte-Carlo round entirely on the GPU (on-device Philox channel + decode + error count): what sim.py --noise device runs
    eng = eng_mod.engine_for(tab)
    frames = 32768
    xone = torch.ones(tab.n, dtype=torch.uint8, device="cuda")
    bufs, rs, cnt = {}, {}, [0]
    def fs():
        rs["o"] = eng.simulate(lib.CH_BIAWGN, lib.MSA, lib.F32, 10 ** (-SNR_DB / 10), frames, seed=1, frame0=cnt[0] * frames,
                               x=xone, max_iter=MAX_ITER, bufs=bufs)
        cnt[0] += 1
    ms = timed_steps(torch, fs, 5, 2, None)
    iters = rs["o"]["iters"].cpu().numpy()
    out.append({"workload": "Monte-Carlo round on the GPU: Philox BIAWGN 2.0 dB noise + MSA f32 decode + bit-error count, LDPC(1200,3,6), cw=1",
                "value": frames * 5 / (ms / 1e3), "unit": UNIT, "mean_iters": float(iters.mean()),
                "wer": float((rs["o"]["bit_errs"] > 0).float().mean().item()),
                "edge_updates_per_s": 2 * tab.E * float(iters.sum()) * 5 / (ms / 1e3)})
    # BEC, config 2
    eng = eng_mod.engine_for(tab)
    frames = 131072
    g = torch.Generator(device="cuda").manual_seed(4)
    yb = torch.where(torch.rand((frames, tab.n), generator=g, device="cuda") < 0.4, 2, 0).to(torch.uint8)
    res = {}
    def fb():
        res["o"] = eng.decode_device_channel(lib.CH_BEC, lib.BEC, lib.F32, 0.0, yb, max_iter=MAX_ITER, out=res.get("o"))
    ms = timed_steps(torch, fb, 3, 2, None)
    iters = res["o"]["iters"].cpu().numpy()
    out.append({"workload": "LDPC(1200,3,6) BEC p=0.40 erasure decoding (on-chip bit planes, resident_bec), max_iter 10, cw=0, 131072 frames",
                "value": frames * 3 / (ms / 1e3), "unit": UNIT, "mean_iters": float(iters.mean()),
                "edge_updates_per_s": 2 * tab.E * float(iters.sum()) * 3 / (ms / 1e3)})
    return out


RESULT = sys.stdout


def claim_stdout():
    """stdout carries exactly one JSON line.  Libraries write there too (NCCL prints its version banner on fd 1 at
    communicator creation, whatever NCCL_DEBUG_FILE says), so keep a private copy of fd 1 for the result and point fd 1
    itself at stderr for everybody else."""
    global RESULT
    sys.stdout.flush()
    RESULT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=32768, help="frames per step per GPU")
    ap.add_argument("--flags", type=int, default=0, help="ldpc_decode flags (8 = register-staged check-node sweep)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-mc", action="store_true")
    ap.add_argument("--mc-rounds", type=int, default=96, help="Monte-Carlo rounds per GPU of the mc leg (n = 1200; n = 64800 runs an eighth)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    from ldpc_decoders_b200 import dist as ldist
    ldist.quiet_nccl_stdout()                         # keep stdout to the one JSON line (NCCL prints its version there)
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    numa_cpus = ldist.bind_near_gpu(local_rank)       # the feeding thread on the CPUs next to the GPU ...
    numa_node = ldist.bind_memory_near_gpu(local_rank)   # ... and the pinned buffers allocated below on its NUMA node
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from ldpc_decoders_b200 import Tables, _lib as lib
    from ldpc_decoders_b200 import engine as eng_mod
    from ldpc_decoders_b200.engine import pinned_empty

    tables = load_code()
    tab = Tables(*tables)
    eng = eng_mod.engine_for(tab)
    B = args.frames
    nv = 10 ** (-SNR_DB / 10)
    peak, peak_src = measured_peak()

    # ---- synthetic received block, resident in HBM: y = (2x-1) + sigma * N(0,1), x = all ones; seed by global rank
    g = torch.Generator(device="cuda").manual_seed(1000 + rank)
    y = 1.0 + nv ** .5 * torch.randn((B, tab.n), generator=g, device="cuda", dtype=torch.float32)
    def measure(flags, algo=lib.MSA, y=y, dtype=lib.F32):
        """K timed steps of the device-resident hot path under `flags`; per-launch events recorded inside."""
        res = {}

        def step():
            res["o"] = eng.decode_device_channel(lib.CH_BIAWGN, algo, dtype, nv, y, max_iter=MAX_ITER,
                                                 out=res.get("o"), flags=flags)

        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        eng.profile(True)
        eng.profile_read()
        l0 = eng.launch_count
        ms = timed_steps(torch, step, args.steps, 0, dist)
        launches = eng.launch_count - l0
        prof = eng.profile_read()
        eng.profile(False)
        return dict(ms=ms, launches=launches, prof=prof, iters=res["o"]["iters"].cpu().numpy(), x_hat=res["o"]["x_hat"].clone())

    def sum_over_ranks(v):
        if dist is None:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    def max_over_ranks(v):
        if dist is None:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 and os.environ.get("LDPC_BENCH_NO_SAMPLER") != "1" else None
    main = measure(args.flags)
    resident = main["prof"]["vn_launches"] == 0          # the on-chip path is ONE kernel per decode
    ms, launches, prof, iters, x_hat = main["ms"], main["launches"], main["prof"], main["iters"], main["x_hat"]
    wer = float((x_hat != 1).any(dim=1).float().mean().item())

    total_frames = B * args.steps * world
    value = total_frames / (ms / 1e3)
    it_sum = float(iters.sum())
    it_sum_all = sum_over_ranks(it_sum)
    edge_updates = 2 * tab.E * it_sum_all * args.steps / (ms / 1e3)
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    speak = smem_peak() if rank == 0 else {"lds128_bytes_per_clk_per_sm": 128.0, "source": "not measured on this rank"}

    def streaming_roofline(m_, s_=4):
        """Roofline of the dominant sweep kernel of a streaming run (this rank), from the per-launch events."""
        pr, it_ = m_["prof"], m_["iters"]
        cn_bytes, vn_bytes = algorithmic_bytes(tab.E, tab.n, s_, it_, MAX_ITER)
        cn_gbs = cn_bytes * args.steps / (pr["cn_ms"] / 1e3) / 1e9 if pr["cn_ms"] > 0 else 0.0
        vn_gbs = vn_bytes * args.steps / (pr["vn_ms"] / 1e3) / 1e9 if pr["vn_ms"] > 0 else 0.0
        dom = "vn_sweep" if pr["vn_ms"] >= pr["cn_ms"] else "cn_sweep"
        achieved = vn_gbs if dom == "vn_sweep" else cn_gbs
        return {
            "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_src, "traffic": traffic_per_launch(dom) if s_ == 4 else None,
            "algorithmic_bytes_per_launch": ((vn_bytes if dom == "vn_sweep" else cn_bytes) * args.steps
                                             / max(1, pr["vn_launches"] if dom == "vn_sweep" else pr["cn_launches"])),
            "avg_launch_ms": (pr["vn_ms"] / max(1, pr["vn_launches"])) if dom == "vn_sweep" else (pr["cn_ms"] / max(1, pr["cn_launches"])),
            "cn_sweep": {"ms_total": pr["cn_ms"], "launches": pr["cn_launches"], "GBps": cn_gbs, "frac": cn_gbs / peak},
            "vn_sweep": {"ms_total": pr["vn_ms"], "launches": pr["vn_launches"], "GBps": vn_gbs, "frac": vn_gbs / peak},
            "sweeps_share_of_step": (pr["cn_ms"] + pr["vn_ms"]) / m_["ms"],
            "step_frac": (cn_bytes + vn_bytes) * args.steps / (m_["ms"] / 1e3) / 1e9 / peak,
            "bytes_per_edge_iteration": 17.5 if s_ == 4 else 34.83,
        }

    def resident_roofline(m_, kernel, s_=4, limiter=""):
        """The on-chip kernels keep every frame in shared memory for all its iterations, so their bound is the
        SHARED-MEMORY pipe, not HBM.  Per frame-iteration the formulation moves 3E + 2n message scalars through shared
        memory (check phase: E marginal gathers + E message stores; variable phase: E message loads + n prior loads + n
        marginal stores).  achieved = those bytes x frame-iterations of the launch / event-timed kernel time; peak = the
        conflict-free LDS.128 rate measured on this GPU (tools/smem_peak).  DRAM sees only `traffic`.
        Returns (roofline, effective_hbm)."""
        pr, it_ = m_["prof"], m_["iters"]
        cn_bytes, vn_bytes = algorithmic_bytes(tab.E, tab.n, s_, it_, MAX_ITER)
        k_ms = pr["cn_ms"]
        n_l = max(1, pr["cn_launches"])
        fi_per_s = float(it_.sum()) * args.steps / (k_ms / 1e3) if k_ms > 0 else 0.0
        smem_bytes = (3 * tab.E + 2 * tab.n) * s_
        achieved = smem_bytes * fi_per_s / 1e9
        if "lds128_GBps" in speak:
            pk, pk_src = float(speak["lds128_GBps"]), "%s: LDS.128 %.0f GB/s = %.1f B/clk/SM at %.0f MHz" % (
                speak["source"], speak["lds128_GBps"], speak["lds128_bytes_per_clk_per_sm"], speak.get("sm_clock_mhz_max", 0))
        else:
            pk, pk_src = speak["lds128_bytes_per_clk_per_sm"] * sm_count * 1.965, speak["source"] + " x 1965 MHz"
        io_bytes = B * (tab.n * 4 + tab.n + 5)
        roof = {
            "bound": "smem", "kernel": kernel, "achieved": achieved, "peak": pk, "unit": "GB/s", "frac": achieved / pk,
            "peak_source": pk_src, "traffic": traffic_per_launch(kernel),
            "algorithmic_bytes_per_launch": smem_bytes * float(it_.sum()), "bytes_per_frame_iteration": smem_bytes,
            "avg_launch_ms": k_ms / n_l, "kernel_share_of_step": k_ms / m_["ms"], "frame_iterations_per_s": fi_per_s,
            "note": "shared-memory roofline of the on-chip kernel (bytes through shared memory / measured LDS.128 rate). " + limiter,
            "compulsory_hbm_bytes_per_launch": io_bytes,
            "compulsory_hbm_GBps": io_bytes * args.steps / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0,
        }
        eff = (cn_bytes + vn_bytes) * args.steps / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0
        effective = {"GBps": eff, "over_hbm_peak": eff / peak, "hbm_peak": peak, "peak_source": peak_src,
                     "note": "algorithmic bytes of the STREAMING layout (SURVEY 8d: %d B per frame-iteration) / kernel time: what an "
                             "HBM-streaming decoder would have to sustain for this rate; not a roofline (the bytes never touch HBM)"
                             % ((4 * tab.E + tab.n) * s_ + (tab.n + tab.E) // 8)}
        return roof, effective

    streaming = None
    effective_hbm = None
    # regular codes in the two-CTA geometry run resident_vq (resident_vp with the frame hand-over fused into the variable
    # phase) unless LDPC_RESIDENT_VP=1 asks for the older kernel; ldpc_resident_kernel names the family
    res_name = kernel_label(eng)
    if resident:
        roofline, effective_hbm = resident_roofline(main, res_name, 4,
                                                    "ncu (profiles/, r2e): no pipe saturated - shared-memory wavefronts 64 %, issue 58 %, ALU 46 %, FMA 15 %.")
        roofline["shared_memory_plan"] = eng.resident_plan()
        sm = measure(args.flags | lib.PATH_STREAMING)
        assert (sm["iters"] == iters).all() and bool((sm["x_hat"] == x_hat).all()), "streaming and resident paths disagree"
        streaming = {"value": total_frames / (sm["ms"] / 1e3), "unit": UNIT, "ms_per_step": sm["ms"] / args.steps,
                     "gpu_launches": int(sm["launches"]), "roofline": streaming_roofline(sm)}
    else:
        roofline = streaming_roofline(main)

    # ---- the reference's own arithmetic: float64 messages (bpa.py computes in the dtype of its priors, and its front
    # ends produce float64), float64 received rows; same code / SNR / frames, min-sum
    y64 = y.double()
    f64_name = "resident_vq<double>" if res_name == "resident_vq" else "resident_vd"     # the float64 instance of the same kernel
    f64 = measure(args.flags, y=y64, dtype=lib.F64)
    f64_res = f64["prof"]["vn_launches"] == 0
    msa_f64 = {"workload": "same code / SNR / frames, min-sum with float64 messages and float64 rows (the reference's arithmetic, bit-exact with the float64 oracle)",
               "value": total_frames / (f64["ms"] / 1e3), "unit": UNIT, "ms_per_step": f64["ms"] / args.steps, "dtype": "f64",
               "mean_iters": float(f64["iters"].mean()), "gpu_launches": int(f64["launches"]),
               "edge_updates_per_s": 2 * tab.E * sum_over_ranks(float(f64["iters"].sum())) * args.steps / (f64["ms"] / 1e3),
               "path": ("on-chip (%s)" % f64_name) if f64_res else "streaming"}
    if f64_res:
        msa_f64["roofline"], msa_f64["effective_hbm"] = resident_roofline(f64, f64_name, 8, "Two frames per 16-byte cell; DSETP/select minima.")
    else:
        msa_f64["roofline"] = streaming_roofline(f64, 8)

    # ---- the other half of the metric: sum-product (float32), same code / SNR / frames, all-zero codeword
    # (src/simulations.py:36 runs SPA with --codeword 0)
    y_spa = y - 2.0
    sp = measure(args.flags, algo=lib.SPA, y=y_spa)
    sp_it = float(sp["iters"].sum())
    sp_it_all = sum_over_ranks(sp_it)
    spa = {"workload": "same code, BIAWGN %.1f dB, sum-product float32 (hyperbolic-pair rule), max_iter %d, codeword=0" % (SNR_DB, MAX_ITER),
           "value": total_frames / (sp["ms"] / 1e3), "unit": UNIT, "ms_per_step": sp["ms"] / args.steps,
           "mean_iters": sp_it / B, "edge_updates_per_s": 2 * tab.E * sp_it_all * args.steps / (sp["ms"] / 1e3),
           "wer": float((sp["x_hat"] != 0).any(dim=1).float().mean().item()), "gpu_launches": int(sp["launches"])}
    if sp["prof"]["vn_launches"] == 0:
        spa["roofline"], spa["effective_hbm"] = resident_roofline(sp, res_name, 4, "Sum-product adds the MUFU pipe (18 ex2/lg2 per check and frame) to the limiters.")
    else:
        spa["roofline"] = streaming_roofline(sp)
    del y_spa

    # ---- Monte-Carlo path at this GPU count: every rank runs device-noise rounds on its own frame indices,
    # counters stay on the device, ONE all-reduce (NCCL) ends the parameter.  Device-timed, max over ranks.
    from ldpc_decoders_b200 import biawgn as gbiawgn, codes as gcodes, sim as gsim
    comm = ldist.Comm()

    def mc_leg(label, code_tab, snr, batch, rounds):
        dec = gbiawgn.MSA(snr, code_tab, max_iter=MAX_ITER, dtype=np.float32)
        x1 = np.ones(code_tab.n, np.int64)
        mc_eng = dec.dec.engine
        frames_all = rounds * batch * world
        gsim.run_fixed_on_device(dec.simulate_round, mc_eng.new_counters, x1, comm, batch, 2 * batch * world, MAX_ITER, seed=11)  # warm
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        l0 = mc_eng.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = gsim.run_fixed_on_device(dec.simulate_round, mc_eng.new_counters, x1, comm, batch, frames_all, MAX_ITER, seed=12)
        e1.record()
        torch.cuda.synchronize()
        mc_ms = max_over_ranks(e0.elapsed_time(e1))
        assert r["tot"] == frames_all
        return {"workload": label, "value": frames_all / (mc_ms / 1e3), "unit": UNIT, "frames": frames_all, "ms": mc_ms,
                "frames_per_round_per_gpu": batch, "rounds": rounds, "n_gpus": world, "wer": r["wer"], "ber": r["ber"],
                "mean_iters": r["dec"]["average"], "gpu_launches": int(mc_eng.launch_count - l0),
                "edge_updates_per_s": 2 * code_tab.E * r["dec"]["average"] * frames_all / (mc_ms / 1e3),
                "exchange": "one all_reduce(int64[%d]) per parameter (%s); no other traffic between GPUs, nothing crosses PCIe inside the loop"
                            % (4 + gsim.hist_bins(MAX_ITER) + 1, "NCCL" if world > 1 else "single GPU: none")}

    mc = {}
    if not args.no_mc:
        try:
            mc["n1200"] = mc_leg("Monte-Carlo: Philox BIAWGN %.1f dB noise + MSA f32 decode + error counters on the device, LDPC(1200,3,6) %s, cw=1"
                                 % (SNR_DB, CODE), tab, SNR_DB, B, args.mc_rounds)
            big = gcodes.random_regular(64800, 3, 6, seed=0).tables
            mc["n64800"] = mc_leg("Monte-Carlo (config 5): synthetic (3,6) n=64800 (seed 0), BIAWGN 2.5 dB, MSA f32, cw=1, streaming path",
                                  big, 2.5, 2048, max(2, args.mc_rounds // 8))
            mc["n64800"]["step_hbm_frac"] = (mc["n64800"]["value"] / world) * mc["n64800"]["mean_iters"] * 3402000.0 / 1e9 / peak
            del big
        except Exception as exc:                # a side leg must not lose the headline line (every rank fails alike: no hang)
            mc["error"] = repr(exc)

    # ---- e2e: host buffers through the host entry point (what decode_batch calls), copies inside the timed region
    def e2e_leg(label, channel, algo, dtype, param, Yh_, packed_in=False, packed_out=True, ref=None):
        """K steps, each with its own H2D of the received block and D2H of the words inside the timed region, submitted
        as a stream of batches (decode_host(wait=False) x K, one host_sync) and, for comparison, as K blocking calls."""
        prow = (2 if channel == lib.CH_BEC else 1) * eng_mod.packed_row_bytes(tab.n)
        xh_ = pinned_empty((B, prow if packed_out else tab.n), np.uint8)
        ith_, rsh_ = pinned_empty((B,), np.int32), pinned_empty((B,), np.uint8)

        def one(wait=True):
            eng.decode_host(channel, algo, dtype, param, Yh_, max_iter=MAX_ITER, x_hat=xh_, iters=ith_, reason=rsh_,
                            flags=args.flags, wait=wait, packed_in=packed_in, packed_out=packed_out)

        def run(stream_of_batches):
            for _ in range(2):
                one()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            l0 = eng.launch_count
            t0_ = time.perf_counter()
            for _ in range(args.steps):
                one(wait=not stream_of_batches)
            eng.host_sync()
            torch.cuda.synchronize()
            return time.perf_counter() - t0_, eng.launch_count - l0

        blk_el = max_over_ranks(run(False)[0])
        el_, launches_ = run(True)
        rank_ms = [1e3 * el_ / args.steps]
        if dist is not None:
            t = torch.zeros(world, dtype=torch.float64, device="cuda")
            t[rank] = el_
            dist.all_reduce(t)
            rank_ms = [1e3 * float(v) / args.steps for v in t.tolist()]
            el_ = float(t.max().item())
        if ref is not None:
            got = eng_mod.unpack_bits(xh_, tab.n) if (packed_out and channel != lib.CH_BEC) else xh_
            assert (ith_ == ref[0]).all() and bool((torch.from_numpy(np.ascontiguousarray(got)).cuda() == ref[1]).all()), "e2e result differs from device path"
        h2d, d2h = int(Yh_.nbytes), int(xh_.nbytes + ith_.nbytes + rsh_.nbytes)
        return {"workload": label, "value": total_frames / el_, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * el_ / args.steps, "blocking_call_value": total_frames / blk_el,
                "h2d_GBps_by_rank": [h2d / (m_ / 1e3) / 1e9 for m_ in rank_ms], "ms_per_step_by_rank": rank_ms,
                "bytes_per_frame": (h2d + d2h) / B, "gpu_launches": int(launches_)}

    Yh = pinned_empty((B, tab.n), np.float32)
    Yh[...] = y.cpu().numpy()
    e2e = e2e_leg("headline workload: pinned float32 rows in (4 B per received value), bit-packed words + iteration counts + reasons out",
                  lib.CH_BIAWGN, lib.MSA, lib.F32, nv, Yh, ref=(iters, x_hat))
    e2e.update({"api": "Engine.decode_host(wait=False, packed_out=True) x K + host_sync -> ldpc_decode_host with LDPC_HOST_ASYNC | LDPC_OUT_PACKED "
                       "(every step copies its own input and output; the decoded words cross PCIe as bits, losslessly)",
                "timer": "host perf_counter around the submitting calls + host_sync, max over ranks",
                "host_cpus_bound": (len(numa_cpus) if numa_cpus else None), "numa_node": numa_node})
    e_launches = e2e.pop("gpu_launches")
    e2e_variants = []
    if world == 1 and not args.no_extras:
        try:
            Y16 = pinned_empty((B, tab.n), np.float16)
            Y16[...] = Yh.astype(np.float16)
            e2e_variants.append(e2e_leg("same frames rounded to binary16 rows (LDPC_F16, 2 B per received value; priors = (-2 float64(y)) / var on those values), MSA f32",
                                        lib.CH_BIAWGN, lib.MSA, lib.F32, nv, Y16))
            Y64 = pinned_empty((B, tab.n), np.float64)
            Y64[...] = Yh.astype(np.float64)
            e2e_variants.append(e2e_leg("float64 rows in, float64 messages (the reference's own types end to end), MSA f64",
                                        lib.CH_BIAWGN, lib.MSA, lib.F64, nv, Y64, ref=(f64["iters"], f64["x_hat"])))
            msa_f64["e2e"] = {k: e2e_variants[-1][k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "ms_per_step", "h2d_GBps_by_rank")}
            del Y16, Y64
            pflip = 0.05
            gb = torch.Generator(device="cuda").manual_seed(5 + rank)
            yb = ((torch.rand((B, tab.n), generator=gb, device="cuda") < pflip) ^ True).to(torch.uint8)      # codeword 1
            llr = math.log(1 - pflip) - math.log(pflip)
            rb = eng.decode_device_channel(lib.CH_BSC, lib.MSA, lib.F32, llr, yb, max_iter=MAX_ITER)
            ms_b = timed_steps(torch, lambda: eng.decode_device_channel(lib.CH_BSC, lib.MSA, lib.F32, llr, yb, max_iter=MAX_ITER, out=rb), args.steps, 2, None)
            Pb = pinned_empty((B, eng_mod.packed_row_bytes(tab.n)), np.uint8)
            Pb[...] = eng_mod.pack_bits(yb.cpu().numpy())
            v = e2e_leg("BSC p=0.05 min-sum f32, cw=1: bit-packed hard bits in (LDPC_IN_PACKED), bit-packed words out", lib.CH_BSC, lib.MSA,
                        lib.F32, llr, Pb, packed_in=True, ref=(rb["iters"].cpu().numpy(), rb["x_hat"]))
            v["device_resident_value"] = B * args.steps / (ms_b / 1e3)
            v["e2e_over_device"] = v["value"] / v["device_resident_value"]
            e2e_variants.append(v)
            Fb = 4 * B
            ye = torch.where(torch.rand((Fb, tab.n), generator=gb, device="cuda") < 0.4, 2, 0).to(torch.uint8)
            re_ = eng.decode_device_channel(lib.CH_BEC, lib.BEC, lib.F32, 0.0, ye, max_iter=MAX_ITER)
            ms_e = timed_steps(torch, lambda: eng.decode_device_channel(lib.CH_BEC, lib.BEC, lib.F32, 0.0, ye, max_iter=MAX_ITER, out=re_), args.steps, 2, None)
            Pe = pinned_empty((Fb, 2 * eng_mod.packed_row_bytes(tab.n)), np.uint8)
            Pe[...] = eng_mod.pack_symbols(ye.cpu().numpy())
            B_keep, total_keep = B, total_frames
            B, total_frames = Fb, Fb * args.steps
            v = e2e_leg("BEC p=0.40 erasure decoding, cw=0, %d frames/step: two bit planes per symbol in and out (LDPC_IN_PACKED | LDPC_OUT_PACKED)" % Fb,
                        lib.CH_BEC, lib.BEC, lib.F32, 0.0, Pe, packed_in=True)
            B, total_frames = B_keep, total_keep
            v["device_resident_value"] = Fb * args.steps / (ms_e / 1e3)
            v["e2e_over_device"] = v["value"] / v["device_resident_value"]
            e2e_variants.append(v)
            del yb, ye, Pb, Pe
        except Exception as exc:                # side measurements must not lose the headline line
            e2e_variants.append({"error": repr(exc)})

    clocks = sampler.stop() if sampler is not None else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, world),
            "edge_updates_per_s": edge_updates, "mean_iters": it_sum / B, "wer": wer,
            "path": "resident (on-chip, LDPC_PATH_AUTO)" if resident else "streaming",
            "roofline": roofline, "effective_hbm": effective_hbm, "roofline_streaming": streaming, "msa_f64": msa_f64, "spa": spa,
            "mc": mc, "e2e": e2e, "e2e_variants": e2e_variants, "clocks": clocks,
            "gpu_launches": int(launches), "gpu_launches_e2e": int(e_launches),
            "precision_note": "headline = float32 messages on float32 rows (north_star's production type, parity at the same dtype); "
                              "the reference's own float64 arithmetic is the msa_f64 block (device rate, roofline, end-to-end with float64 rows)",
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, done, el, mean_it = oracle_rate(tables, SNR_DB, max(2048, 64 * threads), 10.0, threads)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d frames of the same workload in %.1f s, oracle/ldpc_oracle.c float32 "
                                              "min-sum on %d threads (mean %.2f iterations)" % (done, el, threads, mean_it)}
            r1, d1, e1, _ = oracle_rate(tables, SNR_DB, 512, 3.0, 1)
            line["cpu_baseline"]["single_thread_value"] = r1
            line["cpu_baseline"]["reference_python"] = reference_python_record()
        if world == 1 and not args.no_extras:
            try:
                line["extra"] = extra_workloads(torch, lib, eng_mod, Tables, peak)
            except Exception as exc:        # side measurements must not lose the headline line
                line["extra_error"] = repr(exc)
        print(json.dumps(line), file=RESULT, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
