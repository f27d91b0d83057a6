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
  roofline  the dominant kernel: algorithmic bytes / its event-timed launch durations vs measured HBM peak
            (on-chip path: an EFFECTIVE figure, plus `shared` = its shared-memory roofline and `traffic` = real DRAM bytes)
  roofline_streaming   the HBM-streaming path on the same workload (results asserted identical)
  spa       the other half of the metric: float32 sum-product on the same code / SNR / frames
  cpu_baseline / --impl reference: the oracle port (oracle/ldpc_oracle.c, scalar C restatement of src/bpa.py)
            on the host cores — the reference itself is Python and does not travel to the GPU box.
"""
import argparse
import json
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


def load_code(name=CODE):
    import _golden as G
    return G.code_tables(name)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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
    frames = max(2048, 128 * threads)
    from oracle import oracle as O
    m, n, rows, cols = tables
    g = O.Graph(m, n, rows, cols)
    rng = np.random.RandomState(7)
    Y = 1.0 + rng.normal(0, np.sqrt(10 ** (-SNR_DB / 10)), (frames, n))
    pri = O.llr_biawgn(SNR_DB, Y).astype(np.float32)
    its = 0
    for _ in range(args.warmup):
        O.bp_decode(g, O.MSA, pri, max_iter=MAX_ITER, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        its += int(O.bp_decode(g, O.MSA, pri, max_iter=MAX_ITER, nthreads=threads)["iters"].sum())
    el = time.perf_counter() - t0
    val = frames * args.steps / el
    sample = "%d frames/step x %d steps of the same workload (float32 min-sum, max_iter %d)" % (frames, args.steps, MAX_ITER)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(frames, world=1),
        "edge_updates_per_s": 2 * len(rows) * its / el,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
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
               "path": ("on-chip (%s)" % ("resident_vd" if dtype == lib.F64 else eng.resident_kernel)) if on_chip else "streaming"}
        # streaming: fraction of the measured HBM peak the whole step reaches; on-chip: the same algorithmic bytes never
        # touch HBM, so the figure is an EFFECTIVE one (see roofline.note)
        rec["effective_hbm_frac" if on_chip else "step_hbm_frac"] = frac
        out.append(rec)

    tab = Tables(*load_code())
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
                                % (nm, eng.resident_kernel or "streaming"),
                    "value": frames * 3 / (ms / 1e3), "unit": UNIT, "mean_iters": float(iters.mean()),
                    "edge_updates_per_s": 2 * irr.E * float(iters.sum()) * 3 / (ms / 1e3),
                    "path": ("on-chip (%s)" % eng.resident_kernel) if eng.resident_kernel else "streaming"})
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
    out.append({"workload": "LDPC(1200,3,6) BEC p=0.40 erasure decoding (bit planes), max_iter 10, cw=0",
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
    numa_cpus = ldist.bind_near_gpu(local_rank)       # pinned host buffers and the feeding thread on the GPU's NUMA node
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
    def measure(flags, algo=lib.MSA, y=y):
        """K timed steps of the device-resident hot path under `flags`; per-launch events recorded inside."""
        res = {}

        def step():
            res["o"] = eng.decode_device_channel(lib.CH_BIAWGN, algo, lib.F32, nv, y, max_iter=MAX_ITER,
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

    sampler = ClockSampler(local_rank) if rank == 0 and os.environ.get("LDPC_BENCH_NO_SAMPLER") != "1" else None
    main = measure(args.flags)
    resident = main["prof"]["vn_launches"] == 0          # the on-chip path is ONE kernel per decode
    ms, launches, prof, iters, x_hat = main["ms"], main["launches"], main["prof"], main["iters"], main["x_hat"]
    wer = float((x_hat != 1).any(dim=1).float().mean().item())

    total_frames = B * args.steps * world
    value = total_frames / (ms / 1e3)
    it_sum = float(iters.sum())
    if dist is not None:
        t = torch.tensor([it_sum], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        it_sum_all = float(t.item())
    else:
        it_sum_all = it_sum
    edge_updates = 2 * tab.E * it_sum_all * args.steps / (ms / 1e3)

    def streaming_roofline(m_):
        """Roofline of the dominant sweep kernel of a streaming run (this rank), from the per-launch events."""
        pr, it_ = m_["prof"], m_["iters"]
        cn_bytes, vn_bytes = algorithmic_bytes(tab.E, tab.n, 4, it_, MAX_ITER)
        cn_gbs = cn_bytes * args.steps / (pr["cn_ms"] / 1e3) / 1e9 if pr["cn_ms"] > 0 else 0.0
        vn_gbs = vn_bytes * args.steps / (pr["vn_ms"] / 1e3) / 1e9 if pr["vn_ms"] > 0 else 0.0
        dom = "vn_sweep" if pr["vn_ms"] >= pr["cn_ms"] else "cn_sweep"
        achieved = vn_gbs if dom == "vn_sweep" else cn_gbs
        return {
            "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_src, "traffic": traffic_per_launch(dom),
            "algorithmic_bytes_per_launch": ((vn_bytes if dom == "vn_sweep" else cn_bytes) * args.steps
                                             / max(1, pr["vn_launches"] if dom == "vn_sweep" else pr["cn_launches"])),
            "avg_launch_ms": (pr["vn_ms"] / max(1, pr["vn_launches"])) if dom == "vn_sweep" else (pr["cn_ms"] / max(1, pr["cn_launches"])),
            "cn_sweep": {"ms_total": pr["cn_ms"], "launches": pr["cn_launches"], "GBps": cn_gbs, "frac": cn_gbs / peak},
            "vn_sweep": {"ms_total": pr["vn_ms"], "launches": pr["vn_launches"], "GBps": vn_gbs, "frac": vn_gbs / peak},
            "sweeps_share_of_step": (pr["cn_ms"] + pr["vn_ms"]) / m_["ms"],
            "step_frac": (cn_bytes + vn_bytes) * args.steps / (m_["ms"] / 1e3) / 1e9 / peak,
            "bytes_per_edge_iteration": 17.5,
        }

    def resident_roofline(m_, kernel_note):
        """The on-chip kernel keeps every frame in shared memory for all iterations: the algorithmic bytes of the
        streaming layout (63 000 B per frame-iteration) never touch HBM, so "achieved" is an EFFECTIVE
        bandwidth and may exceed the HBM peak; `traffic` is what DRAM really sees, and `shared` is the kernel's own
        (shared-memory) roofline."""
        pr, it_ = m_["prof"], m_["iters"]
        cn_bytes, vn_bytes = algorithmic_bytes(tab.E, tab.n, 4, it_, MAX_ITER)
        k_ms = pr["cn_ms"]
        eff = (cn_bytes + vn_bytes) * args.steps / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0
        io_bytes = B * (tab.n * 4 + tab.n + 5)
        fi_per_s = float(it_.sum()) * args.steps / (k_ms / 1e3) if k_ms > 0 else 0.0
        # shared-memory roofline: per frame-iteration the formulation moves 3E + 2n floats through shared memory
        # (E marginal gathers + E message stores in the check phase, E message gathers + n prior loads + n marginal
        # stores in the variable phase); peak = 128 B/clk/SM (B300_MICROARCH.md, LDS/STS) x SMs x the SM clock.
        sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
        smem_bytes = (3 * tab.E + 2 * tab.n) * 4
        return {
            "bound": "hbm", "kernel": eng.resident_kernel, "achieved": eff, "peak": peak, "unit": "GB/s", "frac": eff / peak,
            "peak_source": peak_src, "traffic": traffic_per_launch(eng.resident_kernel),
            "note": "EFFECTIVE GB/s: algorithmic bytes of the streaming layout (SURVEY 8d, 63 000 B per frame-iteration) / "
                    "kernel time.  frac > 1 is by design, not skipped work: the kernel keeps every frame in shared memory "
                    "and registers for all its iterations, so DRAM only sees `traffic` (= compulsory_hbm_bytes_per_launch). "
                    + kernel_note + "  roofline_streaming is the HBM-bound path on the same workload, results asserted identical.",
            "algorithmic_bytes_per_launch": (cn_bytes + vn_bytes), "avg_launch_ms": k_ms / max(1, pr["cn_launches"]),
            "compulsory_hbm_bytes_per_launch": io_bytes,
            "compulsory_hbm_GBps": io_bytes * args.steps / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0,
            "kernel_share_of_step": k_ms / m_["ms"],
            "frame_iterations_per_s": fi_per_s,
            "shared": {"bound": "shared-memory pipe", "bytes_per_frame_iteration": smem_bytes,
                       "achieved_GBps": smem_bytes * fi_per_s / 1e9, "sm_count": sm_count,
                       "peak_bytes_per_clk_per_sm": 128},
        }

    streaming = None
    if resident:
        roofline = resident_roofline(main, "It is latency-limited between the shared-memory pipe (64 % of peak wavefronts), instruction issue (50 %) and the ALU pipe (44 %), profiles/README.md.")
        roofline["shared_memory_plan"] = eng.resident_plan()
        sm = measure(args.flags | lib.PATH_STREAMING)
        assert (sm["iters"] == iters).all() and bool((sm["x_hat"] == x_hat).all()), "streaming and resident paths disagree"
        streaming = {"value": total_frames / (sm["ms"] / 1e3), "unit": UNIT, "ms_per_step": sm["ms"] / args.steps,
                     "gpu_launches": int(sm["launches"]), "roofline": streaming_roofline(sm)}
    else:
        roofline = streaming_roofline(main)

    # ---- the other half of the metric: sum-product (float32), same code / SNR / frames, all-zero codeword
    # (src/simulations.py:36 runs SPA with --codeword 0)
    y_spa = y - 2.0
    sp = measure(args.flags, algo=lib.SPA, y=y_spa)
    sp_it = float(sp["iters"].sum())
    if dist is not None:
        t = torch.tensor([sp_it], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        sp_it_all = float(t.item())
    else:
        sp_it_all = sp_it
    spa = {"workload": "same code, BIAWGN %.1f dB, sum-product float32 (hyperbolic-pair rule), max_iter %d, codeword=0" % (SNR_DB, MAX_ITER),
           "value": total_frames / (sp["ms"] / 1e3), "unit": UNIT, "ms_per_step": sp["ms"] / args.steps,
           "mean_iters": sp_it / B, "edge_updates_per_s": 2 * tab.E * sp_it_all * args.steps / (sp["ms"] / 1e3),
           "wer": float((sp["x_hat"] != 0).any(dim=1).float().mean().item()), "gpu_launches": int(sp["launches"]),
           "roofline": (resident_roofline(sp, "Sum-product adds the MUFU pipe (18 ex2/lg2 per check and frame, 45 %) to the limiters.")
                        if sp["prof"]["vn_launches"] == 0 else streaming_roofline(sp))}
    del y_spa

    # ---- e2e: host buffers through the host entry point (what decode_batch calls), copies inside the timed region
    Yh = pinned_empty((B, tab.n), np.float32)
    Yh[...] = y.cpu().numpy()
    xh, ith, rsh = pinned_empty((B, tab.n), np.uint8), pinned_empty((B,), np.int32), pinned_empty((B,), np.uint8)

    def e2e_step(wait=True):
        eng.decode_host(lib.CH_BIAWGN, lib.MSA, lib.F32, nv, Yh, max_iter=MAX_ITER, x_hat=xh, iters=ith, reason=rsh,
                        flags=args.flags, wait=wait)

    def e2e_run(stream_of_batches):
        """K steps, each with its own H2D of the received block and D2H of the words inside the timed region.
        stream_of_batches: the steps are submitted back to back (decode_host(wait=False)) and completed by one
        host_sync, the way a caller with many batches uses the entry point; otherwise every call blocks."""
        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        l0 = eng.launch_count
        t0_ = time.perf_counter()
        for _ in range(args.steps):
            e2e_step(wait=not stream_of_batches)
        eng.host_sync()
        torch.cuda.synchronize()
        return time.perf_counter() - t0_, eng.launch_count - l0

    blk_el, _ = e2e_run(False)
    if dist is not None:
        t = torch.tensor([blk_el], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        blk_el = float(t.item())
    e_el, e_launches = e2e_run(True)
    e_rank_ms = [1e3 * e_el / args.steps]
    if dist is not None:
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = e_el
        dist.all_reduce(t)
        e_rank_ms = [1e3 * float(v) / args.steps for v in t.tolist()]
        e_el = float(t.max().item())
    assert (ith == iters).all() and bool((torch.from_numpy(xh).cuda() == x_hat).all()), "e2e result differs from device path"
    e2e = {"value": total_frames / e_el, "unit": UNIT, "h2d_bytes_per_step": int(Yh.nbytes),
           "d2h_bytes_per_step": int(xh.nbytes + ith.nbytes + rsh.nbytes), "ms_per_step": 1e3 * e_el / args.steps,
           "api": "Engine.decode_host(wait=False) x K + host_sync -> ldpc_decode_host with LDPC_HOST_ASYNC (pinned float32 y in; "
                  "x_hat, iters, reason out; every step copies its own input and output)",
           "blocking_call_value": total_frames / blk_el,
           "timer": "host perf_counter around blocking calls, max over ranks",
           "ms_per_step_by_rank": e_rank_ms, "host_cpus_bound": (len(numa_cpus) if numa_cpus else None)}

    clocks = sampler.stop() if sampler is not None else None
    for rf in (roofline, spa["roofline"]):
        sh = rf.get("shared") if isinstance(rf, dict) else None
        if sh is not None:
            mhz = (clocks or {}).get("sm_mhz") or 1965.0
            sh["peak_GBps"] = sh["peak_bytes_per_clk_per_sm"] * sh["sm_count"] * mhz * 1e6 / 1e9
            sh["frac"] = sh["achieved_GBps"] / sh["peak_GBps"]
            sh["sm_mhz"] = mhz

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, world),
            "edge_updates_per_s": edge_updates, "mean_iters": it_sum / B, "wer": wer,
            "path": "resident (on-chip, LDPC_PATH_AUTO)" if resident else "streaming",
            "roofline": roofline, "roofline_streaming": streaming, "spa": spa, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(launches), "gpu_launches_e2e": int(e_launches),
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, done, el, mean_it = oracle_rate(tables, SNR_DB, max(2048, 64 * threads), 10.0, threads)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d frames of the same workload in %.1f s, oracle/ldpc_oracle.c float32 "
                                              "min-sum on %d threads (mean %.2f iterations)" % (done, el, threads, mean_it)}
            r1, d1, e1, _ = oracle_rate(tables, SNR_DB, 512, 3.0, 1)
            line["cpu_baseline"]["single_thread_value"] = r1
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
