"""CPU tier: host-side logic of the batched driver and the multi-process path (gloo, world_size 2).

The decode function injected here is the oracle (tests may use it as a stand-in checker); the product's
decoders need a GPU and are covered by tests/test_gpu_parity.py / test_gpu_sim.py.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import _golden as G
from ldpc_decoders_b200 import dist, sim
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_round_slices_tile_the_frame_axis():
    seen = []
    for rnd in range(3):
        for r in range(4):
            g0, g1 = dist.round_slice(rnd, r, 4, 5)
            seen.extend(range(g0, g1))
    assert seen == list(range(60))


def test_sequential_stop_matches_frame_by_frame_loop():
    rng = np.random.RandomState(0)
    for _ in range(200):
        errs = rng.binomial(3, .2, size=rng.randint(1, 40))
        wec0, min_wec = rng.randint(0, 5), rng.randint(1, 8)
        wec, used = wec0, 0
        for e in errs:                      # the reference's loop, src/main.py:37-45
            if wec >= min_wec:
                break
            wec += e > 0
            used += 1
        assert dist.sequential_stop(errs, wec0, min_wec) == used


def oracle_decoder(name, snr):
    g = O.Graph(*G.code_tables(name))
    def decode_batch(Y):
        r = O.bp_decode(g, O.MSA, O.llr_biawgn(snr, Y), max_iter=10, nthreads=2)
        return r["x_hat"], r["iters"]
    return g, decode_batch


def test_run_param_reproduces_the_sequential_reference_loop():
    """Batched + stopping rule == decoding the same RNG stream one frame at a time (what main.test does)."""
    name, snr, min_wec = "512_3_6_rand_ldpc_1", 1.5, 7
    g, decode_batch = oracle_decoder(name, snr)
    x = np.ones(g.n, np.int64)
    std = np.sqrt(10 ** (-snr / 10))
    send = lambda X: (2 * X - 1) + np.random.normal(0, std, X.shape)
    np.random.seed(3)
    tot = wec = bec = 0
    while wec < min_wec:                                  # src/main.py:37-45
        y = send(x)
        x_hat, _ = decode_batch(y[None, :])
        e = int((x_hat[0] != x).sum())
        wec += e > 0; bec += e; tot += 1
    for batch in (1, 4, 16):
        np.random.seed(3)
        r = sim.run_param(decode_batch, send, x, dist.Comm(), batch, min_wec)
        assert (r["tot"], r["wec"], r["bec"]) == (tot, wec, bec), batch
        assert sum(r["dec"]["iter"]) == tot


def test_saver_schema_matches_reference(tmp_path):
    ids = [("channel", "biawgn"), ("code", "c"), ("decoder", "MSA"), ("codeword", 1), ("min_wec", 100), ("max_iter", 10)]
    s = sim.Saver(str(tmp_path), ids)
    s.add(2.0, dict(tot=10, wec=1, wer=.1, bec=3, ber=.001))
    s.add(2.5, dict(tot=20, wec=1, wer=.05, bec=2, ber=.0005))
    s.add(2.0, dict(tot=11, wec=2, wer=.2, bec=4, ber=.002))
    assert os.path.basename(s.file_path) == "biawgn-c-MSA-1-100-10.json"
    d = json.load(open(s.file_path))
    assert list(d)[:6] == [k for k, _ in ids]
    assert d["tot"] == {"2.0": 11, "2.5": 20} and d["wer"]["2.5"] == .05


WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import _golden as G
from ldpc_decoders_b200 import dist, sim
from oracle import oracle as O
comm = dist.Comm("gloo")
g = O.Graph(*G.code_tables("512_3_6_rand_ldpc_1"))
snr = 1.5
def decode_batch(Y):
    r = O.bp_decode(g, O.MSA, O.llr_biawgn(snr, Y), max_iter=10)
    return r["x_hat"], r["iters"]
x = np.ones(g.n, np.int64)
std = np.sqrt(10 ** (-snr / 10))
send = lambda X: (2 * X - 1) + np.random.normal(0, std, X.shape)
np.random.seed(3)
r = sim.run_param(decode_batch, send, x, comm, 4, 7)
tot = comm.allreduce_sum(np.array([r["tot"], comm.rank]))
if comm.rank == 0:
    print("RESULT " + json.dumps(dict(tot=r["tot"], wec=r["wec"], bec=r["bec"], sum_tot=int(tot[0]), ranks=int(tot[1]))))
comm.close()
'''


def test_two_process_gloo_run_matches_single_process(tmp_path):
    name, snr, min_wec = "512_3_6_rand_ldpc_1", 1.5, 7
    g, decode_batch = oracle_decoder(name, snr)
    x = np.ones(g.n, np.int64)
    std = np.sqrt(10 ** (-snr / 10))
    send = lambda X: (2 * X - 1) + np.random.normal(0, std, X.shape)
    np.random.seed(3)
    ref = sim.run_param(decode_batch, send, x, dist.Comm(), 8, min_wec)

    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][0]
    r = json.loads(line[7:])
    assert (r["tot"], r["wec"], r["bec"]) == (ref["tot"], ref["wec"], ref["bec"])
    assert r["sum_tot"] == 2 * ref["tot"] and r["ranks"] == 1      # all_reduce(SUM) over 2 ranks


# ---- fixed-length runs with the counters kept "on the device" (run_fixed_on_device): one all-reduce per parameter.
# The stand-in round below draws frame-index-keyed noise (like the Philox generator) and decodes with the oracle; its
# counters live in a CPU torch tensor, which is where a gloo group wants them.
FIXED_HELPER = r"""
import numpy as np, torch
def make_round(O, g, snr, calls):
    std = np.sqrt(10 ** (-snr / 10))
    def simulate_round(x, nb, seed, frame0, c, nh):
        Y = np.stack([(2 * x - 1) + np.random.RandomState((seed + frame0 + i) % (2 ** 31)).normal(0, std, x.size) for i in range(nb)])
        r = O.bp_decode(g, O.MSA, O.llr_biawgn(snr, Y), max_iter=10)
        e = (r["x_hat"] != x[None, :]).sum(1)
        c[0] += nb; c[1] += int((e > 0).sum()); c[2] += int(e.sum()); c[3] += int(r["iters"].sum())
        for it in r["iters"]:
            c[4 + min(int(it), nh - 1)] += 1
        calls.append((frame0, nb))
    return simulate_round
new_counters = lambda k: torch.zeros(4 + k, dtype=torch.int64)
"""


def _fixed_ref(frames):
    ns = {}
    exec(FIXED_HELPER, ns)
    g = O.Graph(*G.code_tables("512_3_6_rand_ldpc_1"))
    x = np.ones(g.n, np.int64)
    calls = []
    r = sim.run_fixed_on_device(ns["make_round"](O, g, 1.5, calls), ns["new_counters"], x, dist.Comm(), 8, frames, seed=11)
    return r, calls


def test_fixed_length_device_counters_single_process():
    r, calls = _fixed_ref(37)
    assert r["tot"] == 37 and sum(r["dec"]["iter"]) == 37 and r["wec"] <= 37
    assert calls == [(0, 8), (8, 8), (16, 8), (24, 8), (32, 5)]          # the last round is short, nothing past `frames`
    r2, calls2 = _fixed_ref(37)
    assert r2 == r
    # status callbacks (partial results) do not change the outcome
    ns = {}
    exec(FIXED_HELPER, ns)
    g = O.Graph(*G.code_tables("512_3_6_rand_ldpc_1"))
    seen = []
    r3 = sim.run_fixed_on_device(ns["make_round"](O, g, 1.5, []), ns["new_counters"], np.ones(g.n, np.int64), dist.Comm(),
                                 8, 37, on_status=lambda *a: seen.append(a[0]), log_freq=0., seed=11)
    assert r3 == r and seen and seen == sorted(seen) and seen[-1] < 37


FIXED_WORKER = r"""
import os, sys, json
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import _golden as G
from ldpc_decoders_b200 import dist, sim
from oracle import oracle as O
{helper}
comm = dist.Comm("gloo")
g = O.Graph(*G.code_tables("512_3_6_rand_ldpc_1"))
x = np.ones(g.n, np.int64)
calls, seen = [], []
r = sim.run_fixed_on_device(make_round(O, g, 1.5, calls), new_counters, x, comm, 4, 37,
                            on_status=lambda *a: seen.append(a[0]), log_freq=0., seed=11)
if comm.rank == 0:
    print("RESULT " + json.dumps(dict(r=r, calls=calls, seen=seen)))
comm.close()
"""


def test_fixed_length_device_counters_two_process_gloo(tmp_path):
    ref, _ = _fixed_ref(37)
    script = tmp_path / "worker_fixed.py"
    script.write_text(FIXED_WORKER.format(root=ROOT, helper=FIXED_HELPER))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         env=dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533"),
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    got = json.loads([l for l in out.stdout.splitlines() if l.startswith("RESULT ")][0][7:])
    assert got["r"] == json.loads(json.dumps(ref))                       # same frames, same counters, any GPU count
    assert [tuple(c) for c in got["calls"]] == [(0, 4), (8, 4), (16, 4), (24, 4), (32, 4)]      # rank 0's slices
    assert got["seen"] == sorted(got["seen"]) and all(0 < t < 37 for t in got["seen"])


def test_run_param_fixed_frames_host_noise_sums_locally():
    """--frames with numpy noise: local sums + ONE all-reduce; equals the min_wec-free sequential loop."""
    name, snr = "512_3_6_rand_ldpc_1", 1.5
    g, decode_batch = oracle_decoder(name, snr)
    x = np.ones(g.n, np.int64)
    std = np.sqrt(10 ** (-snr / 10))
    send = lambda X: (2 * X - 1) + np.random.normal(0, std, X.shape)
    np.random.seed(5)
    Y = send(np.tile(x, (21, 1)))
    xh, it = decode_batch(Y)
    e = (xh != x[None, :]).sum(1)
    np.random.seed(5)
    r = sim.run_param(decode_batch, send, x, dist.Comm(), 7, 100, frames=21)
    assert (r["tot"], r["wec"], r["bec"]) == (21, int((e > 0).sum()), int(e.sum()))
    assert r["dec"]["average"] == it.sum() / 21


def test_status_persists_partial_results(tmp_path):
    """sim.main saves the counters at every status report (main.py:30-35 log_status -> saver.add)."""
    ids = [("channel", "biawgn"), ("code", "c"), ("decoder", "MSA"), ("codeword", 1), ("min_wec", 100), ("max_iter", 10)]
    s = sim.Saver(str(tmp_path), ids)
    r = sim._result(10, 2, 5, 93, np.array([0, 0, 1, 9]), 1200)
    s.add(2.0, {k: r[k] for k in ("tot", "wec", "wer", "bec", "ber", "dec")})
    d = json.load(open(s.file_path))
    assert d["tot"]["2.0"] == 10 and d["dec"]["2.0"]["iter"] == [0, 0, 1, 9] and abs(d["dec"]["2.0"]["average"] - 9.3) < 1e-12


def test_named_simulation_cases_match_the_reference():
    """ldpc_decoders_b200.simulations expands REG_ENS / IREG_ENS / REG_BAD / MAR / HMG to the same main.py command
    lines as the reference's simulations.py (tests/golden/simulations_cases.txt = its own output, SPA / MSA lines)."""
    import os
    from ldpc_decoders_b200 import simulations, sim

    def norm(tokens):
        ap = sim.setup_parser()
        a = ap.parse_args(tokens)
        return (a.channel, a.code, a.decoder, a.codeword, a.max_iter, a.min_wec, tuple(round(v, 9) for v in a.params))

    here = os.path.dirname(os.path.abspath(__file__))
    gold = [norm(line.split()) for line in open(os.path.join(here, "golden", "simulations_cases.txt")) if line.strip()]
    ours = [norm(c) for name in ("REG_ENS", "IREG_ENS", "REG_BAD", "MAR", "HMG") for c in simulations.all_cases[name]()]
    assert len(gold) == 150 and ours == gold


def test_parse_cpulist_and_bind_is_harmless_without_gpu():
    from ldpc_decoders_b200 import dist as D
    assert D.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert D.parse_cpulist("") == set()
    assert D.bind_near_gpu(0) is None or isinstance(D.bind_near_gpu(0), set)     # no CUDA device here: a no-op
    assert D.bind_memory_near_gpu(0) is None


def test_random_irregular_sampler_follows_the_reference_construction():
    """codes.random_irregular (src/ldpc.py:149-192 without the dense matrix): seeded, the requested variable degrees up
    to the pairwise cancellation of parallel edges, checks of degree dc or dc - 2k, and the result decodes."""
    import _golden as G
    from ldpc_decoders_b200 import Tables, codes
    from oracle import oracle as O
    ref = Tables(*G.code_tables("1200_rho_x5_rand_ldpc_1"))
    want = codes.variable_degree_counts(ref)
    # the shipped file lost a few edges to cancellation: top the profile up to a multiple of dc like the reference's `extra`
    prof = dict(want)
    short = (-sum(d * c for d, c in prof.items())) % 6
    if short:
        prof[2] -= 1
        prof[2 + short] = prof.get(2 + short, 0) + 1
    a = codes.random_irregular(prof, 6, seed=3).tables
    b = codes.random_irregular(prof, 6, seed=3).tables
    c = codes.random_irregular(prof, 6, seed=4).tables
    assert (a.edge_chk == b.edge_chk).all() and (a.edge_var == b.edge_var).all()
    assert a.E != c.E or (a.edge_var != c.edge_var).any()
    assert a.n == 1200 and a.m == sum(d * k for d, k in prof.items()) // 6
    dc = a.check_degrees
    assert set(dc.tolist()) <= {6, 4, 2} and (dc == 6).mean() > 0.95
    lost = sum(d * k for d, k in prof.items()) - a.E
    assert 0 <= lost <= 60 and lost % 2 == 0
    got = codes.variable_degree_counts(a)
    assert sum(abs(got.get(d, 0) - prof.get(d, 0)) for d in set(got) | set(prof)) <= 2 * lost
    # decodes: clean BSC words stay, a few flips are corrected
    g = O.Graph(a.m, a.n, a.edge_chk.astype(np.int64), a.edge_var.astype(np.int64))
    y = np.zeros((8, a.n), np.uint8)
    y[1:, ::151] = 1
    out = O.bp_decode(g, O.MSA, O.llr_bsc(.02, y), y_hard=y, max_iter=20)
    assert out["iters"][0] == 0 and (out["x_hat"] == 0).all()
