"""Batched Monte-Carlo driver — the GPU sibling of the reference's main.test (src/main.py:10-51).

Same CLI vocabulary (src/utils.py:21-55) plus --batch / --dtype / --frames / --seed / --noise, same counters
(tot, wec, wer, bec, ber), same stopping rule (`while wec < min_wec`, applied to the global frame order so the
result does not depend on batch size or GPU count), same result-file schema as utils.Saver (src/utils.py:118-140)
so src/graph.py / plot_results.py read our outputs, and `dec` = iteration histogram in ADMM's stats() shape
(src/admm.py:36-40, hook at main.py:34).

    python -m ldpc_decoders_b200.sim biawgn 1200_3_6_rand_ldpc_1 MSA --codeword 1 --params 2.0 2.5 --batch 8192
    torchrun --nproc-per-node 8 -m ldpc_decoders_b200.sim ...      # frames sharded over GPUs, counters all-reduced
"""
import argparse
import json
import logging
import os
import time
from collections import OrderedDict

import numpy as np

from .dist import Comm, round_slice, sequential_stop

decoder_names = ['SPA', 'MSA']


def setup_parser():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument('channel', choices=['bsc', 'bec', 'biawgn'])
    p.add_argument('code')
    p.add_argument('decoder', choices=decoder_names)
    p.add_argument('--codeword', default=0, type=int, choices=[0, 1])
    p.add_argument('--min-wec', default=100, type=int)
    p.add_argument('--params', nargs='+', type=float, default=[.1, .01])
    p.add_argument('--max-iter', default=10, type=int)
    p.add_argument('--log-freq', default=5., type=float)
    p.add_argument('--data_dir', default=os.path.abspath(os.path.join('data', 'output_b200')))
    p.add_argument('--console', action='store_true')
    p.add_argument('--debug', action='store_true')
    # new
    p.add_argument('--batch', default=8192, type=int, help='frames per GPU per round')
    p.add_argument('--dtype', default='f64', choices=['f32', 'f64'], help='message arithmetic (f64 = the reference)')
    p.add_argument('--frames', default=0, type=int, help='fixed number of frames per parameter instead of --min-wec')
    p.add_argument('--seed', default=None, type=int, help='np.random.seed before each parameter (reference: unseeded)')
    p.add_argument('--noise', default='host', choices=['host', 'device'],
                   help="host: numpy's global RNG like the reference (parity runs); device: Philox on the GPU, keyed by "
                        "(seed, global frame index) - same frames for any --batch / GPU count, nothing but counters crosses PCIe")
    p.add_argument('--codes-dir', default=None)
    return p


class Saver:
    """Result file with the reference's schema (src/utils.py:118-140): run ids first, then one dict per metric
    keyed by str(param); re-runs merge per parameter."""

    def __init__(self, data_dir, run_ids):
        self.ids = OrderedDict(run_ids)
        os.makedirs(data_dir, exist_ok=True)
        self.file_path = os.path.join(data_dir, '%s.json' % '-'.join(str(v) for v in self.ids.values()))

    def add(self, param, val_dict):
        data = None
        try:
            with open(self.file_path) as fp:
                data = json.load(fp, object_pairs_hook=OrderedDict)
        except Exception:
            pass
        if data is None:
            data = OrderedDict(self.ids)
        for key, val in val_dict.items():
            data.setdefault(key, OrderedDict())[str(param)] = val
        with open(self.file_path, 'w') as fp:
            json.dump(data, fp, indent=4)


def _result(tot, wec, bec, it_sum, hist, n):
    hist = np.asarray(hist, np.int64)
    return dict(tot=int(tot), wec=int(wec), wer=wec / max(tot, 1), bec=int(bec), ber=bec / max(tot * n, 1),
                dec={'average': it_sum / max(tot, 1),
                     'iter': hist[:int(np.flatnonzero(hist)[-1]) + 1].tolist() if hist.any() else []})


def hist_bins(max_iter, iter_cap=1000):
    """Bins of the iteration histogram: 0 .. max_iter (+1 overflow); 'unlimited' (max_iter <= 0) runs up to iter_cap."""
    return (max_iter if max_iter > 0 else iter_cap) + 2


def run_fixed_on_device(simulate_round, new_counters, x, comm, batch, frames, max_iter=10, on_status=None,
                        log_freq=5., seed=0):
    """Fixed number of frames per parameter, noise drawn on the GPU: every rank runs its rounds back to back with the
    counters [tot, wec, bec, sum iters, histogram] accumulated ON THE DEVICE and never looks at them; ONE all-reduce of
    that vector (NCCL over NVLink) ends the parameter.  Nothing else is exchanged and nothing is copied to
    the host inside the loop.

    Progress reports / partial results (main.py:46-48 logs and saves every log_freq seconds): all ranks must enter a
    collective at the same round, so the reporting interval is counted in ROUNDS and adapted from the time rank 0
    measured, which travels inside the reduced vector itself (last element) - every rank derives the same next interval.

    simulate_round(x, nb, seed, frame0, counters, nhist) adds one round to `counters`; new_counters(k) gives a zeroed
    int64 tensor of 4 + k elements on the device of the process group.
    """
    nh = hist_bins(max_iter)
    c = new_counters(nh + 1)                               # [... | ms rank 0 spent since the last report]
    per_round = comm.world * batch
    rounds = (frames + per_round - 1) // per_round
    every, nxt = 1, 1
    t_last = time.time()
    for rnd in range(rounds):
        g0, g1 = round_slice(rnd, comm.rank, comm.world, batch)
        nb = min(g1, frames) - g0                          # the last round may be short (or empty) for this rank
        if nb > 0:
            simulate_round(x, nb, seed, g0, c, nh)
        if rnd + 1 == nxt and rnd + 1 < rounds and on_status is not None:
            g = c.clone()
            g[-1] = int(1e3 * (time.time() - t_last)) if comm.rank == 0 else 0
            tot = comm.allreduce_tensor(g).cpu().numpy()   # the only synchronisation inside the loop, ~ every log_freq / 2 s
            ms = int(tot[-1])
            if ms < 500. * log_freq:
                every = min(every * 2, 1 << 20)
            else:
                on_status(int(tot[0]), int(tot[1]), int(tot[2]), int(tot[3]), tot[4:4 + nh])
                t_last = time.time()
            nxt = rnd + 1 + every
    c[-1] = 0
    tot = comm.allreduce_tensor(c).cpu().numpy()
    return _result(tot[0], tot[1], tot[2], tot[3], tot[4:4 + nh], x.size)


def run_param(decode_batch, send, x, comm, batch, min_wec, frames=0, max_iter=10, on_status=None, log_freq=5.,
              simulate_batch=None, seed=0):
    """Monte-Carlo loop for one channel parameter.

    decode_batch(Y) -> (X_hat [b,n], iters [b]);  send(X) -> received block for X [b,n] (global RNG stream).
    Every rank draws the whole round (world*batch frames, so the stream equals the reference's sequential draws,
    SURVEY H8) and decodes its own slice.
      * `while wec < min_wec` (the reference's rule, main.py:37): the per-frame (bit errors, iterations) of a round are
        all-gathered - on the device when the decoder returns CUDA tensors, one read-back per round - and consumed in
        GLOBAL frame order, so the counters equal a sequential run whatever the batch size or GPU count.
      * fixed `frames`: counters are summed locally and all-reduced ONCE at the end.
    """
    n = x.size
    nh = hist_bins(max_iter)
    tot = wec = bec = it_sum = 0
    hist = np.zeros(nh, np.int64)
    rnd = 0
    start = time.time()

    def consume(e, it):
        nonlocal tot, wec, bec, it_sum
        tot += int(e.size)
        wec += int((e > 0).sum())
        bec += int(e.sum())
        it_sum += int(it.sum())
        np.add.at(hist, np.minimum(it, nh - 1), 1)

    while (wec < min_wec) if frames <= 0 else (rnd * comm.world * batch < frames):
        g0, g1 = round_slice(rnd, comm.rank, comm.world, batch)
        on_dev = False
        if simulate_batch is not None:                   # noise drawn on the GPU, keyed by the global frame index
            try:
                errs, iters = simulate_batch(x, batch, seed, g0, on_device=True)
                on_dev = hasattr(errs, "is_cuda")
            except TypeError:                            # a stand-in without the keyword (tests)
                errs, iters = simulate_batch(x, batch, seed, g0)
        else:
            Y = send(np.tile(x, (comm.world * batch, 1)))
            lo = g0 - rnd * comm.world * batch
            X_hat, iters = decode_batch(Y[lo:lo + batch])
            errs = (np.asarray(X_hat) != x[None, :]).sum(axis=1)
        if frames > 0:                                   # fixed length: local sums, one all-reduce at the end
            if on_dev:
                errs, iters = errs.cpu().numpy(), iters.cpu().numpy()
            keep = max(0, min(g1, frames) - g0)
            consume(np.asarray(errs, np.int64)[:keep], np.asarray(iters, np.int64)[:keep])
        else:
            if on_dev:
                import torch
                both = comm.allgather_tensor(torch.cat([errs, iters]).to(torch.int64)).cpu().numpy()
            else:
                both = comm.allgather(np.concatenate([np.asarray(errs, np.int64), np.asarray(iters, np.int64)]))
            errs_g = both[:, :batch].reshape(-1)
            iters_g = both[:, batch:].reshape(-1)
            take = sequential_stop(errs_g, wec, min_wec)
            consume(errs_g[:take], iters_g[:take])
        rnd += 1
        if on_status is not None and frames <= 0 and time.time() - start > log_freq:
            start = time.time()
            on_status(tot, wec, bec, it_sum, hist)       # min_wec mode: every rank holds the global counters
    if frames > 0:
        v = comm.allreduce_sum(np.concatenate([[tot, wec, bec, it_sum], hist]))
        tot, wec, bec, it_sum, hist = int(v[0]), int(v[1]), int(v[2]), int(v[3]), v[4:]
    return _result(tot, wec, bec, it_sum, hist, n)


def main(argv=None, comm=None):
    """One simulation case.  `comm`: a process group that outlives this call (simulations.main runs many cases on one)."""
    args = setup_parser().parse_args(argv)
    own_comm = comm is None
    if own_comm:
        comm = Comm()
    import torch
    if torch.cuda.is_available():
        torch.cuda.set_device(comm.local_rank)
    from . import codes
    from .models import models
    level = logging.DEBUG if args.debug else logging.INFO
    if args.console:
        logging.basicConfig(format='%(name)s|%(message)s', level=level)
    else:
        os.makedirs(args.data_dir, exist_ok=True)
        logging.basicConfig(filename=os.path.join(args.data_dir, 'test.log'), filemode='a', level=level,
                            format='%(asctime)s,%(msecs)03d|%(name)s|%(levelname)s|%(message)s', datefmt='%H:%M:%S')
    model = models[args.channel]
    dec_fac = getattr(model, args.decoder)
    id_keys = ['channel', 'code', 'decoder', 'codeword', 'min_wec'] + dec_fac.id_keys          # main.py:13
    id_val = [vars(args)[k] for k in id_keys]
    log = logging.getLogger('.'.join(str(v) for v in id_val))
    code = codes.get_code(args.code, args.codes_dir)
    x = np.zeros(code.get_n(), np.int64) + args.codeword                                       # main.py:18
    saver = Saver(args.data_dir, list(zip(id_keys, id_val))) if comm.rank == 0 else None
    dt = np.float32 if args.dtype == 'f32' else np.float64
    results = OrderedDict()
    for param in args.params:
        log.info('Starting parameter: %f' % param)
        if args.seed is not None:
            np.random.seed(args.seed)
        channel = model.Channel(param)
        decoder = dec_fac(param, code, **dict(vars(args), dtype=dt))

        def status(tot, wec, bec, it_sum, hist):
            """main.py:30-35 log_status: log the counters AND save them, so a killed run keeps its partial result."""
            if comm.rank != 0:
                return
            r = _result(tot, wec, bec, it_sum, hist, x.size)
            log.info('TOT:%d, WEC:%d, WER:%s, BEC:%d, BER:%s' % (r['tot'], r['wec'], r['wer'], r['bec'], r['ber']))
            saver.add(param, OrderedDict((k, r[k]) for k in ('tot', 'wec', 'wer', 'bec', 'ber', 'dec')))

        seed = (args.seed or 0) * 1000003 + int(round(param * 1e6))
        if args.noise == 'device' and args.frames > 0:
            eng = getattr(decoder, 'dec', decoder).engine
            r = run_fixed_on_device(decoder.simulate_round, eng.new_counters, x, comm, args.batch, args.frames,
                                    args.max_iter, status, args.log_freq, seed=seed)
        else:
            r = run_param(decoder.decode_batch, channel.send, x, comm, args.batch, args.min_wec, args.frames,
                          args.max_iter, status, args.log_freq,
                          simulate_batch=decoder.simulate_batch if args.noise == 'device' else None, seed=seed)
        results[str(param)] = r
        status(r['tot'], r['wec'], r['bec'], r['dec']['average'] * r['tot'], _pad_hist(r['dec']['iter'], args.max_iter))
    log.info('Done!')
    if own_comm:
        comm.close()
    return results


def _pad_hist(hist, max_iter):
    out = np.zeros(max(hist_bins(max_iter), len(hist)), np.int64)
    out[:len(hist)] = hist
    return out


if __name__ == '__main__':
    main()
