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


def run_param(decode_batch, send, x, comm, batch, min_wec, frames=0, max_iter=10, on_status=None, log_freq=5.,
              simulate_batch=None, seed=0):
    """Monte-Carlo loop for one channel parameter.

    decode_batch(Y) -> (X_hat [b,n], iters [b]);  send(X) -> received block for X [b,n] (global RNG stream).
    Every rank draws the whole round (world*batch frames, so the stream equals the reference's sequential draws,
    SURVEY H8) and decodes its own slice; per-frame results are all-gathered and consumed in global order.
    """
    n = x.size
    tot = wec = bec = it_sum = 0
    hist = np.zeros(max(max_iter, 0) + 2, np.int64)
    rnd = 0
    start = time.time()
    while (wec < min_wec) if frames <= 0 else (tot < frames):
        g0, g1 = round_slice(rnd, comm.rank, comm.world, batch)
        if simulate_batch is not None:                   # noise drawn on the GPU, keyed by the global frame index
            errs, iters = simulate_batch(x, batch, seed, g0)
            errs = np.asarray(errs, np.int64)
        else:
            Y = send(np.tile(x, (comm.world * batch, 1)))
            lo = g0 - rnd * comm.world * batch
            X_hat, iters = decode_batch(Y[lo:lo + batch])
            errs = (np.asarray(X_hat) != x[None, :]).sum(axis=1).astype(np.int64)
        both = comm.allgather(np.concatenate([errs, np.asarray(iters, np.int64)]))
        errs_g = both[:, :batch].reshape(-1)
        iters_g = both[:, batch:].reshape(-1)
        if frames > 0:
            take = min(errs_g.size, frames - tot)
        else:
            take = sequential_stop(errs_g, wec, min_wec)
        e, it = errs_g[:take], iters_g[:take]
        tot += take
        wec += int((e > 0).sum())
        bec += int(e.sum())
        it_sum += int(it.sum())
        np.add.at(hist, np.minimum(it, hist.size - 1), 1)
        rnd += 1
        if on_status is not None and time.time() - start > log_freq:
            start = time.time()
            on_status(tot, wec, bec, it_sum, hist)
    return dict(tot=tot, wec=wec, wer=wec / max(tot, 1), bec=bec, ber=bec / max(tot * n, 1),
                dec={'average': it_sum / max(tot, 1), 'iter': hist[:int(np.flatnonzero(hist)[-1]) + 1].tolist() if hist.any() else []})


def main(argv=None):
    args = setup_parser().parse_args(argv)
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
    for param in args.params:
        log.info('Starting parameter: %f' % param)
        if args.seed is not None:
            np.random.seed(args.seed)
        channel = model.Channel(param)
        decoder = dec_fac(param, code, **dict(vars(args), dtype=dt))

        def status(tot, wec, bec, it_sum, hist, final=False):
            if comm.rank != 0:
                return
            wer, ber = wec / max(tot, 1), bec / max(tot * x.size, 1)
            log.info('TOT:%d, WEC:%d, WER:%s, BEC:%d, BER:%s' % (tot, wec, wer, bec, ber))

        r = run_param(decoder.decode_batch, channel.send, x, comm, args.batch, args.min_wec, args.frames,
                      args.max_iter, status, args.log_freq,
                      simulate_batch=decoder.simulate_batch if args.noise == 'device' else None,
                      seed=(args.seed or 0) * 1000003 + int(round(param * 1e6)))
        if comm.rank == 0:
            log.info('TOT:%d, WEC:%d, WER:%s, BEC:%d, BER:%s' % (r['tot'], r['wec'], r['wer'], r['bec'], r['ber']))
            saver.add(param, OrderedDict((k, r[k]) for k in ('tot', 'wec', 'wer', 'bec', 'ber', 'dec')))
    log.info('Done!')
    comm.close()


if __name__ == '__main__':
    main()
