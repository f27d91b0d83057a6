"""Named simulation cases — the GPU sibling of the reference's simulations.py + run_sims.sh.

The reference prints one `src/main.py` command line per (channel, code, decoder) case and run_sims.sh starts one
OS process for each (simulations.py:24-85, run_sims.sh:15-27).  Here the same named cases (restricted to the
message-passing decoders this package replaces: SPA, MSA) run IN-PROCESS through sim.main on one engine cache, so
a code's tables are uploaded once and every case of an ensemble reuses them.

    python -m ldpc_decoders_b200.simulations REG_ENS --print                    # the reference's lines, verbatim
    python -m ldpc_decoders_b200.simulations REG_ENS IREG_ENS --noise device --dtype f32 --data_dir out/
    torchrun --nproc-per-node 8 -m ldpc_decoders_b200.simulations REG_ENS --noise device
"""
import argparse

from . import sim

P_BEC = '.5 .475 .45 .425 .4 .375 .35 .34 .33 .325 .32 .31 .3'
P_BSC_MSA = '.081 .0751 .071 .0651 .061 .0551 .051 .0451 .041 .0351 .031 .0251 .021 .0151 .01'
P_AWGN_MSA = '.5 .75 1. 1.25 1.5 1.75 2. 2.2 2.3 2.4 2.5 2.6 2.7 2.8 2.9 3.0'
P_AWGN_SPA = '.5 .75 1. 1.25 1.5 1.75 2. 2.25 2.5 2.75 3.'

stp = lambda init, step, count: [init + cnt * step for cnt in range(count)]       # simulations.py:15
sp = lambda ll: ' '.join('%g' % v for v in ll)                                     # simulations.py:14


def default_cases(code, mi=10, mw=100):
    """simulations.py:24-37 (exc_def_cases)."""
    tail = lambda cw: ['--codeword=%d' % cw, '--max-iter=%d' % mi, '--min-wec=%d' % mw]
    return [['bec', code, 'SPA'] + tail(0) + ['--params'] + P_BEC.split(),
            ['bsc', code, 'MSA'] + tail(1) + ['--params'] + P_BSC_MSA.split(),
            ['biawgn', code, 'MSA'] + tail(1) + ['--params'] + P_AWGN_MSA.split(),
            ['bsc', code, 'SPA'] + tail(0) + ['--params'] + sp(stp(.1, -.01, 7)).split(),
            ['biawgn', code, 'SPA'] + tail(0) + ['--params'] + P_AWGN_SPA.split()]


def ensemble(prefix, count):
    return [c for i in range(count) for c in default_cases('%s_%d' % (prefix, i + 1))]


def HMG():
    """simulations.py:48-60, decoders SPA / MSA only."""
    p_bec = '.5 .4 .3 .2 .1 .08 .06 .04 .02'
    p_bsc = p_bec + ' .25 .15 .01 .008 .006 .004 .002'
    code, config = '7_4_hamming', ['--codeword=1', '--min-wec=300']
    return [['bec', code, 'SPA', '--params'] + p_bec.split() + config] + \
           [['bsc', code, d, '--params'] + p_bsc.split() + config for d in ('SPA', 'MSA')] + \
           [['biawgn', code, d, '--params'] + sp(stp(2, .5, 11)).split() + config for d in ('SPA', 'MSA')]


def MAR():
    """simulations.py:62-71 (the ADMM lines are not this package's decoders)."""
    return default_cases('margulis')


def REG_BAD():
    """simulations.py:73-76."""
    return default_cases('1200_3_6_ldpc') + [c for mi in (0, 1, 2, 3, 6, 40, 100) for c in default_cases('1200_3_6_ldpc', mi)]


def REG_ENS():
    return ensemble('1200_3_6_rand_ldpc', 10)


def IREG_ENS():
    return ensemble('1200_rho_x5_rand_ldpc', 10)


all_cases = dict(HMG=HMG, MAR=MAR, REG_BAD=REG_BAD, REG_ENS=REG_ENS, IREG_ENS=IREG_ENS)


def main(argv=None):
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument('case', nargs='+', choices=sorted(all_cases))
    p.add_argument('--print', action='store_true', help='only print the case lines (what the reference does)')
    p.add_argument('--limit', type=int, default=0, help='run only the first LIMIT case lines of each named case')
    args, extra = p.parse_known_args(argv)
    comm = None
    out = []
    for name in args.case:
        cases = all_cases[name]()
        for case in (cases[:args.limit] if args.limit > 0 else cases):
            if args.print:
                print(' '.join(case + extra), flush=True)
                continue
            if comm is None:                       # ONE process group for every case (not an init / destroy per case)
                from .dist import Comm
                comm = Comm()
            out.append((case[:3], sim.main(case + extra, comm=comm)))
    if comm is not None:
        comm.close()
    return out


if __name__ == '__main__':
    main()
