#!/usr/bin/env python
"""Wall-clock throughput of the batched Monte-Carlo driver (ldpc_decoders_b200.sim.main, --noise device): the whole
application path — Python loop, on-device channel, decode, error counting, counters — not just the kernel."""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import _golden as G
    from ldpc_decoders_b200 import sim
    name = sys.argv[1] if len(sys.argv) > 1 else "1200_3_6_rand_ldpc_1"
    frames = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20000000
    d = tempfile.mkdtemp()
    m, n, rows, cols = G.code_tables(name)
    with open(os.path.join(d, name + ".txt"), "w") as fp:
        for c in range(m):
            fp.write(" ".join(str(v + 1) for v in cols[rows == c]) + "\n")
    for decoder, cw, dtype in (("MSA", 1, "f32"), ("SPA", 0, "f32"), ("MSA", 1, "f64")):
        for batch in (32768, 131072):
            args = ["biawgn", name, decoder, "--codeword", str(cw), "--params", "2.0", "--max-iter", "10", "--batch", str(batch),
                    "--dtype", dtype, "--seed", "5", "--noise", "device", "--frames", str(frames), "--data_dir", d, "--codes-dir", d]
            warm = list(args)
            warm[warm.index("--frames") + 1] = str(batch * 2)
            sim.main(warm)                                         # builds the engine caches, loads the kernels
            t0 = time.perf_counter()
            sim.main(args)
            dt = time.perf_counter() - t0
            print("%s %s %s batch %6d: %.2f s for %d frames = %.2f M frames/s" % (name, decoder, dtype, batch, dt, frames, frames / dt / 1e6), flush=True)


if __name__ == "__main__":
    main()
