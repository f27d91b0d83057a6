#!/usr/bin/env python
"""profiles/r2/sass_*.txt: mnemonic histogram + the data-movement lines of the hot kernels, from
`cuobjdump -sass ldpc_decoders_b200/libldpc_b200.so` (runs without a GPU).   python scripts/sass_excerpts.py [outdir]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"resident_vq": "resident_vqILi0ELi6ELi3ELi320ELi1200ELi2ELi4EE", "resident_bec": "resident_becILb0ELi2ELi4ELi320EE",
        "cn_sweep_tma": "cn_sweep_tmaIfLi4ELi0ELi6ELb1EE", "vn_sweep": "vn_sweepIfLi4ELi3ELb1EE",
        "resident_vd": "resident_vdILi6ELi3ELi320ELi1200ELb0EE"}
KEY = r"(UBLKCP|SYNCS|LDS\.128|STS\.128|FMNMX3|MUFU|REDUX|VOTE|SHFL|DMUL|DSETP|LDG\.E\.128|STG\.E\.128|LOP3|ATOMS|BAR)"


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2")
    os.makedirs(out, exist_ok=True)
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "ldpc_decoders_b200", "libldpc_b200.so")],
                         capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    for name, key in WANT.items():
        f = [x for x in funcs if key in x.split("\n")[0]]
        if not f:
            print("missing", name)
            continue
        body = f[0]
        ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body)
        short = collections.Counter(".".join(i.split(".")[:2]) if i.startswith(("LDS", "STS", "LDG", "STG", "UBLKCP", "SYNCS", "MUFU", "ATOMS", "RED", "BAR", "SHFL"))
                                    else i.split(".")[0] for i in ins)
        with open(os.path.join(out, "sass_%s.txt" % name), "w") as fp:
            fp.write("# cuobjdump -sass ldpc_decoders_b200/libldpc_b200.so, function %s\n# %d static instructions (sm_100a); mnemonic "
                     "histogram, then the lines that show how data moves\n" % (body.split("\n")[0].strip(), len(ins)))
            for k, v in short.most_common(40):
                fp.write("%-14s %5d\n" % (k, v))
            fp.write("\n# bulk-copy engine / mbarrier / 128-bit shared accesses / special instructions (first occurrences)\n")
            seen = collections.Counter()
            for line in body.split("\n"):
                m = re.search(r"/\*[0-9a-f]{4}\*/\s+((?:@!?U?P\d+\s+)?" + KEY + r"[^;]*;)", line)
                if m and seen[m.group(2)] < 3:
                    seen[m.group(2)] += 1
                    fp.write(line.strip()[:140] + "\n")
        print(name, len(ins), dict(short.most_common(5)))


if __name__ == "__main__":
    main()
