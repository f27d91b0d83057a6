#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export (SASS view): dynamic warp-instruction mix by opcode, shared-memory
wavefronts and stall samples, optionally split at the kernel's barriers into phases.

    python scripts/ncu_source_summary.py gpurun_out/resident_msa_source.csv [--phases]
"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        src = r[ix["Source"]]
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        op = m.group(2) if m else src[:12]
        out.append(dict(addr=r[ix["Address"]], src=src, op=op, n=int(r[ix["Instructions Executed"]] or 0),
                        samples=int(r[ix["# Samples"]] or 0), wf=int(r[ix["L1 Wavefronts Shared"]] or 0),
                        wf_ideal=int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)))
    return out


def short(op):
    if op.startswith(("LDS", "STS", "LDL", "STL", "MUFU", "LDG", "STG", "FMNMX", "ATOMS", "BAR")):
        return ".".join(op.split(".")[:2])
    return op.split(".")[0]


def table(ins, title):
    tot = sum(i["n"] for i in ins) or 1
    ts = sum(i["samples"] for i in ins) or 1
    c, s = collections.Counter(), collections.Counter()
    for i in ins:
        c[short(i["op"])] += i["n"]
        s[short(i["op"])] += i["samples"]
    print("== %s: %d warp-instructions, %d samples, shared wavefronts %d (ideal %d)"
          % (title, tot, ts, sum(i["wf"] for i in ins), sum(i["wf_ideal"] for i in ins)))
    for op, n in c.most_common(22):
        print("   %-12s %6.2f%%   samples %5.2f%%" % (op, 100 * n / tot, 100 * s[op] / ts))
    return tot


def main():
    ins = load(sys.argv[1])
    total = table(ins, "whole kernel")
    if "--phases" in sys.argv:
        cuts = [k for k, i in enumerate(ins) if i["op"].startswith("BAR")]
        prev = 0
        for k in cuts + [len(ins)]:
            seg = ins[prev:k + 1]
            n = sum(i["n"] for i in seg)
            if n > 0.02 * total:
                table(seg, "SASS %s..%s (%.1f%% of instructions)" % (seg[0]["addr"], seg[-1]["addr"], 100 * n / total))
            prev = k + 1


if __name__ == "__main__":
    main()
