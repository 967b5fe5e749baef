#!/usr/bin/env python
"""Dynamic SASS opcode histogram of the (single) kernel in an .ncu-rep captured with --import-source on:
warp-level instructions executed per opcode, optionally divided by a tile count.

    python tools/ncu_opcodes.py gpurun_out/prof.ncu-rep [tiles]
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
    i_s, i_e, i_n = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, samp, tot = collections.Counter(), collections.Counter(), 0
    for r in rows:
        if len(r) <= i_e or r is hdr:
            continue
        m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", r[i_s])
        if not m:
            continue
        try:
            n, s = int(r[i_e] or 0), int(r[i_n] or 0)
        except ValueError:
            continue
        ops[m.group(1)] += n
        samp[m.group(1)] += s
        tot += n
    print(f"warp instructions executed: {tot}  ({tot / tiles:.1f} per tile, tiles = {tiles:g})")
    for op, n in ops.most_common(30):
        print(f"  {op:12s} {n / tiles:10.1f} per tile  {100 * n / max(tot, 1):5.1f} %   stall samples {samp[op]}")


if __name__ == "__main__":
    main()
