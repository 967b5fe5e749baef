#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel of an object / shared library (no GPU needed).

    python tools/sass_hist.py <file.o|.so> <substring of the mangled kernel name> [top]
e.g. the metric kernel:  tools/sass_hist.py fft_b200/_C/obj/spectre_mix_inst_4096.o 'ELi2ELi512ELi1EffLb0ELb0ELb1ELb1ELb0E'
"""
import collections
import re
import subprocess
import sys


def main():
    path, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, hist = None, collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and pat in cur:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
            if m:
                hist[m.group(1)] += 1
    tot = sum(hist.values())
    print("kernel pattern", pat, "total SASS instructions", tot)
    for op, n in hist.most_common(top):
        print(f"  {op:12s} {n:6d}")


if __name__ == "__main__":
    main()
