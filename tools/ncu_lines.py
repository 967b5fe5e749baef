#!/usr/bin/env python
"""Aggregate warp-stall samples of an .ncu-rep per CUDA source line (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    lines = []
    fname = ""
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        if len(r) >= 5 and r[0].strip().isdigit() and r[2] == "-":
            try:
                lines.append((fname, int(r[0]), r[1].strip()[:100], int(r[4] or 0)))
            except ValueError:
                pass
    tot = sum(x[3] for x in lines)
    print("total samples", tot)
    for f, ln, src, s in sorted(lines, key=lambda x: -x[3])[:top]:
        print(f"{s:6d} {100 * s / max(tot, 1):5.1f}%  {f}:{ln}: {src}")


if __name__ == "__main__":
    main()
