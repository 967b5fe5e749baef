#!/usr/bin/env python
"""SASS evidence per kernel of fft_b200/_C/libspectre_mix.so (no GPU needed): counts of the Blackwell-specific mnemonics
(TMA: UTMALDG / UTMASTG / UTMAPF / UBLKCP; tensor memory: LDTM / STTM; packed fp32x2: FADD2 / FFMA2 / FMUL2; SHFL; tensor-core
MMA: UTC*MMA, expected 0 -- the path is butterfly / element-wise work; programmatic dependent launch: ACQBULK = griddepcontrol.wait,
PREEXIT = griddepcontrol.launch_dependents) and the total instruction count.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fft_b200", "_C", "libspectre_mix.so")
KEYS = ["UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "FADD2", "FFMA2", "FMUL2", "SHFL", "MUFU", "UTCMMA", "LDS", "STS", "MOV", "ACQBULK", "PREEXIT"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
            if m:
                op = m.group(1)
                per[cur]["TOTAL"] += 1
                if op.startswith("UTC") and "MMA" in op:
                    per[cur]["UTCMMA"] += 1
                elif op in KEYS:
                    per[cur][op] += 1
    names = subprocess.run(["c++filt"] + list(per), capture_output=True, text=True).stdout.strip().splitlines()
    tot = collections.Counter()
    rows = []
    for (mangled, c), name in zip(per.items(), names):
        name = name.replace("spx::", "").replace("(bool)", "").replace("(int)", "").replace("__nv_bfloat16", "bf16")
        name = re.sub(r"\((anonymous namespace::)?MixParams.*", "", name)
        name = re.sub(r"^void ", "", name)
        tot.update(c)
        rows.append((name, c))
    print(f"{len(rows)} kernels in {os.path.relpath(LIB, ROOT)}; whole library: " + ", ".join(f"{k} {tot[k]}" for k in KEYS + ["TOTAL"]))
    print("template arguments of spectre_mix_kernel: <Plan<radices, SUB>, mode (0 QUAD / 1 PAIR / 2 REAL), tile columns, compute threads, "
          "min CTAs/SM, in type, out type, HAS_MEM, RFFT_ONLY, TMA_IN, TMEM_IO, ANCH (gate from anchors), DGATE (gate gradient)>")
    want = sys.argv[1:] or ["Plan<16, 16, 16, 1, 0>, 0, 2, 512, 1, float, float, false, false, true, true, false, false",   # metric kernel (4096, TMEM-staged)
                            "Plan<16, 16, 16, 1, 0>, 0, 2, 512, 1, float, float, false, false, true, true, true, false",    # + gate from anchors
                            "Plan<16, 16, 16, 1, 0>, 0, 2, 512, 1, float, float, false, false, true, true, false, true",    # gate gradient
                            "Plan<16, 16, 16, 1, 2>, 0, 2, 512, 1, float, float, false, false, true, true, false, false",   # 8192 in one kernel (DIT2)
                            "Plan<16, 16, 16, 1, 1>, 0, 2, 512, 1, float, float, false, false, true, true, false, false",   # 16384 middle kernel (SUB)
                            "Plan<16, 16, 16, 1, 0>, 0, 2, 512, 1, bf16, bf16, false, false, true, true, false, false",     # bf16 I/O
                            "Plan<16, 16, 4, 1, 0>, 0, 8, 512, 1, float, float, false, false, true, true, false, false",    # 1024 wide rows
                            "Plan<4, 16, 16, 1, 0>, 0, 4, 256, 2, float, float, false, false, true, false, false, false",   # 1024, 16-channel tiles (cfg2)
                            "long_pass_kernel<4, float", "gate_expand_kernel", "decode_kernel<3>", "decode_reduce_kernel", "gate_interp_table"]
    for w in want:
        for name, c in rows:
            if w in name:
                print(f"\n{name}\n    " + "  ".join(f"{k} {c[k]}" for k in KEYS + ["TOTAL"]))


if __name__ == "__main__":
    main()
