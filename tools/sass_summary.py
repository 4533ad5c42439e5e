#!/usr/bin/env python
"""profiles/sass_summary.txt: per-kernel SASS evidence of the shipped libb200sparse.so (cuobjdump -sass).

    python tools/sass_summary.py > profiles/r2_sass_summary.txt

Counts, per kernel, the mnemonics that show HOW the kernel moves and computes: UBLKCP (1-D TMA bulk copy
global->shared), SYNCS (mbarrier), LDG/STG widths, shuffles, and the FP64/FP32 arithmetic.  No tensor-core
instruction is expected (the path is bandwidth-bound, 0.13 flop/B)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "eigen-git-mirror_b200", "lib", "libb200sparse.so")
PATTERNS = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDG.128", r"\bLDG\.[A-Z.]*128"), ("LDG.64", r"\bLDG\.[A-Z.]*64\b"),
            ("LDG(all)", r"\bLDG\b"), ("STG.128", r"\bSTG\.[A-Z.]*128"), ("STG(all)", r"\bSTG\b"), ("LDS", r"\bLDS\b"),
            ("SHFL", r"\bSHFL"), ("DFMA", r"\bDFMA"), ("DMUL", r"\bDMUL"), ("DADD", r"\bDADD"), ("FFMA", r"\bFFMA"),
            ("FMUL", r"\bFMUL"), ("FADD", r"\bFADD"), ("ATOM/RED", r"\b(ATOMG|ATOM|RED)\b"),
            ("HMMA/UTCMMA(tensor)", r"\b(HMMA|IMMA|DMMA|UTC[A-Z]*MMA)"), ("MEMBAR/FENCE", r"\b(MEMBAR|FENCE)")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", sass)
    if names:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print(f"# {os.path.relpath(LIB, ROOT)}  ({os.path.getsize(LIB)} bytes), cubin arch: {', '.join(arch)}")
    print("# columns: " + " | ".join(n for n, _ in PATTERNS) + " | instructions")
    cur, counts, total = None, None, 0
    rows = []

    def flush():
        if cur is not None:
            rows.append((demangle.get(cur, cur), counts, total))

    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            cur, counts, total = m.group(1), collections.Counter(), 0
            continue
        if cur is None or "/*" not in line:
            continue
        ins = re.search(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not ins:
            continue
        total += 1
        for name, pat in PATTERNS:
            if re.search(pat, ins.group(1)):
                counts[name] += 1
    flush()
    for name, c, tot in sorted(rows):
        short = re.sub(r"\(.*", "", name).replace("b200s::", "")
        print(f"{short:58s} " + " ".join(f"{c.get(n, 0):5d}" for n, _ in PATTERNS) + f" {tot:7d}")
    tensor = sum(c.get("HMMA/UTCMMA(tensor)", 0) for _, c, _ in rows)
    print(f"# kernels: {len(rows)}; tensor-core instructions: {tensor}; "
          f"UBLKCP sites: {sum(c.get('UBLKCP', 0) for _, c, _ in rows)}")


if __name__ == "__main__":
    sys.exit(main())
