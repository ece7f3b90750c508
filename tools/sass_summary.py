"""Mnemonic counts per kernel from `cuobjdump -sass` of the built library (the evidence table of profiles/*_sass_summary.md).
usage: python tools/sass_summary.py [lib.so] > profiles/rN_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "freesplat_b200", "libfreesplat_b200.so")
COLS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "ELECT", "FFMA2", "FMUL2", "FADD2", "RED", "REDG", "ATOMG", "LDGSTS", "MUFU.EX2", "CCTL"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
dem = {}
counts = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1)
        counts[cur]["total"] += 1
        for c in COLS:
            if op == c or op.startswith(c + "."):
                counts[cur][c] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS evidence (`cuobjdump -sass freesplat_b200/libfreesplat_b200.so`, `python tools/sass_summary.py`)\n")
print("Mnemonic counts per kernel: `UTCHMMA` = tcgen05.mma (kind::tf32), `LDTM` / `STTM` = tcgen05.ld / st, `UTMALDG` = TMA tensor load, "
      "`UTCBAR` = tcgen05.commit, `ELECT` = elect.sync (single-lane MMA issue), `FFMA2` / `FMUL2` / `FADD2` = packed fp32, `RED` / `REDG` / "
      "`ATOMG` = global reductions / atomics (incl. the peer reductions of the fused reduce-scatter), `LDGSTS` = cp.async, `CCTL` = "
      "prefetch.global.L1.\n")
print("| kernel | " + " | ".join(COLS) + " | total instructions |")
print("|---|" + "---|" * (len(COLS) + 1))
for mangled, name in sorted(zip(counts, names), key=lambda kv: kv[1]):
    c = counts[mangled]
    short = re.sub(r"\(.*", "", name)
    print(f"| `{short}` | " + " | ".join(str(c[k]) if c[k] else "" for k in COLS) + f" | {c['total']} |")
