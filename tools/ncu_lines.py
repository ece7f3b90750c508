"""Per-source-line stall samples / executed instructions from `ncu -i rep --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_lines.py export.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if len(r) > 20 and r[0] == 'Line No')
iS, iE = hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
agg = {}
for r in rows:
    if len(r) < len(hdr) or r[0] == 'Line No' or r[2] != '-':
        continue
    agg[int(r[0])] = (int(r[iS] or 0), int(r[iE] or 0), r[1], {n: int(r[hdr.index(n)] or 0) for n in stalls})
totS = sum(v[0] for v in agg.values()); totE = sum(v[1] for v in agg.values())
print("samples", totS, "warp instructions", totE)
tot = {n: sum(v[3][n] for v in agg.values()) for n in stalls}
print({k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
for ln in sorted(agg, key=lambda l: -agg[l][0])[:top_n]:
    s, e, t, st = agg[ln]
    big = {k[6:]: v for k, v in st.items() if v > 0.15 * s and v > 20}
    print(f"{ln:5d} {s:6d} {100 * s / totS:5.1f}% {e / 1e6:7.1f}M  {t.strip()[:100]}  {big}")
