#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass` output: dynamic instruction mix by opcode,
stall samples by opcode, and the hottest SASS addresses.  Usage: ncu_src_summary.py src.csv [cell-warps] [kernel index in the file]"""
import csv, re, sys
from collections import Counter, defaultdict
rows = list(csv.reader(open(sys.argv[1])))
cells_warps = float(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 4096 / 32
# the file may hold several kernels; take the first, or the one named by the third argument (0-based)
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hdr = rows[starts[which]]
end = starts[which + 1] - 1 if len(starts) > which + 1 else len(rows)
data = rows[starts[which] + 1:end]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ex, smp = Counter(), Counter()
stall_by = defaultdict(Counter)
tot_ex = tot_smp = 0
hot = []
for r in data:
    if len(r) <= iex or not r[iex]:
        continue
    m = re.match(r"\s*(@!?U?P\d\s+)?([A-Z0-9_]+)", r[isrc])
    if not m:
        continue
    op = m.group(2)
    e = int(r[iex] or 0); s = int(r[ismp] or 0)
    ex[op] += e; smp[op] += s; tot_ex += e; tot_smp += s
    for c in stall_cols:
        v = int(r[c] or 0)
        if v:
            stall_by[hdr[c]][op] += v
    hot.append((s, e, r[ia], r[isrc][:70]))
print("total warp-instr %d  = %.1f per cell-warp; samples %d" % (tot_ex, tot_ex / cells_warps, tot_smp))
print("%-10s %12s %9s %8s" % ("opcode", "executed", "/cellwarp", "samples%"))
for op, e in ex.most_common(28):
    print("%-10s %12d %9.1f %8.1f" % (op, e, e / cells_warps, 100.0 * smp[op] / max(tot_smp, 1)))
print("\nstall reasons (share of samples) and the opcodes they sit on:")
for name, cnt in sorted(stall_by.items(), key=lambda kv: -sum(kv[1].values()))[:8]:
    t = sum(cnt.values())
    print("  %-22s %5.1f%%  %s" % (name, 100.0 * t / max(tot_smp, 1), ", ".join("%s %.0f%%" % (o, 100.0 * v / t) for o, v in cnt.most_common(4))))
print("\nhottest instructions:")
for s, e, a, src in sorted(hot, reverse=True)[:25]:
    print("  %6d smp %10d ex  %s  %s" % (s, e, a, src))
