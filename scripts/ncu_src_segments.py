#!/usr/bin/env python
"""Split the SASS of a kernel in an `ncu --page source --csv` dump into runs of instructions with the same execution count (= the same
set of warps: all four, one role, the loader threads, per-chunk code) and print size, FP64 / memory instruction counts and the share of
stall samples per run.  usage: ncu_src_segments.py src.csv [kernel index in the file]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = rows[starts[which]]
end = starts[which + 1] - 1 if len(starts) > which + 1 else len(rows)
data = rows[starts[which] + 1:end]
iex, ismp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
segs, cur = [], None
for k, r in enumerate(data):
    if len(r) <= iex or not r[iex]:
        continue
    m = re.match(r"\s*(@!?U?P\d\s+)?([A-Z0-9_.]+)", r[isrc])
    if not m:
        continue
    e = int(r[iex]); s = int(r[ismp] or 0); op = m.group(2)
    if cur is None or abs(e - cur["e"]) > 0.02 * max(e, cur["e"], 1):
        cur = dict(e=e, n=0, smp=0, k0=k, fp64=0, shfl=0, lds=0, sts=0, gmem=0); segs.append(cur)
    cur["n"] += 1; cur["smp"] += s
    cur["fp64"] += op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
    cur["shfl"] += op.startswith("SHFL"); cur["lds"] += op.startswith("LDS"); cur["sts"] += op.startswith("STS")
    cur["gmem"] += op.startswith(("LDG", "STG", "LDGSTS", "LDL", "STL"))
tot = sum(s["smp"] for s in segs)
print("kernel", which, "samples", tot)
for s in segs:
    if s["n"] >= 6 and s["e"] > 100000:
        print("  at %5d  executed %8d  instr %4d  fp64 %4d  lds %3d  sts %3d  shfl %3d  gmem/local %3d  samples %5.1f%%" % (s["k0"], s["e"], s["n"], s["fp64"], s["lds"], s["sts"], s["shfl"], s["gmem"], 100 * s["smp"] / tot))
