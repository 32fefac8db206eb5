#!/usr/bin/env python
"""Dynamic opcode histogram + stall-sample histogram from `ncu --page source --csv --print-source sass`."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iX, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops = collections.Counter(); samp = collections.Counter(); tot = 0; totS = 0
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
stalls = collections.Counter()
for r in rows[2:]:
    if len(r) <= iX: continue
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[iS])
    if not m: continue
    op = m.group(2)
    n = int(r[iX] or 0); s = int(r[iN] or 0)
    ops[op] += n; samp[op] += s; tot += n; totS += s
    for i in stall_cols:
        stalls[hdr[i]] += int(r[i] or 0)
cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
print("total warp-inst", tot, "samples", totS, ("thread-inst per cell %.0f" % (tot * 32 / cells)) if cells else "")
for op, n in ops.most_common(28):
    print("%-8s %6.2f%% inst  %6.2f%% samples" % (op, 100.0 * n / tot, 100.0 * samp[op] / max(1, totS)))
print({k: round(100.0 * v / max(1, totS), 1) for k, v in stalls.most_common(10)})
