#!/usr/bin/env python
"""Top source lines (samples / instructions) from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None
smp = collections.Counter(); ins = collections.Counter(); txt = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iS = hdr.index("# Samples"); iX = hdr.index("Instructions Executed"); continue
    if hdr is None or r[0] == "": continue
    try: ln = int(r[0])
    except ValueError: continue
    try: s = int(r[iS]); x = int(r[iX])
    except ValueError: continue
    smp[(cur, ln)] += s; ins[(cur, ln)] += x; txt[(cur, ln)] = r[1]
tot = sum(smp.values()); toti = sum(ins.values())
print("total samples", tot, "warp-inst", toti)
for key, s in smp.most_common(top):
    print("%-16s %5d %5.2f%% smp %5.2f%% inst  %s" % (key[0], key[1], 100 * s / tot, 100 * ins[key] / toti, txt[key].strip()[:100]))
byfile = collections.Counter()
for k, s in smp.items(): byfile[k[0]] += s
print({k: round(100 * v / tot, 1) for k, v in byfile.items()})
