#!/usr/bin/env python
"""Per source line, the executed SASS instructions by opcode (from `ncu --page source --csv --print-source cuda,sass`).
Usage: ncu_line_ops.py file.csv <warp-levels> [opcode-regex] [top]  -> lines ranked by executed instructions matching
the regex, per warp-level."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
unit = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else re.compile(".")
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
cur = None; hdr = None; line = None
cnt = collections.Counter(); smp = collections.Counter(); txt = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iX = hdr.index("Instructions Executed"); iN = hdr.index("# Samples"); continue
    if hdr is None: continue
    if r[0] != "":
        line = (cur, int(r[0])); txt[line] = r[1].strip(); continue
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[3])
    if not m or not pat.search(m.group(2)): continue
    try: cnt[line] += int(r[iX] or 0); smp[line] += int(r[iN] or 0)
    except ValueError: pass
tot = sum(cnt.values())
print("matching instructions per unit: %.1f" % (tot / unit))
for k, v in cnt.most_common(top):
    print("%-14s %5d %7.2f  smp %5d  %s" % (k[0], k[1], v / unit, smp[k], txt[k][:110]))
