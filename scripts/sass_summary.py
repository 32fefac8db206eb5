#!/usr/bin/env python
"""Static summary of the built product library: per kernel instantiation the ptxas resource line
(registers, spill bytes, stack frame, static shared memory) and the SASS instruction counts that
show how the kernel is built (FP64 arithmetic, reciprocal seeds, global / shared accesses, barriers,
bulk copies through the TMA unit, mbarrier operations).

Runs where the library is built (no GPU needed):
    python scripts/sass_summary.py > profiles/sass_summary_rNN.txt
"""
import collections
import glob
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(REPO, "ocean-bgc_b200", "csrc")
CUOBJDUMP = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")
CUFILT = os.environ.get("CUFILT", "/usr/local/cuda/bin/cu++filt")

# column -> regular expression on the SASS mnemonic (first token after the predicate)
COUNTS = [
    ("DFMA", r"DFMA"), ("DMUL", r"DMUL"), ("DADD", r"DADD"), ("DSETP", r"DSETP"),
    ("RCP64H", r"MUFU\.RCP64H"), ("LDG", r"LDG"), ("STG", r"STG"), ("LDS", r"LDS"), ("STS", r"STS"),
    ("LDL", r"LDL"), ("STL", r"STL"), ("LDC", r"LDC"), ("BAR", r"BAR"), ("UBLKCP", r"UBLKCP"),
    ("SYNCS", r"SYNCS"), ("SHFL", r"SHFL"), ("VOTE", r"VOTE"),
]


def demangle(names):
    out = subprocess.run([CUFILT] + names, capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(name):
    """bgc::(anonymous namespace)::eco_columns_kernel<1, 256, 1>(bgc::EcoArgs) -> eco_columns_kernel<1,256,1>"""
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = re.sub(r"\((int|bool)\)", "", name)
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*\)$", "", name)
    name = name.replace("bgc::", "").replace(", ", ",")
    return name


def ptxas_resources():
    """mangled entry name -> dict(registers, spill_st, spill_ld, stack, smem) from build/*.ptxas.log"""
    res = {}
    for log in sorted(glob.glob(os.path.join(CSRC, "build", "*.ptxas.log"))):
        cur = None
        for line in open(log):
            m = re.search(r"Compiling entry function '([^']+)'", line)
            if m:
                cur = res.setdefault(m.group(1), {})
                continue
            if cur is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and "stack" not in cur:      # the first such line after the entry is the entry's own
                cur.update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
            m = re.search(r"Used (\d+) registers", line)
            if m and "registers" not in cur:
                cur["registers"] = int(m.group(1))
                s = re.search(r"(\d+) bytes smem", line)
                cur["smem"] = int(s.group(1)) if s else 0
    return res


def sass_counts(lib):
    sass = subprocess.run([CUOBJDUMP, "-sass", lib], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        cur["total"] += 1
        for col, rx in COUNTS:
            if re.match(rx + r"(\.|$)", op):
                cur[col] += 1
    return per


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(CSRC, "libbgc_b200.so")
    head = subprocess.run(["git", "-C", REPO, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    res = ptxas_resources()
    per = sass_counts(lib)
    names = demangle(list(per.keys()))
    print(f"# {os.path.relpath(lib, REPO)} (sm_100a) at commit {head}: ptxas -v resources and static SASS instruction")
    print("# counts per kernel instantiation (scripts/sass_summary.py).  UBLKCP = cp.async.bulk through the TMA unit,")
    print("# SYNCS = mbarrier operations, RCP64H = MUFU reciprocal seed, LDL/STL = local memory (spills), LDC = constant")
    print("# bank loads that are not FMA operands.  regs / spill st+ld bytes / stack bytes / static smem bytes from ptxas.")
    rows = []
    for mangled, c in per.items():
        r = res.get(mangled, {})
        rows.append((short(names.get(mangled, mangled)), r, c))
    rows.sort(key=lambda t: t[0])
    w = max(len(t[0]) for t in rows) + 2
    for name, r, c in rows:
        rtxt = (f"regs={r.get('registers', '?')} spill={r.get('spill_st', '?')}+{r.get('spill_ld', '?')} "
                f"stack={r.get('stack', '?')} smem={r.get('smem', '?')}")
        ctxt = " ".join(f"{col}={c[col]}" for col, _ in COUNTS)
        print(f"{name:<{w}}{rtxt:<42} total={c['total']} {ctxt}")


if __name__ == "__main__":
    main()
