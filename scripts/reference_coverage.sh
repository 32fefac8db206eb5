#!/bin/bash
# Statement coverage of the REFERENCE (its translated form, oracle/_ref/bgc_ref.c) under the
# oracle-vs-reference tests and the differential fuzzer: which reference statements does the
# pinning actually execute?  Prints the Fortran file:line of every statement never reached.
#   scripts/reference_coverage.sh [fuzz rounds]        (needs gcov; run from the repository root)
set -e
ROUNDS=${1:-300}
W=$(mktemp -d)
cp oracle/_ref/bgc_ref.c $W/
( cd $W && gcc -O0 --coverage -ffp-contract=off -fno-math-errno -fPIC -w -shared -o libbgc_ref_cov.so bgc_ref.c -lm )
export BGC_REF_LIB=$W/libbgc_ref_cov.so
python scripts/fuzz_oracle_vs_reference.py $ROUNDS 5000 | tail -1
python -m pytest tests/test_reference_translated.py -q -x 2>&1 | tail -1
( cd $W && gcov -o . libbgc_ref_cov.so-bgc_ref.gcno > gcov.log 2>&1 )
python - $W/bgc_ref.c.gcov <<'PY'
import re, sys
cur = func = None
missing, total, hit = {}, 0, 0
for ln in open(sys.argv[1]).read().splitlines():
    m = re.match(r'\s*([^:]+):\s*(\d+):(.*)', ln)
    if not m:
        continue
    cnt, txt = m.group(1).strip(), m.group(3)
    c = re.search(r'/\* (\w+\.F90):(\d+) (subroutine|function) (\w+) \*/', txt)
    if c:
        func = c.group(4)
    c = re.match(r'\s*/\* (\w+\.F90):(\d+) \*/', txt)
    if c:
        cur = (c.group(1), int(c.group(2)))
        continue
    if cur is None or cnt == '-':
        continue
    total += 1
    if cnt == '#####':
        missing.setdefault((cur[0], func), set()).add(cur[1])
    else:
        hit += 1
print("translated statements of the reference executed: %d of %d (%.1f %%)" % (hit, total, 100.0 * hit / total))
for (f, fn), s in sorted(missing.items()):
    print("  never reached: %s %s lines %s" % (f, fn, ", ".join(str(x) for x in sorted(s))))
PY
rm -rf $W
