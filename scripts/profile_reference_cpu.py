"""Where the CPU time of the REFERENCE goes: the translated reference (oracle/f90c.py) built with
-DREF_PROFILE counts inclusive rdtsc cycles and calls per procedure.  One thread, one block of the
synthetic EC60to30 columns, cold pass then warm passes.

    python scripts/profile_reference_cpu.py [columns]     (writes nothing; prints a table)
"""
import ctypes as C
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import parity                    # noqa: E402
import ref_translated as rt      # noqa: E402

so = "/tmp/libbgc_ref_prof.so"
subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-math-errno", "-fPIC", "-w", "-DREF_PROFILE",
                       "-shared", "-o", so, os.path.join(rt.REFDIR, "bgc_ref.c"), "-lm"])
L = rt.TLib(so, rt.META_PATH)
o = parity.oracle()
po = o.Parms()
rp = rt.RefParms(po, L=L)
nC = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
bgc, dms, mac = parity.make_bgc(60, nC, po, with_dms=True, with_macros=True)
lib = L.lib()
lib.ref_prof_name.restype = C.c_char_p
lib.ref_prof_cycles.restype = C.c_ulonglong
lib.ref_prof_ncalls.restype = C.c_ulonglong


def table(title, cells):
    rows = [(lib.ref_prof_name(i).decode(), lib.ref_prof_cycles(i), lib.ref_prof_ncalls(i))
            for i in range(lib.ref_prof_count())]
    rows = [r for r in rows if r[2]]
    top = max(r[1] for r in rows)
    print("\n%s  (%d cells; inclusive cycles, %% of the largest entry)" % (title, cells))
    for n, cyc, calls in sorted(rows, key=lambda r: -r[1]):
        print("  %-40s %6.1f %%  %10d calls  %8.0f cycles/call  %7.1f calls/cell" % (
            n, 100.0 * cyc / top, calls, cyc / calls, calls / cells))


cells = int(bgc.active_mask().sum())
rt.BGC_SourceSink(rp, bgc, True)
table("BGC_SourceSink, cold brackets", cells)
lib.ref_prof_reset()
t0 = time.perf_counter()
rt.BGC_SourceSink(rp, bgc, True)
dt = time.perf_counter() - t0
table("BGC_SourceSink, warm brackets (%.2f us per cell on one thread)" % (dt / cells * 1e6), cells)
lib.ref_prof_reset()
rt.BGC_SurfaceFluxes(rp, bgc); rt.DMS_SourceSink(rp, dms); rt.DMS_SurfaceFluxes(rp, dms); rt.MACROS_SourceSink(rp, mac)
table("BGC_SurfaceFluxes + DMS_* + MACROS_SourceSink", cells)
