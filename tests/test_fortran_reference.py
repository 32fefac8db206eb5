"""Pins the CPU oracle against the UNMODIFIED reference Fortran through a small driver PROGRAM
(tests/fortran/ref_driver.F90: reads one block of synthetic columns from a flat binary file, calls
the reference's own *_parms_init, *_init, BGC_SourceSink (cold, warm), BGC_SurfaceFluxes, DMS_*,
MACROS_* and dumps every output).  Two ways to build the driver + reference, same test body:

  * a Fortran compiler, if one exists (none in this image nor on the GPU box): the Makefile of
    tests/fortran/ compiles the reference sources where they lie - tolerance 1e-13 normalised
    (SURVEY.md 8c), since another compiler version may order a few operations differently;
  * oracle/f90c.py: driver and reference translated to C and compiled by gcc - bit-exact.  This
    also proves the driver and the file protocol of this test before any Fortran compiler sees them.

Nothing of the reference is copied into this repository; both builds go to oracle/_ref/.
"""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import parity  # noqa: E402

REF = os.environ.get("BGC_REFERENCE_DIR", "/root/reference")
FC = os.environ.get("FC", "gfortran")
OUT = os.path.join(parity.REPO, "oracle", "_ref")
MODS = ["BGC_parms", "co2calc", "BGC_mod", "DMS_parms", "DMS_mod", "MACROS_parms", "MACROS_mod"]

abi = parity.abi


def build_with_fortran_compiler():
    if shutil.which(FC) is None or not os.path.isdir(REF):
        pytest.skip("needs a Fortran compiler (%s) and the reference tree (%s)" % (FC, REF))
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "fortran"), "FC=" + FC, "REF=" + REF])
    return os.path.join(OUT, "ref_driver"), 1e-13


def build_with_f90c():
    exe = os.path.join(OUT, "ref_driver_translated")
    if os.path.isfile(os.path.join(REF, "BGC_mod.F90")):
        fdir = os.path.join(HERE, "fortran")
        subprocess.check_call([sys.executable, "gen_alloc.py"], cwd=fdir, stdout=subprocess.DEVNULL)
        os.makedirs(OUT, exist_ok=True)
        csrc = os.path.join(OUT, "ref_driver.c")
        subprocess.check_call([sys.executable, os.path.join(parity.REPO, "oracle", "f90c.py"), "-o", csrc]
                              + [os.path.join(REF, m + ".F90") for m in MODS]
                              + [os.path.join(fdir, "ref_driver.F90")], stdout=subprocess.DEVNULL)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-math-errno", "-w", "-DREF_TLS=",
                               "-o", exe, csrc, "-lm"])
    elif not os.path.exists(exe):
        pytest.skip("needs the reference tree (%s) once, or the prebuilt oracle/_ref/ref_driver_translated" % REF)
    return exe, 0.0


def _dump(f, *arrays):
    for a in arrays:
        np.asfortranarray(a).ravel(order="F").tofile(f)


@pytest.mark.parametrize("build", [build_with_fortran_compiler, build_with_f90c], ids=["fortran-compiler", "f90c"])
def test_oracle_matches_the_reference_driver(tmp_path, build):
    exe, TOL = build()
    o = parity.oracle()
    po = o.Parms()
    nL, nC, nCols = 60, 96, 90
    cols, dms, mac = parity.make_bgc(nL, nC, po, nColumns=nCols, ragged=True, with_dms=True, with_macros=True)
    parity.poison_outputs(cols)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    F = cols.forcing
    with open(fin, "wb") as f:
        np.array([nL, nC, nCols, 1], dtype=np.int32).tofile(f)
        np.array([po.bgc.T0_Kelvin_BGC], dtype=np.float64).tofile(f)
        _dump(f, cols.BGC_tracers, cols.PotentialTemperature, cols.Salinity, cols.cell_center_depth,
              cols.cell_thickness, cols.cell_bottom_depth, cols.cell_latitude)
        cols.number_of_active_levels.astype(np.int32).tofile(f)
        _dump(f, F["FESEDFLUX"], F["dust_FLUX_IN"], F["ShortWaveFlux_surface"], F["surfacePressure"], F["iceFraction"],
              F["windSpeedSquared10m"], F["atmCO2"], F["atmCO2_ALT_CO2"], F["surface_pH"], F["surface_pH_alt_co2"],
              F["surfaceDepth"], F["SST"], F["SSS"], F["depositionFlux"], F["riverFlux"], F["gasFlux"],
              F["seaIceFlux"], F["netFlux"])
        _dump(f, cols.PH_PREV_3D, cols.PH_PREV_ALT_CO2_3D, dms.DMS_tracers, mac.MACROS_tracers)
    subprocess.check_call([exe, fin, fout])

    # the oracle on the same inputs, same call sequence as the driver
    o.BGC_SourceSink(po, cols, True)
    ph_cold = cols.PH_PREV_3D.copy()
    o.BGC_SourceSink(po, cols, True)
    o.BGC_SurfaceFluxes(po, cols)
    o.DMS_SourceSink(po, dms); o.DMS_SurfaceFluxes(po, dms)
    o.MACROS_SourceSink(po, mac)

    raw = np.fromfile(fout, dtype=np.float64)
    pos = 0

    def take(shape):
        nonlocal pos
        n = int(np.prod(shape))
        a = raw[pos:pos + n].reshape(shape, order="F")
        pos += n
        return a
    n2, nT = (nL, nC), abi.BGC_TRACER_CNT
    checks = [("BGC_tendencies", take((nL, nC, nT)), cols.BGC_tendencies), ("ph_cold", take(n2), ph_cold),
              ("PH_PREV_3D", take(n2), cols.PH_PREV_3D), ("PH_PREV_ALT_CO2_3D", take(n2), cols.PH_PREV_ALT_CO2_3D),
              ("netFlux", take((nC, nT)), F["netFlux"]), ("gasFlux", take((nC, nT)), F["gasFlux"]),
              ("surface_pH", take((nC,)), F["surface_pH"]), ("surface_pH_alt_co2", take((nC,)), F["surface_pH_alt_co2"]),
              ("iceFraction", take((nC,)), F["iceFraction"]),
              ("DMS_tendencies", take((nL, nC, abi.DMS_TRACER_CNT)), dms.DMS_tendencies),
              ("DMS netFlux", take((nC, abi.DMS_TRACER_CNT)), dms.forcing["netFlux"]),
              ("MACROS_tendencies", take((nL, nC, abi.MACROS_TRACER_CNT)), mac.MACROS_tendencies)]
    act = (np.arange(1, nL + 1)[:, None] <= np.where(np.arange(nC) < nCols, cols.number_of_active_levels, 0)[None, :])
    for name in abi.BGC_DIAG_K2:
        checks.append((name, take(n2), cols.diag[name]))
    for name in abi.BGC_DIAG_KA:
        checks.append((name, take((nL, nC, abi.BGC_AUTOTROPH_CNT)), cols.diag[name]))
    for name in abi.BGC_DIAG_CA:
        checks.append((name, take((nC, abi.BGC_AUTOTROPH_CNT)), cols.diag[name]))
    for name in abi.BGC_DIAG_C1:
        checks.append((name, take((nC,)), cols.diag[name]))
    for name in abi.BGC_FLUX_DIAG:
        checks.append((name, take((nC,)), cols.flux_diag[name]))
    masked = []   # DMS / MACROS diagnostics are defined on active cells only
    for name in abi.DMS_DIAG:
        masked.append((name, take(n2), dms.diag[name]))
    for name in abi.DMS_FLUX_DIAG:
        checks.append((name, take((nC,))[:nCols], dms.flux_diag[name][:nCols]))
    for name in abi.MACROS_DIAG:
        masked.append((name, take(n2), mac.diag[name]))
    assert pos == raw.size
    worst = {}
    for name, ref, got in checks:
        worst[name] = parity.nerr(got, ref)
    for name, ref, got in masked:
        worst[name] = parity.nerr(got[act], ref[act])
    bad = {k: v for k, v in worst.items() if not v <= TOL}
    assert not bad, bad
