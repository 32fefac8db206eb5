"""The drop-in Fortran shim's public interface against the reference's, read by a Fortran parser that
shares nothing with this repo: numpy.f2py's crackfortran.

tests/test_fortran_shim.py executes the shim through oracle/f90c.py - the translator written for
this repo.  Here an independent parser reads both `ocean-bgc_b200/fortran/<module>.F90` and
`/root/reference/<module>.F90` and the test holds, for every public procedure of the reference
(`BGC_mod.F90:59-63`, `DMS_mod.F90`, `MACROS_mod.F90`, `co2calc.F90:24`):
  * the same name, the same dummy-argument names in the same order;
  * per argument the same type / derived-type name / kind / rank / `optional`;
  * the same intent - except that the shim may say `intent(in)` where the reference declares none
    (every caller that compiles against the reference compiles against that), and may add `target`
    (module procedures have explicit interfaces);
  * the same public module entities (the shim may export more, e.g. the batched `*_points` forms).
Needs the reference sources: skipped where `/root/reference` does not exist (the GPU box).
"""
import contextlib
import io
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("BGC_REFERENCE_DIR", "/root/reference")
SHIM = os.path.join(REPO, "ocean-bgc_b200", "fortran")
MODULES = ["BGC_mod", "DMS_mod", "MACROS_mod", "co2calc"]

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present")


def crack(path):
    """The module block of `path` as numpy.f2py.crackfortran sees it."""
    cf = pytest.importorskip("numpy.f2py.crackfortran")
    sink = io.StringIO()
    cwd = os.getcwd()
    os.chdir(os.path.dirname(path))         # include files are looked up beside the source
    try:
        with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
            cf.verbose = 0
            blocks = cf.crackfortran([path])
    finally:
        os.chdir(cwd)
    mods = [b for b in blocks if b["block"] == "module"]
    assert len(mods) == 1, path
    return mods[0]


def procedures(mod):
    return {b["name"]: b for b in mod["body"] if b["block"] in ("subroutine", "function")}


def public_names(mod):
    return {k for k, v in mod["vars"].items() if "public" in v.get("attrspec", [])}


def characteristics(var):
    c = {k: var.get(k) for k in ("typespec", "typename", "kindselector", "charselector", "dimension")}
    c["attrspec"] = sorted(a for a in var.get("attrspec", []) if a != "target")
    return c


@pytest.mark.parametrize("module", MODULES)
def test_shim_interface_is_the_reference_interface(module):
    ref = crack(os.path.join(REF, module + ".F90"))
    shim = crack(os.path.join(SHIM, module + ".F90"))
    assert ref["name"] == shim["name"]
    missing = public_names(ref) - public_names(shim)
    assert not missing, "public entities of the reference the shim does not export: %s" % sorted(missing)
    rp, sp = procedures(ref), procedures(shim)
    checked = 0
    for name in sorted(public_names(ref) & set(rp)):
        assert name in sp, name
        r, s = rp[name], sp[name]
        assert r["block"] == s["block"], name
        assert r["args"] == s["args"], (name, r["args"], s["args"])
        for a in r["args"]:
            rv, sv = r["vars"][a], s["vars"][a]
            assert characteristics(rv) == characteristics(sv), (name, a, characteristics(rv), characteristics(sv))
            ri, si = rv.get("intent"), sv.get("intent")
            assert ri == si or (ri is None and si == ["in"]), (name, a, ri, si)
            checked += 1
    assert checked >= {"BGC_mod": 18, "DMS_mod": 15, "MACROS_mod": 8, "co2calc": 39}[module]


def test_shim_links_the_reference_parameter_modules_unchanged():
    """The derived types ARE the API: the shim must not carry its own copies of the *_parms modules
    (its Makefile compiles the reference's), so that a host built against the reference's types
    passes the very same types."""
    for name in ("BGC_parms.F90", "DMS_parms.F90", "MACROS_parms.F90"):
        assert not os.path.exists(os.path.join(SHIM, name)), name
    mk = open(os.path.join(SHIM, "Makefile")).read()
    for name in ("BGC_parms", "DMS_parms", "MACROS_parms"):
        assert name in mk
    assert "$(REF)/co2calc" not in mk       # the GPU-backed co2calc module replaces the reference's CPU one
