"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every
function include/bgc_b200.h declares, its structs match the ctypes mirror, the
host-side parameter initialisers reproduce the reference defaults, and the
compute entry points fail LOUDLY without a GPU (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import parity

pkg = parity.pkg
abi = pkg.abi
HEADER = os.path.join(parity.REPO, "include", "bgc_b200.h")


def declared_functions():
    src = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    names = re.findall(r"^\s*(?:const\s+char\s*\*|int)\s*(\w+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_functions_are_listed_and_exported(built):
    host = pkg.host
    decl = declared_functions()
    assert len(decl) >= 35
    assert sorted(host.ABI_SYMBOLS) == decl
    for flavour in ("prod", "strict"):
        L = host.lib(flavour)
        for name in decl:
            assert hasattr(L, name), (flavour, name)
    assert b"sm_100a" in host.lib().bgc_version()


def test_library_is_sm100a_and_has_no_cpu_path(built):
    so = os.path.join(parity.REPO, "ocean-bgc_b200", "csrc", "libbgc_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    # the product library must not link or reference the oracle
    nm = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
    assert "oracle_" not in nm


def test_sweep_sass_has_no_local_memory_and_stages_through_tma(built):
    """Static check on the built library (what profiles/sass_summary_r02.txt records): every
    instantiation of the column sweep is free of local-memory (spill) accesses, brings its inputs in
    by bulk copies through the TMA unit completed on an mbarrier, and the hot kernels are all present."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("sass_summary", os.path.join(parity.REPO, "scripts", "sass_summary.py"))
    ss = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ss)
    so = os.path.join(parity.REPO, "ocean-bgc_b200", "csrc", "libbgc_b200.so")
    per = ss.sass_counts(so)
    sweep = {k: c for k, c in per.items() if "eco_columns_kernel" in k}
    assert len(sweep) == 6          # three diagnostics modes x two block shapes
    for k, c in sweep.items():
        assert c["LDL"] == 0 and c["STL"] == 0, k
        assert c["UBLKCP"] >= 1 and c["SYNCS"] >= 1, k
        assert c["DFMA"] > 500 and c["RCP64H"] > 50, k     # FP64 body, reciprocal-seed divisions
    for kernel in ("co3_cells_kernel", "dms_cells_kernel", "dms_columns_kernel", "macros_cells_kernel",
                   "surface_fluxes_kernel", "zsat_columns_kernel", "inventory_fold_kernel", "co2calc_points_kernel"):
        assert any(kernel in k for k in per), kernel
    res = ss.ptxas_resources()
    if res:     # the ptxas logs exist where the library was built (not on the GPU box)
        for k, r in res.items():
            if "eco_columns_kernel" in k:
                assert r["spill_st"] == 0 and r["spill_ld"] == 0 and r["stack"] == 0, (k, r)


def test_struct_layouts_match_the_c_compiler(built):
    names = ["BgcParams", "BgcAutotroph", "BgcIndices", "DmsParams", "DmsIndices", "MacrosParams",
             "MacrosIndices", "BgcInput", "BgcForcing", "BgcOutput", "BgcFluxDiagnostics", "BgcDiagnostics",
             "DmsInput", "DmsForcing", "DmsOutput", "DmsFluxDiagnostics", "DmsDiagnostics", "MacrosInput",
             "MacrosOutput", "MacrosDiagnostics", "BgcStatus"]
    prog = '#include <stdio.h>\n#include "bgc_b200.h"\nint main(void){\n'
    for n in names:
        prog += 'printf("%s %%zu\\n", sizeof(%s));\n' % (n, n)
    prog += 'printf("off_epsC %zu\\n", __builtin_offsetof(BgcParams, epsC));\n'
    prog += 'printf("off_kFe %zu\\n", __builtin_offsetof(BgcAutotroph, kFe));\nreturn 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.dirname(HEADER), src, "-o", exe])
        out = dict(l.split() for l in subprocess.check_output([exe], text=True).splitlines())
    for n in names:
        assert int(out[n]) == C.sizeof(getattr(abi, n)), n
    assert int(out["off_epsC"]) == abi.BgcParams.epsC.offset
    assert int(out["off_kFe"]) == abi.BgcAutotroph.kFe.offset


def test_every_member_offset_matches_the_c_compiler(built):
    """abi.py parses the header with regular expressions; gcc is the authority on where each member
    of each struct lies."""
    names = [n for n in abi._STRUCT_FIELDS if hasattr(abi, n)]
    assert len(names) >= 21
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "bgc_b200.h"\nint main(void){\n'
    want = {}
    for n in names:
        for fname, _ in abi._STRUCT_FIELDS[n]:
            prog += 'printf("%s.%s %%zu %%zu\\n", offsetof(%s, %s), sizeof(((%s *)0)->%s));\n' % (n, fname, n, fname, n, fname)
            fld = getattr(getattr(abi, n), fname)
            want["%s.%s" % (n, fname)] = (fld.offset, fld.size)
    prog += "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "o.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "o")
        subprocess.check_call(["gcc", "-I", os.path.dirname(HEADER), src, "-o", exe])
        got = {}
        for line in subprocess.check_output([exe], text=True).splitlines():
            k, off, size = line.split()
            got[k] = (int(off), int(size))
    assert got == want
    assert len(got) >= 350


def test_parameter_tables_match_reference_defaults(built):
    """bgc_host_parms.c (product) and parms_oracle.c (oracle) are two independent
    restatements of BGC_parms_init / DMS_parms_init / MACROS_parms_init."""
    o = parity.oracle()
    p, q = pkg.host.Parms(), o.Parms()
    for a, b in ((p.bgc, q.bgc), (p.ind, q.ind), (p.dms, q.dms), (p.macros, q.macros),
                 (p.dms_ind, q.dms_ind), (p.macros_ind, q.macros_ind), (p.autotrophs, q.autotrophs)):
        assert bytes(a) == bytes(b), type(a).__name__
    # spot values straight from BGC_parms.F90:524-697
    assert p.bgc.parm_POC_diss == 88.0e2 and p.bgc.parm_kappa_nitrif == 0.06 * (1.0 / 86400.0)
    sp, diat, diaz, phaeo = p.autotrophs
    assert sp.imp_calcifier == 1 and diaz.Nfixer == 1 and phaeo.grazee_ind == 2
    assert phaeo.temp_function == abi.DEFINES["BGC_TFNC_QUASI_MMRT"] and diat.kSiO3 == 0.8
    assert diaz.Qp == 0.002735 and sp.Qp == 0.00855
    assert diat.Si_ind == p.ind.diatSi_ind and sp.CaCO3_ind == p.ind.spCaCO3_ind and diaz.Si_ind == 0
    # permuting the host-chosen tracer slots re-wires the autotroph indices (BGC_mod.F90:271-321)
    perm = np.random.default_rng(3).permutation(30)
    p.permute_tracers(perm); q.permute_tracers(perm)
    assert bytes(p.autotrophs) == bytes(q.autotrophs) and bytes(p.ind) == bytes(q.ind)


def test_no_gpu_means_loud_failure(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    host = pkg.host
    with pytest.raises(host.BgcError) as e:
        host.Context(10, 10)
    assert "no CPU fallback" in str(e.value)
    L = host.lib()
    assert L.bgc_ctx_create(C.c_int(0), C.c_int(10), C.c_int(10), C.byref(C.c_void_p())) == abi.DEFINES["BGC_ERR_NO_DEVICE"]


def test_missing_library_raises(built, monkeypatch):
    host = pkg.host
    monkeypatch.setitem(host.LIB_NAME, "ghost", "libbgc_does_not_exist.so")
    with pytest.raises(host.BgcError):
        host.lib("ghost")


def test_sharding_slabs():
    sh = pkg.sharding
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 235160, 3693225):
            s = sh.slabs(world, n)
            assert s[0][0] == 0 and sum(x[1] for x in s) == n
            for (a, na), (b, nb) in zip(s, s[1:]):
                assert b == a + na and 0 <= na - nb <= 1
    with pytest.raises(ValueError):
        sh.slab(2, 2, 10)


def test_generated_fortran_bindings_are_in_sync_with_the_header(tmp_path):
    """ocean-bgc_b200/fortran/bgc_b200_capi.F90 and the *_ptrs.inc files are generated from the
    X-macro lists of include/bgc_b200.h (gen_capi.py): regenerating them must change nothing."""
    import hashlib
    import shutil
    import sys
    pkg_dir = os.path.join(parity.REPO, "ocean-bgc_b200")
    shutil.copytree(os.path.join(parity.REPO, "include"), tmp_path / "include")
    os.makedirs(tmp_path / "ocean-bgc_b200")
    shutil.copy(os.path.join(pkg_dir, "abi.py"), tmp_path / "ocean-bgc_b200" / "abi.py")
    shutil.copytree(os.path.join(pkg_dir, "fortran"), tmp_path / "ocean-bgc_b200" / "fortran")
    work = tmp_path / "ocean-bgc_b200" / "fortran"
    subprocess.check_call([sys.executable, str(work / "gen_capi.py")], stdout=subprocess.DEVNULL)
    fdir = os.path.join(pkg_dir, "fortran")
    names = ["bgc_b200_capi.F90"] + sorted(n for n in os.listdir(fdir) if n.endswith(".inc"))
    assert len(names) == 6
    for n in names:
        a = hashlib.sha256(open(os.path.join(fdir, n), "rb").read()).hexdigest()
        b = hashlib.sha256((work / n).read_bytes()).hexdigest()
        assert a == b, "%s is out of date: run ocean-bgc_b200/fortran/gen_capi.py" % n
