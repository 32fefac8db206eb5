"""Unit tests of oracle/f90c.py itself — the Fortran-90-subset -> C translator that runs the
reference in an image without a Fortran compiler.  The pinning of the
oracle is only as good as the translator's reading of Fortran, so every rule the bit-exact
comparison leans on is tested here in isolation on a small Fortran module written for the
purpose (tests/f90c_cases/semantics.F90), against expectations derived BY HAND from the
language rules and from gfortran's documented code generation - never from the translator.
"""
import ctypes as C
import math
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(REPO, "oracle"))
import ref_translated as rt   # noqa: E402  (test infrastructure only)


@pytest.fixture(scope="module")
def L(tmp_path_factory):
    d = tmp_path_factory.mktemp("f90c")
    c, meta, so = str(d / "sem.c"), str(d / "sem.json"), str(d / "libsem.so")
    subprocess.check_call([sys.executable, os.path.join(REPO, "oracle", "f90c.py"), "-o", c, "-m", meta,
                           os.path.join(HERE, "f90c_cases", "semantics.F90")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-math-errno", "-fPIC", "-w", "-shared",
                           "-o", so, c, "-lm"])
    return rt.TLib(so, meta)


def _vec(n):
    return (C.c_double * n)()


def test_arithmetic_rules(L):
    x, y = 0.1, 3.0
    out = _vec(12)
    L.call("sem_mod__arith", x, y, 7, 2, out)
    assert out[0] == 3.0 and out[1] == -3.0                       # truncation toward zero
    assert out[2] == (x / y) * y and out[3] == (x + y) - y        # left to right, no reassociation
    assert out[4] == -(x * x)                                     # unary minus binds weaker than **
    assert out[5] == 0.0                                          # integer(8) constant 0.5 -> 0
    assert out[6] == x / 10.0
    assert out[7] == float(np.float32(1.00e-8)) != 1.00e-8        # REAL(4) literal widened
    assert out[8] == 1.00e-8
    assert out[9] == 7.0                                          # private names do not leak through USE
    assert out[10] == 1.0 / (365.0 * 86400.0)
    assert out[11] == 6.0 * x


def test_power_lowering(L):
    x, n = 1.7, 5
    out = _vec(8)
    L.call("sem_mod__powers", x, n, out)
    assert out[0] == x * x                                        # __builtin_powi(x, 2)
    assert out[1] == x * x * x                                    # (x*x)*x
    # x**n with a run-time n: libgcc's __powidf2 (square and multiply)
    y, p, m = (x if n % 2 else 1.0), x, n
    while True:
        m >>= 1
        if not m:
            break
        p = p * p
        if m % 2:
            y = y * p
    assert out[2] == y
    assert out[3] == math.pow(x, 1.5)                             # libm pow, no sqrt rewrite
    assert out[4] == math.pow(10.0, -x) == out[5]
    assert out[6] == 32.0
    assert out[7] == 1.0 / (x * x)


def test_min_max_merge(L):
    out = _vec(6)
    L.call("sem_mod__minmax", 0.25, 0.75, out)
    assert list(out) == [0.75, 0.25, 0.75, 0.75, 0.0, 0.5]
    L.call("sem_mod__minmax", 0.0, -0.0, out)
    # gfortran: m = a1; if (a2 > m) m = a2  ->  the first argument survives a tie of signed zeros
    assert math.copysign(1.0, out[0]) == 1.0 and math.copysign(1.0, out[1]) == 1.0
    L.call("sem_mod__minmax", 2.0, 1.0, out)
    assert out[3] == 2.0 and out[4] == -1.0


def test_sum_and_sections(L):
    v = (C.c_double * 4)(1e16, 1.0, -1e16, 1.0)
    m = np.asfortranarray(np.arange(1.0, 9.0).reshape(2, 4, order="F"))   # m(i,j) = i + 2(j-1)
    pair = L.struct("kinds_mod__pair")
    pick = (pair * 4)()
    for i, n in enumerate((4, 4, 1, 2)):
        pick[i].n = n
    out = _vec(6)
    L.call("sem_mod__sums", v, rt.describe(m), pick, out)
    assert out[0] == ((((0.0 + 1e16) + 1.0) - 1e16) + 1.0) == 1.0         # sequential: the first 1.0 is lost
    assert out[1] == 2.0
    assert out[2] == 2.0 + 4.0 + 6.0 + 8.0
    assert out[3] == 7.0 + 7.0 + 1.0 + 3.0
    assert out[4] == 8.0
    assert math.isinf(out[5])                                     # exp(1e16) overflows, as in libm


def test_loops(L):
    out = _vec(5)
    L.call("sem_mod__loops", 5, out)
    # i=1:1, i=2 skipped, i=3:3, i=4: j=1,2 then cycle outer at j=3 -> 2, i=5:5
    assert list(out) == [11.0, 6.0, 7.0, 22.0, 6.0]


def test_select_case(L):
    got = []
    for k in (1, 2, 5, 7, 8, 9, 10, 0):
        _, b = L.call("sem_mod__selects", k, 0.0)
        got.append(b[1].value)
    assert got == [10.0, 20.0, 20.0, 30.0, 30.0, 30.0, -1.0, -1.0]


def test_derived_types_and_allocatables(L):
    n, m = 3, 4
    grid = np.zeros((n, m), order="F")
    line, idx = np.ones(5), np.zeros(6, dtype=np.int32)
    b, keep = L.fill("kinds_mod__bag", {"grid": grid, "line": line, "idx": idx}, {"flag": 0})
    out = _vec(4)
    L.call("sem_mod__bags", b, n, m, out)
    i, j = np.meshgrid(np.arange(1, n + 1), np.arange(1, m + 1), indexing="ij")
    assert np.array_equal(grid, 1.5 + i + 10 * j)                 # column-major, 1-based indexing
    assert np.all(line == 0.25) and np.all(idx == 3) and b.flag == 1
    assert list(out) == [1.5 + n + 10 * m, float(n), float(m), 1.0]


def test_character_assignment(L):
    names = np.full((4, 16), ord("?"), dtype=np.uint8)
    b = L.struct("kinds_mod__bag")()
    d = rt.FA()
    d.p, d.n1, d.n2, d.n3 = names.ctypes.data, 4, 1, 1
    b.names = d
    p = L.struct("kinds_mod__pair")()
    p.n = 2
    L.call("sem_mod__strings", b, p)
    txt = [bytes(r).decode() for r in names]
    assert txt[0] == "abChl".ljust(16)                            # trim, //, blank padding
    assert txt[1] == "ab".ljust(16)                               # 17 characters truncated to 16
    assert txt[2] == "this text is lon"                           # truncation on assignment
    assert txt[3] == "x".ljust(16)                                # (:) section assignment


def test_module_state_and_argument_passing(L):
    r1, _ = L.call("sem_mod__saved", 1.5)
    r2, _ = L.call("sem_mod__saved", 2.0)
    assert (r1, r2) == (1.5, 3.5)
    assert L.var("kinds_mod__module_state").value == 3.5
    p = L.struct("kinds_mod__pair")()
    p.a, p.b, p.n = 1.0, 10.0, 3
    out = _vec(3)
    L.call("sem_mod__by_ref", p, out)
    assert p.b == 12.0 and list(out) == [12.0, 13.0, -13.0]       # by reference; expressions by temporary
    out2 = _vec(2)
    L.call("sem_mod__keyword_caller", out2)
    assert list(out2) == [-3.0, 3.0]                              # keyword argument


def test_named_constants_are_exported(L):
    assert L.const("kinds_mod__r8") == 8 and L.const("kinds_mod__i8") == 8 and L.const("kinds_mod__i4") == 4
    assert L.const("kinds_mod__int_half") == 0 and L.meta()["consts"]["kinds_mod__int_half"] == "i8"
    assert L.const("kinds_mod__single_lit") == float(np.float32(1e-8))


@pytest.fixture(scope="module")
def LI(tmp_path_factory):
    """tests/f90c_cases/interop.F90 + its C side: the ISO_C_BINDING group of constructs"""
    d = tmp_path_factory.mktemp("f90c_interop")
    c, meta, so = str(d / "interop.c"), str(d / "interop.json"), str(d / "libinterop.so")
    subprocess.check_call([sys.executable, os.path.join(REPO, "oracle", "f90c.py"), "-o", c, "-m", meta,
                           os.path.join(HERE, "f90c_cases", "interop.F90")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-w", "-shared", "-o", so, c,
                           os.path.join(HERE, "f90c_cases", "interop_c.c"), "-lm"])
    return rt.TLib(so, meta)


def test_c_interoperability(LI):
    """bind(C) types with default-initialised c_ptr components, interface bodies with value and
    by-reference dummies, c_loc of an allocatable component, c_associated, array constructors with
    kind suffixes, module variables with initialisers."""
    h = LI.struct("interop_mod__holder")()
    vals = np.arange(1.0, 6.0)
    h.values = rt.describe(vals)
    rc, _ = LI.call("interop_mod__scale_through_c", h, 2.5)
    assert rc == 5 + 1000 * 6                       # n, spare was set by C (no +100), c_sum3 = 6
    assert np.array_equal(vals, 2.5 * np.arange(1.0, 6.0))
    empty = LI.struct("interop_mod__holder")()      # not allocated -> NULL data, n = 0
    rc, _ = LI.call("interop_mod__scale_through_c", empty, 2.0)
    assert rc == 0 + 1000 * 6
    assert LI.var("interop_mod__calls").value == 2 and LI.var("interop_mod__seen").value == 1


def test_internal_procedures_and_character_dummies(LI):
    h = LI.struct("interop_mod__holder")()
    tags = np.full((3, 8), ord("?"), dtype=np.uint8)
    d = rt.FA()
    d.p, d.n1, d.n2, d.n3 = tags.ctypes.data, 3, 1, 1
    h.tags = d
    # a character(len=*) dummy travels as the translator's string value {int n; char s[1024]}
    class FStr(rt.C.Structure):
        _fields_ = [("n", rt.C.c_int), ("s", rt.C.c_char * 1024)]
    f = FStr(4, b"ab  ")
    fn = getattr(LI.lib(), "interop_mod__label_all")
    fn.restype = None
    fn(rt.C.byref(h), f)
    assert [bytes(r).decode() for r in tags] == ["ab-odd  ", "ab-even ", "ab-odd  "]


def test_c_strings_and_while_loops(LI):
    n, _ = LI.call("interop_mod__read_c_string")
    assert n == len("thirteen char")
    calls = LI.var("interop_mod__calls").value
    r, _ = LI.call("interop_mod__count_to", 10)
    assert r == 12 + calls                          # 3, 6, 9, 12
    r, _ = LI.call("interop_mod__count_to", 1000)
    assert r == 102 + calls                         # the `;`-separated exit inside the loop


def test_unsupported_constructs_stop_the_translation(tmp_path):
    src = tmp_path / "bad.F90"
    src.write_text("module m\n implicit none\ncontains\n subroutine s(x)\n  real(8) :: x\n"
                   "  where (x > 0) x = 1\n end subroutine\nend module\n")
    r = subprocess.run([sys.executable, os.path.join(REPO, "oracle", "f90c.py"), "-o", str(tmp_path / "o.c"),
                        str(src)], capture_output=True, text=True)
    assert r.returncode != 0 and "bad.F90:6" in (r.stderr + r.stdout)
