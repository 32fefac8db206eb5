#!/usr/bin/env python3
"""f90c — a Fortran-90-subset to C translator, written for ONE purpose: to turn the
unmodified reference sources of E3SM-Project/Ocean-BGC (/root/reference/*.F90) into a
shared library that can be RUN in an image that has no Fortran compiler.

TEST INFRASTRUCTURE ONLY (see bgc_oracle.h).  The output goes to oracle/_ref/ (git-ignored;
nothing derived from the reference sources is committed) and is used to PIN the hand-written
oracle: the translation is mechanical — statement by statement, expression by expression,
parentheses and left-to-right evaluation order kept, `x**n` handed to the same GCC builtin
(`__builtin_powi`) that gfortran's front end emits, literals kept in the precision they are
written in (a `1.00e-8` stays a float32 constant) — and the C is compiled by the same GCC
middle/back end that `gfortran -O2` uses (`gcc -O2 -ffp-contract=off -fno-math-errno`, no
`-march`: plain x86-64 has no FMA, so gfortran contracts nothing either).

Supported subset = what the reference uses, nothing more: modules, `use`, named constants,
derived types with scalar / fixed-shape / allocatable components, module variables (emitted
thread-local: the reference keeps solver state in module SAVE variables), subroutines and
functions with by-reference arguments (positional and keyword), `if / else if / else`, counted,
endless and `while` loops, construct names with `exit` / `cycle`, `return`, `select case` on
integers, `allocate / deallocate`, whole-array and `(:)`-section assignments, vector subscripts
inside `sum(..., dim=1)`, `merge`, `max`, `min`, `size`, array constructors, character assignment
with `trim` and `//`.  A second group of constructs exists so that the drop-in shim of
ocean-bgc_b200/fortran/ can be executed as well: `use, intrinsic :: iso_c_binding`, `type, bind(C)`,
interface bodies with `bind(C, name=...)` and `value` dummies (external C functions), `c_ptr`,
`c_loc`, `c_associated`, `c_f_pointer`, internal procedures (emitted as GCC nested functions),
`character(len=*)` dummies, pointer arrays, default-initialised components, `include`, `;`, and
the little I/O an error path needs (`write` to a unit, internal `read` of an integer,
`get_environment_variable`, `error stop`).  A third, small group runs the stand-alone driver of
tests/fortran/: a main `program`, `open / read / write / close` of unformatted stream files,
`get_command_argument`.  Anything else stops the translation with the file
and line.  The translator is lenient where a compiler is strict: it does not check conformance.

-DREF_POISON fills every ALLOCATE with NaN patterns; -DREF_PROFILE counts inclusive cycles and
calls per procedure (ref_prof_*); -DREF_TLS= makes the module variables plain globals.

Usage:  f90c.py -o OUT.c -m META.json  A.F90 B.F90 ...   (files in module-dependency order)
"""
import json
import re
import sys

# ----------------------------------------------------------------------------- source reader


class Line:
    __slots__ = ("text", "file", "no")

    def __init__(self, text, file, no):
        self.text, self.file, self.no = text, file, no


def strip_comment(s):
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out)


def split_semicolons(s):
    out, cur, q = [], [], None
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == ";":
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur))
    return [x.strip() for x in out if x.strip()]


def read_source(path, defines=()):
    """cpp conditionals (#ifdef/#ifndef/#else/#endif), comments, continuation lines, INCLUDE
    lines, `;` statement separators."""
    short = path.rsplit("/", 1)[-1]
    lines, stack, active = [], [], True
    cur, cur_no = "", 0
    for no, raw in enumerate(open(path, encoding="latin-1"), 1):
        s = raw.rstrip("\n")
        m = re.match(r"""\s*include\s+(['"])(.+?)\1\s*$""", s, re.I)
        if m and active and not cur:
            inc = path.rsplit("/", 1)[0] + "/" + m.group(2) if "/" in path else m.group(2)
            lines.extend(read_source(inc, defines))
            continue
        if s.lstrip().startswith("#"):
            d = s.lstrip()[1:].split()
            if d[0] in ("ifdef", "ifndef"):
                stack.append(active)
                cond = d[1] in defines
                active = active and (cond if d[0] == "ifdef" else not cond)
            elif d[0] == "else":
                active = stack[-1] and not active
            elif d[0] == "endif":
                active = stack.pop()
            else:
                raise SystemExit(f"{short}:{no}: unsupported cpp directive {s!r}")
            continue
        if not active:
            continue
        s = strip_comment(s).replace("\t", " ").rstrip()
        if not s.strip():
            continue
        if cur:
            t = s.lstrip()
            if t.startswith("&"):
                t = t[1:]
            cur += " " + t
        else:
            cur, cur_no = s, no
        if cur.endswith("&"):
            cur = cur[:-1]
            continue
        for part in split_semicolons(cur):
            lines.append(Line(part, short, cur_no))
        cur = ""
    return lines


# ----------------------------------------------------------------------------- tokens

TOK = re.compile(r"""
   (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
 | (?P<dot>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.)
 | (?P<num>(?:\d+\.(?![a-zA-Z]+\.)\d*|\.\d+|\d+)(?:[edED][+-]?\d+)?(?:_\w+)?)
 | (?P<id>[A-Za-z]\w*)
 | (?P<op>\*\*|//|==|/=|<=|>=|=>|::|\(/|/\)|[-+*/()=,:%<>\[\]])
 | (?P<ws>\s+)
""", re.X | re.I)


def tokenize(s, where):
    out, pos = [], 0
    while pos < len(s):
        m = TOK.match(s, pos)
        if not m:
            raise SystemExit(f"{where}: cannot tokenize at {s[pos:pos+20]!r}")
        pos = m.end()
        k = m.lastgroup
        v = m.group(k)
        if k == "ws":
            continue
        if k in ("id", "dot"):
            v = v.lower()
        if k == "num":
            v = v.lower()
        out.append((k, v))
    return out


# ----------------------------------------------------------------------------- AST

class Node:
    pass


class Num(Node):
    def __init__(self, text):
        self.text = text


class Str(Node):
    def __init__(self, val):
        self.val = val


class Log(Node):
    def __init__(self, val):
        self.val = val


class Bin(Node):
    def __init__(self, op, l, r):
        self.op, self.l, self.r = op, l, r


class Un(Node):
    def __init__(self, op, e):
        self.op, self.e = op, e


class Paren(Node):
    def __init__(self, e):
        self.e = e


class ArrCons(Node):
    def __init__(self, items):
        self.items = items


class Colon(Node):
    def __init__(self, lo=None, hi=None):
        self.lo, self.hi = lo, hi


class Kw(Node):
    def __init__(self, name, e):
        self.name, self.e = name, e


class Ref(Node):
    """designator: parts = [(name, args|None), ...] joined by %"""

    def __init__(self, parts):
        self.parts = parts


class Parser:
    def __init__(self, toks, where):
        self.t, self.i, self.where = toks, 0, where

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else (None, None)

    def next(self):
        tk = self.peek()
        self.i += 1
        return tk

    def accept(self, v):
        if self.peek()[1] == v and self.peek()[0] != "str":
            self.i += 1
            return True
        return False

    def expect(self, v):
        if not self.accept(v):
            raise SystemExit(f"{self.where}: expected {v!r}, found {self.peek()[1]!r}")

    def at_end(self):
        return self.i >= len(self.t)

    # precedence climbing, Fortran order
    def expr(self):
        l = self.p_or()
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.next()[1]
            l = Bin(op, l, self.p_or())
        return l

    def p_or(self):
        l = self.p_and()
        while self.peek()[1] == ".or.":
            self.next()
            l = Bin(".or.", l, self.p_and())
        return l

    def p_and(self):
        l = self.p_not()
        while self.peek()[1] == ".and.":
            self.next()
            l = Bin(".and.", l, self.p_not())
        return l

    def p_not(self):
        if self.peek()[1] == ".not.":
            self.next()
            return Un(".not.", self.p_not())
        return self.p_rel()

    REL = {"==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">=",
           ".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}

    def p_rel(self):
        l = self.p_cat()
        k, v = self.peek()
        if k != "str" and v in self.REL:
            self.next()
            return Bin(self.REL[v], l, self.p_cat())
        return l

    def p_cat(self):
        l = self.p_add()
        while self.peek()[1] == "//" and self.peek()[0] == "op":
            self.next()
            l = Bin("//", l, self.p_add())
        return l

    def p_add(self):
        k, v = self.peek()
        if k == "op" and v in "+-":
            self.next()
            l = self.p_mul()
            if v == "-":
                l = Un("-", l)
        else:
            l = self.p_mul()
        while self.peek()[0] == "op" and self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            l = Bin(op, l, self.p_mul())
        return l

    def p_mul(self):
        l = self.p_pow()
        while self.peek()[0] == "op" and self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            l = Bin(op, l, self.p_pow())
        return l

    def p_pow(self):
        l = self.p_primary()
        if self.peek() == ("op", "**"):
            self.next()
            k, v = self.peek()
            if k == "op" and v in "+-":      # a ** -b (extension); keep it working
                self.next()
                r = self.p_pow()
                r = Un("-", r) if v == "-" else r
            else:
                r = self.p_pow()             # right associative
            return Bin("**", l, r)
        return l

    def p_primary(self):
        k, v = self.next()
        if k == "num":
            return Num(v)
        if k == "str":
            q = v[0]
            return Str(v[1:-1].replace(q + q, q))
        if k == "dot" and v in (".true.", ".false."):
            return Log(v == ".true.")
        if k == "op" and v == "(":
            e = self.expr()
            self.expect(")")
            return Paren(e)
        if k == "op" and v in ("(/", "["):
            items = [self.expr()]
            while self.accept(","):
                items.append(self.expr())
            self.expect("/)" if v == "(/" else "]")
            return ArrCons(items)
        if k == "id":
            parts = []
            name = v
            while True:
                args = None
                if self.peek() == ("op", "("):
                    self.next()
                    args = self.arglist()
                parts.append((name, args))
                if self.peek() == ("op", "%"):
                    self.next()
                    k2, name = self.next()
                    if k2 != "id":
                        raise SystemExit(f"{self.where}: component name expected")
                    continue
                break
            return Ref(parts)
        raise SystemExit(f"{self.where}: unexpected token {v!r}")

    def arglist(self):
        args = []
        if self.accept(")"):
            return args
        while True:
            args.append(self.arg())
            if self.accept(","):
                continue
            self.expect(")")
            return args

    def arg(self):
        if self.peek() == ("op", "*") and self.peek(1)[1] in (",", ")"):
            self.next()
            return Colon()            # assumed size / assumed length
        if self.peek()[0] == "id" and self.peek(1) == ("op", "="):
            name = self.next()[1]
            self.next()
            return Kw(name, self.expr())
        if self.peek() == ("op", ":"):
            self.next()
            if self.peek()[1] in (",", ")"):
                return Colon()
            return Colon(None, self.expr())
        e = self.expr()
        if self.peek() == ("op", ":"):
            self.next()
            if self.peek()[1] in (",", ")"):
                return Colon(e, None)
            return Colon(e, self.expr())
        return e


# ----------------------------------------------------------------------------- types / symbols

class T:
    """base: int | real | logical | char | type ; kind in bytes ; rank"""

    def __init__(self, base, kind=4, rank=0, tname=None, clen=None):
        self.base, self.kind, self.rank, self.tname, self.clen = base, kind, rank, tname, clen

    def scalar(self):
        return T(self.base, self.kind, 0, self.tname, self.clen)

    def ctype(self):
        if self.base == "cptr":
            return "void *"
        if self.base == "int":
            return {8: "long long", 1: "signed char", 2: "short"}.get(self.kind, "int")
        if self.base == "real":
            return "double" if self.kind == 8 else "float"
        if self.base == "logical":
            return "int"
        if self.base == "char":
            return "char"
        return f"struct {self.tname}"

    def code(self):
        if self.base == "cptr":
            return "cptr"
        if self.base == "int":
            return {8: "i8", 1: "i1", 2: "i2"}.get(self.kind, "i4")
        if self.base == "real":
            return "r8" if self.kind == 8 else "r4"
        if self.base == "logical":
            return "log"
        if self.base == "char":
            return f"char{self.clen}"
        return f"type:{self.tname}"


class Sym:
    def __init__(self, name, typ, dims=None, alloc=False, param=False, init=None, intent=None,
                 private=False):
        self.name, self.typ, self.dims, self.alloc = name, typ, dims, alloc
        self.param, self.init, self.intent, self.private = param, init, intent, private
        self.dummy = False
        self.cname = None        # C identifier of the object
        self.module = None
        self.is_result = False
        self.local_const = False
        self.pointer = False
        self.value = False


class DType:
    def __init__(self, name, cname):
        self.name, self.cname, self.fields = name, cname, []   # list of Sym


class Proc:
    def __init__(self, name, kind, args, result, module, line):
        self.name, self.kind, self.args, self.result, self.module, self.line = \
            name, kind, args, result, module, line
        self.syms, self.body = {}, []
        self.cname = f"{module.name}__{name}"
        self.rtype = None
        self.parent = None       # host procedure of an internal procedure
        self.children = []
        self.external = False    # interface body with bind(C): a C function defined elsewhere
        self.is_program = False  # the body of a main PROGRAM
        self.prefix_type = None


class Module:
    def __init__(self, name):
        self.name, self.uses, self.syms, self.types, self.procs = name, [], {}, {}, {}
        self.default_private = False


# ----------------------------------------------------------------------------- translator

INTRINSIC_ELEMENTAL = {"exp": "exp", "log": "log", "log10": "log10", "sqrt": "sqrt",
                       "tanh": "tanh", "sin": "sin", "cos": "cos", "atan": "atan"}

PRELUDE = r"""/* GENERATED by oracle/f90c.py from the reference Fortran sources — do not edit, do not commit. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifndef REF_TLS
#define REF_TLS __thread
#endif
typedef struct { void *p; int n1, n2, n3; } fa_t;
typedef struct { int n; char s[1024]; } fstr_t;
static inline fstr_t f_lit(const char *s, int n) { fstr_t r; r.n = n; memcpy(r.s, s, n); return r; }
static inline fstr_t f_var(const char *s, int n) { fstr_t r; r.n = n; memcpy(r.s, s, n); return r; }
static inline fstr_t f_trim(fstr_t a) { while (a.n > 0 && a.s[a.n - 1] == ' ') a.n--; return a; }
static inline fstr_t f_cat(fstr_t a, fstr_t b) { memcpy(a.s + a.n, b.s, b.n); a.n += b.n; return a; }
static inline int f_cmp(fstr_t a, fstr_t b) {   /* blank-padded comparison */
  int n = a.n > b.n ? a.n : b.n;
  for (int i = 0; i < n; i++) { unsigned char x = i < a.n ? a.s[i] : ' ', y = i < b.n ? b.s[i] : ' ';
    if (x != y) return x < y ? -1 : 1; }
  return 0; }
static inline void f_assign(char *d, int dn, fstr_t v) {
  int n = v.n < dn ? v.n : dn; memcpy(d, v.s, n); if (n < dn) memset(d + n, ' ', dn - n); }
/* MAX/MIN as gfortran expands them without -ffast-math: m = a1; if (a2 > m) m = a2; ... */
#define F_MAX2(T, a, b) ({ T _m = (a); T _b = (b); if (_b > _m) _m = _b; _m; })
#define F_MIN2(T, a, b) ({ T _m = (a); T _b = (b); if (_b < _m) _m = _b; _m; })
static inline long long f_ipow(long long b, long long e) {
  long long r = 1; if (e < 0) return (b == 1) ? 1 : (b == -1 ? ((e & 1) ? -1 : 1) : 0);
  while (e) { if (e & 1) r *= b; b *= b; e >>= 1; } return r; }
/* -DREF_PROFILE: inclusive cycles (rdtsc) and call counts per procedure, per thread */
#ifdef REF_PROFILE
#include <x86intrin.h>
#define REF_PROF_MAX 256
REF_TLS unsigned long long ref_prof_cyc[REF_PROF_MAX], ref_prof_calls[REF_PROF_MAX];
#define PROF_ENTER(id) const unsigned long long _pt0 = __rdtsc(); ref_prof_calls[id]++;
#define PROF_EXIT(id) ref_prof_cyc[id] += __rdtsc() - _pt0;
#else
#define PROF_ENTER(id)
#define PROF_EXIT(id)
#endif
/* unformatted stream I/O and the command line (only a main PROGRAM uses them) */
static FILE *f_units[100];
static int f_argc; static char **f_argv;
static void f_open(int u, fstr_t name, int old) {
  name = f_trim(name); name.s[name.n] = 0;
  f_units[u] = fopen(name.s, old ? "rb" : "wb");
  if (!f_units[u]) { fprintf(stderr, "cannot open %s\n", name.s); abort(); } }
static void f_io(int u, void *p, size_t bytes, int wr) {
  size_t n = wr ? fwrite(p, 1, bytes, f_units[u]) : fread(p, 1, bytes, f_units[u]);
  if (n != bytes) { fprintf(stderr, "unit %d: short %s (%zu of %zu bytes)\n", u, wr ? "write" : "read", n, bytes); abort(); } }
#ifdef REF_POISON
static void *f_alloc(size_t nbytes) { void *p = malloc(nbytes ? nbytes : 1); memset(p, 0xFF, nbytes); return p; }
#else
static void *f_alloc(size_t nbytes) { return malloc(nbytes ? nbytes : 1); }
#endif
"""


ISO_C_KINDS = {"c_int": 4, "c_long_long": 8, "c_size_t": 8, "c_double": 8, "c_float": 4, "c_char": 1,
               "c_signed_char": 1, "c_short": 2, "c_long": 8, "c_int64_t": 8, "c_int32_t": 4, "c_bool": 1}


class Translator:
    def __init__(self):
        self.modules = {}
        self.order = []
        iso = Module("iso_c_binding")
        for n, v in ISO_C_KINDS.items():
            sy = Sym(n, T("int", 4), param=True, init=Num(str(v)))
            sy.cname, sy.module = str(v), iso
            iso.syms[n] = sy
        sy = Sym("c_null_ptr", T("cptr", 8), param=True, init=Num("0"))
        sy.cname, sy.module = "((void *)0)", iso
        iso.syms["c_null_ptr"] = sy
        sy = Sym("c_null_char", T("char", 1, 0, None, 1), param=True, init=Str("\0"))
        sy.cname, sy.module = "f_lit(\"\\0\", 1)", iso
        iso.syms["c_null_char"] = sy
        self.modules["iso_c_binding"] = iso
        self.out = []
        self.tmp = 0
        self.cur_line = None
        self.scope_proc = None
        self.scope_mod = None

    # ---------------------------------------------------------------- errors
    def err(self, msg):
        ln = self.cur_line
        raise SystemExit(f"{ln.file}:{ln.no}: {msg}\n    {ln.text}")

    # ---------------------------------------------------------------- pass 1: structure
    def load(self, path):
        lines = read_source(path)
        mod = dtype = None
        stack = []            # open procedures: [module procedure, internal procedure]
        in_interface = False
        for ln in lines:
            self.cur_line = ln
            toks = tokenize(ln.text, f"{ln.file}:{ln.no}")
            w0 = toks[0][1] if toks[0][0] == "id" else None
            w1 = toks[1][1] if len(toks) > 1 and toks[1][0] == "id" else None
            proc = stack[-1] if stack else None
            if w0 == "module" and w1 and w1 != "procedure":
                mod = Module(w1)
                self.modules[w1] = mod
                self.order.append(mod)
                continue
            if w0 == "program" and w1 and len(toks) == 2:
                mod = Module(w1)
                self.modules[w1] = mod
                self.order.append(mod)
                pr = Proc("main", "subroutine", [], None, mod, ln)
                pr.is_program = True
                mod.procs["main"] = pr
                stack.append(pr)
                continue
            if w0 == "end" or w0 in ("endmodule", "endsubroutine", "endfunction", "endtype", "endinterface",
                                     "endprogram"):
                what = w1 if w0 == "end" else w0[3:]
                if what == "module":
                    mod = None
                    continue
                if what == "program":
                    stack.pop()
                    mod = None
                    continue
                if what in ("subroutine", "function"):
                    stack.pop()
                    continue
                if what == "type":
                    dtype = None
                    continue
                if what == "interface":
                    in_interface = False
                    continue
                if proc is None:
                    self.err("unexpected END")
            if mod is None:
                self.err("statement outside a module")
            self.scope_mod, self.scope_proc = mod, None
            if dtype is not None:
                self.decl(toks, dtype=dtype, mod=mod)
                continue
            if w0 == "interface" and len(toks) == 1:
                in_interface = True
                continue
            if proc is None or (w0 == "contains") or self._is_proc_header(toks):
                if w0 == "contains":
                    continue
                hdr = self.proc_header(toks, mod, ln, parent=proc)
                if hdr is not None:
                    hdr.external = in_interface
                    stack.append(hdr)
                    continue
            if proc is None:
                if w0 == "use":
                    i = 1
                    if toks[1][1] == ",":          # use, intrinsic :: name
                        i = [v for _, v in toks].index("::") + 1
                    mod.uses.append(toks[i][1])
                    continue
                if w0 == "implicit" or w0 == "save":
                    continue
                if w0 == "private" and len(toks) == 1:
                    mod.default_private = True
                    continue
                if w0 in ("private", "public") and (len(toks) == 1 or toks[1][1] == "::"):
                    continue
                if w0 == "public" and len(toks) > 1 and toks[1][0] == "id":
                    continue
                if w0 == "type" and toks[1][1] != "(":
                    name = [v for k, v in toks if k == "id"][-1]
                    dtype = DType(name, f"{mod.name}__{name}")
                    mod.types[name] = dtype
                    continue
                self.decl(toks, mod=mod)
                continue
            # inside a procedure: keep the raw statement, classify in pass 2
            if proc.is_program and w0 == "use":
                mod.uses.append(toks[1][1])
                continue
            proc.body.append((ln, toks))
        return self

    @staticmethod
    def _is_proc_header(toks):
        ids = [v for k, v in toks if k == "id"]
        if not ids or "::" in [v for _, v in toks]:
            return False
        if ids[0] in ("subroutine", "function"):
            return True
        return ids[0] in ("real", "integer", "logical", "type", "pure", "elemental", "recursive") \
            and "function" in ids and "=" not in [v for k, v in toks if k == "op"][:1]

    def proc_header(self, toks, mod, ln, parent=None):
        ids = [v for k, v in toks]
        for kw in ("subroutine", "function"):
            if kw in ids and toks[ids.index(kw)][0] == "id":
                j = ids.index(kw)
                # `integer(c_int) function f(x)` allowed; a declaration would have '::' before
                if "::" in ids[:j] or (j > 0 and ids[0] not in
                                       ("real", "integer", "logical", "type", "pure", "elemental", "recursive")):
                    return None
                name = ids[j + 1]
                args, result = [], name
                k = j + 2
                if k < len(ids) and ids[k] == "(":
                    k += 1
                    while ids[k] != ")":
                        if ids[k] != ",":
                            args.append(ids[k])
                        k += 1
                    k += 1
                bind_name = None
                while k < len(ids):
                    if ids[k] == "result":
                        result = ids[k + 2]
                        k += 4
                    elif ids[k] == "bind":
                        e = ids.index(")", k)
                        for t in range(k, e):
                            if toks[t][0] == "str":
                                bind_name = toks[t][1][1:-1]
                        k = e + 1
                    else:
                        k += 1
                p = Proc(name, kw, args, result if kw == "function" else None, mod, ln)
                if j > 0 and ids[0] in ("real", "integer", "logical", "type"):
                    self.scope_mod = mod
                    p.prefix_type = self.parse_typespec(Parser(toks[:j], f"{ln.file}:{ln.no}"))
                if bind_name:
                    p.cname = bind_name
                if parent is not None:
                    p.parent = parent
                    p.cname = f"{parent.cname}__{name}"
                    parent.children.append(p)
                else:
                    mod.procs[name] = p
                return p
        return None

    # ---------------------------------------------------------------- declarations
    def const_int(self, e, scope_syms):
        """Integer constant folding for kinds, lengths and extents."""
        if isinstance(e, Num):
            if re.fullmatch(r"\d+", e.text):
                return int(e.text)
            self.err(f"integer constant expected, got {e.text}")
        if isinstance(e, Paren):
            return self.const_int(e.e, scope_syms)
        if isinstance(e, Un) and e.op == "-":
            return -self.const_int(e.e, scope_syms)
        if isinstance(e, Bin) and e.op in "+-*/":
            a, b = self.const_int(e.l, scope_syms), self.const_int(e.r, scope_syms)
            return {"+": a + b, "-": a - b, "*": a * b, "/": int(a / b) if b else 0}[e.op]
        if isinstance(e, Kw):
            return self.const_int(e.e, scope_syms)
        if isinstance(e, Ref) and len(e.parts) == 1:
            name, args = e.parts[0]
            if args is None:
                s = self.lookup(name)
                if s is None or not s.param or s.init is None:
                    self.err(f"{name} is not a named integer constant")
                return self.const_int(s.init, scope_syms)
            if name == "kind":
                a = args[0]
                if isinstance(a, Log):
                    return 4
                if isinstance(a, Num):
                    return self.num_type(a.text).kind
            if name == "selected_int_kind":
                r = self.const_int(args[0], scope_syms)
                return 1 if r <= 2 else 2 if r <= 4 else 4 if r <= 9 else 8
            if name == "selected_real_kind":
                p = self.const_int(args[0], scope_syms)
                return 4 if p <= 6 else 8 if p <= 15 else 16
        self.err("unsupported constant expression")

    def parse_typespec(self, p):
        """p: Parser positioned at the type keyword.  Returns T (scalar)."""
        k, w = p.next()
        if w == "double":
            p.next()
            return T("real", 8)
        if w == "type":
            p.expect("(")
            name = p.next()[1]
            p.expect(")")
            if name == "c_ptr":
                return T("cptr", 8)
            return T("type", 0, 0, self.find_type(name).cname)
        base = {"real": "real", "integer": "int", "logical": "logical", "character": "char"}[w]
        kind, clen = 4, 1
        if p.accept("*"):
            kind = int(p.next()[1])
        elif p.peek() == ("op", "("):
            p.next()
            if p.peek()[0] == "id" and p.peek()[1] in ("kind", "len") and p.peek(1) == ("op", "="):
                p.next()
                p.next()
            if p.peek() == ("op", "*"):
                p.next()
                p.expect(")")
                return T("char", 1, 0, None, -1)      # assumed length: travels as an fstr_t value
            e = p.expr()
            p.expect(")")
            v = self.const_int(e, None)
            if base == "char":
                clen = v
            else:
                kind = v
        if base == "char":
            return T("char", 1, 0, None, clen)
        return T(base, kind)

    def decl(self, toks, dtype=None, mod=None, proc=None):
        p = Parser(toks, f"{self.cur_line.file}:{self.cur_line.no}")
        typ = self.parse_typespec(p)
        attrs = {"dims": None}
        while p.accept(","):
            a = p.next()[1]
            if a == "dimension":
                p.expect("(")
                attrs["dims"] = p.arglist()
            elif a == "intent":
                p.expect("(")
                attrs["intent"] = p.next()[1]
                if attrs["intent"] == "in" and p.peek()[1] == "out":
                    p.next()
                    attrs["intent"] = "inout"
                p.expect(")")
            elif a in ("parameter", "allocatable", "public", "private", "save", "target", "pointer", "value"):
                attrs[a] = True
            else:
                self.err(f"unsupported attribute {a}")
        p.accept("::")
        while True:
            k, name = p.next()
            if k != "id":
                self.err("entity name expected")
            dims = attrs["dims"]
            if p.peek() == ("op", "("):
                p.next()
                dims = p.arglist()
            init = None
            if p.accept("="):
                init = p.expr()
            t = T(typ.base, typ.kind, len(dims) if dims else 0, typ.tname, typ.clen)
            s = Sym(name, t, dims, attrs.get("allocatable", False) or attrs.get("pointer", False),
                    attrs.get("parameter", False), init, attrs.get("intent"), attrs.get("private", False))
            s.pointer, s.value = attrs.get("pointer", False), attrs.get("value", False)
            if dims and not s.alloc and any(isinstance(d, Colon) for d in dims):
                s.assumed = True          # assumed size / shape dummy: a bare pointer
            else:
                s.assumed = False
            if dtype is not None:
                s.cname = name
                dtype.fields.append(s)
            elif proc is not None:
                if name in proc.syms:
                    self.err(f"duplicate declaration of {name}")
                s.cname = name + "_"
                proc.syms[name] = s
            else:
                s.cname = f"{mod.name}__{name}"
                s.module = mod
                mod.syms[name] = s
            if not p.accept(","):
                break
        if not p.at_end():
            self.err(f"trailing tokens in declaration: {p.peek()[1]!r}")

    # ---------------------------------------------------------------- lookup
    def find_type(self, name, mod=None, seen=None):
        mod = mod or self.scope_mod
        seen = seen if seen is not None else set()
        if mod.name in seen:
            return None
        seen.add(mod.name)
        if name in mod.types:
            return mod.types[name]
        for u in mod.uses:
            if u in self.modules:
                r = self.find_type(name, self.modules[u], seen)
                if r:
                    return r
        if len(seen) == 1:
            self.err(f"unknown derived type {name}")
        return None

    def type_by_cname(self, cname):
        for m in self.order:
            for d in m.types.values():
                if d.cname == cname:
                    return d
        self.err(f"unknown struct {cname}")

    def lookup_mod(self, name, mod, seen, through_use):
        if mod.name in seen:
            return None
        seen.add(mod.name)
        s = mod.syms.get(name)
        if s is not None and not (through_use and s.private):
            return s
        for u in mod.uses:
            if u in self.modules:
                r = self.lookup_mod(name, self.modules[u], seen, True)
                if r is not None:
                    return r
        return None

    def lookup(self, name):
        pr = self.scope_proc
        while pr is not None:
            if name in pr.syms:
                return pr.syms[name]
            pr = pr.parent
        return self.lookup_mod(name, self.scope_mod, set(), False)

    def lookup_proc(self, name, mod=None, seen=None):
        if mod is None:
            pr = self.scope_proc
            while pr is not None:
                for ch in pr.children:
                    if ch.name == name:
                        return ch
                pr = pr.parent
        mod = mod or self.scope_mod
        seen = seen if seen is not None else set()
        if mod.name in seen:
            return None
        seen.add(mod.name)
        if name in mod.procs:
            return mod.procs[name]
        for u in mod.uses:
            if u in self.modules:
                r = self.lookup_proc(name, self.modules[u], seen)
                if r:
                    return r
        return None

    # ---------------------------------------------------------------- expression typing / emission
    def num_type(self, text):
        m = re.fullmatch(r"([\d.]+)(?:([ed])([+-]?\d+))?(?:_(\w+))?", text)
        if not m:
            self.err(f"bad numeric literal {text}")
        mant, ech, ex, ksuf = m.groups()
        is_real = "." in mant or ech is not None
        kind = 4
        if ech == "d":
            kind = 8
        if ksuf:
            kind = int(ksuf) if ksuf.isdigit() else self.const_int(Ref([(ksuf, None)]), None)
        return T("real" if is_real else "int", kind)

    def num_c(self, text):
        m = re.fullmatch(r"([\d.]+)(?:([ed])([+-]?\d+))?(?:_(\w+))?", text)
        mant, ech, ex, ksuf = m.groups()
        t = self.num_type(text)
        if t.base == "int":
            return (mant + "LL") if t.kind == 8 else mant, t
        if "." not in mant:
            mant += "."
        s = mant + (("e" + ex) if ex is not None else "")
        if s.startswith("."):
            s = "0" + s
        if s.endswith(".") or ".e" in s:
            s = s.replace(".e", ".0e") if ".e" in s else s + "0"
        return (s + "f") if t.kind == 4 else s, t

    @staticmethod
    def promote(a, b):
        order = {"int": 0, "real": 1}
        if a.base not in order or b.base not in order:
            return a
        if a.base == b.base:
            return T(a.base, max(a.kind, b.kind))
        r = a if a.base == "real" else b
        return T("real", r.kind)

    def sym_total_size(self, s, obj):
        """C expression: number of elements of array object `obj` (C lvalue / name) of Sym s."""
        if s.alloc:
            return " * ".join(f"({obj}).n{d + 1}" for d in range(s.typ.rank))
        return " * ".join(f"({self.const_int(d, None)})" for d in s.dims)

    def sym_extent(self, s, obj, d):
        if s.alloc:
            return f"({obj}).n{d + 1}"
        return str(self.const_int(s.dims[d], None))

    def resolve_ref(self, ref, ivar):
        """Walk a designator.  Returns (cexpr, T, is_lvalue).  `ivar`: loop index (0-based C
        expression) that replaces whole-array references and `:` subscripts, or None."""
        name, args = ref.parts[0]
        s = self.lookup(name)
        if s is None:
            return None
        if s.param and not s.local_const and s.typ.rank == 0 and s.module is not None:
            obj, typ, lval = s.cname, s.typ, False
        else:
            obj = s.cname
            if s.dummy and (s.typ.rank == 0 or s.alloc) and not s.value and s.typ.base != "char":
                obj = f"(*{s.cname})"
            typ, lval = s.typ, not s.param
        cur = s
        parts = ref.parts
        for pi, (pname, pargs) in enumerate(parts):
            if pi > 0:
                d = self.type_by_cname(typ.tname)
                f = next((x for x in d.fields if x.name == pname), None)
                if f is None:
                    self.err(f"type {d.name} has no component {pname}")
                obj = f"{obj}.{f.cname}"
                cur, typ = f, f.typ
            self.last_sym = cur
            if pargs is not None:
                if typ.rank == 0:
                    self.err(f"{pname} is not an array")
                if len(pargs) != typ.rank:
                    self.err(f"rank mismatch on {pname}")
                idx, stride = [], "1"
                for dnum, a in enumerate(pargs):
                    if isinstance(a, Colon):
                        if a.lo is not None or a.hi is not None:
                            self.err("only full-range (:) sections are supported")
                        if ivar is None:
                            self.err("array section in scalar context")
                        sub = f"({ivar})"
                    else:
                        ce, ct = self.emit(a, ivar)
                        sub = f"(({ce}) - 1)"
                    idx.append(sub if stride == "1" else f"{stride} * {sub}")
                    stride = (f"{stride} * {self.sym_extent(cur, obj, dnum)}" if stride != "1"
                              else self.sym_extent(cur, obj, dnum))
                flat = " + ".join(idx)
                obj = self.elem(cur, obj, flat)
                typ = typ.scalar()
            elif typ.rank > 0 and (ivar is not None) and not self.want_whole:
                obj = self.elem(cur, obj, f"({ivar})")
                typ = typ.scalar()
        return obj, typ, lval

    def elem(self, s, obj, flat):
        et = s.typ.scalar()
        if s.alloc:
            if et.base == "char":
                return f"((char (*)[{et.clen}])({obj}).p)[{flat}]"
            return f"(({et.ctype()} *)({obj}).p)[{flat}]"
        return f"{obj}[{flat}]"

    want_whole = False

    def extent_of(self, e):
        """C expression for the element count if `e` is array-valued, else None."""
        if isinstance(e, (Num, Str, Log)):
            return None
        if isinstance(e, Paren):
            return self.extent_of(e.e)
        if isinstance(e, Un):
            return self.extent_of(e.e)
        if isinstance(e, Bin):
            return self.extent_of(e.l) or self.extent_of(e.r)
        if isinstance(e, ArrCons):
            return str(len(e.items))
        if isinstance(e, Kw):
            return self.extent_of(e.e)
        if isinstance(e, Ref):
            name, args = e.parts[0]
            s = self.lookup(name)
            if s is None:
                elemental = name in INTRINSIC_ELEMENTAL or name in ("abs", "max", "min", "merge", "real",
                                                                   "int", "mod")
                if not elemental:
                    return None          # inquiry / transformational intrinsics, user functions
                if args is not None:
                    for a in args:
                        x = self.extent_of(a)
                        if x:
                            return x
                return None
            obj = f"(*{s.cname})" if (s.dummy and (s.typ.rank == 0 or s.alloc) and not s.value) else s.cname
            cur, typ = s, s.typ
            for pi, (pname, pargs) in enumerate(e.parts):
                if pi > 0:
                    d = self.type_by_cname(typ.tname)
                    f = next((x for x in d.fields if x.name == pname), None)
                    if f is None:
                        self.err(f"type {d.name} has no component {pname}")
                    obj = f"{obj}.{f.cname}"
                    cur, typ = f, f.typ
                if pargs is None:
                    if typ.rank > 0:
                        return self.sym_total_size(cur, obj)
                else:
                    for dnum, a in enumerate(pargs):
                        if isinstance(a, Colon):
                            return self.sym_extent(cur, obj, dnum)
                        x = self.extent_of(a)      # vector subscript
                        if x is not None:
                            return x
                    # element selected: continue into components with a scalar object
                    obj = self.elem(cur, obj, "0")
                    typ = typ.scalar()
            return None
        return None

    def newtmp(self, base="_t"):
        self.tmp += 1
        return f"{base}{self.tmp}"

    def emit(self, e, ivar=None):
        """Returns (C expression string, T)."""
        if isinstance(e, Num):
            return self.num_c(e.text)
        if isinstance(e, Log):
            return ("1" if e.val else "0"), T("logical")
        if isinstance(e, Str):
            esc = e.val.replace("\\", "\\\\").replace('"', '\\"')
            return f'f_lit("{esc}", {len(e.val)})', T("char", 1, 0, None, 0)
        if isinstance(e, Paren):
            c, t = self.emit(e.e, ivar)
            return f"({c})", t
        if isinstance(e, Kw):
            return self.emit(e.e, ivar)
        if isinstance(e, Un):
            c, t = self.emit(e.e, ivar)
            if e.op == ".not.":
                return f"(!({c}))", T("logical")
            return f"(-({c}))", t
        if isinstance(e, Bin):
            lc, lt = self.emit(e.l, ivar)
            rc, rt = self.emit(e.r, ivar)
            if e.op in (".and.", ".or."):
                # Fortran does not short-circuit by rule, but every operand here is side-effect free
                return f"(({lc}) {'&&' if e.op == '.and.' else '||'} ({rc}))", T("logical")
            if e.op in (".eqv.", ".neqv."):
                return f"((!!({lc})) {'==' if e.op == '.eqv.' else '!='} (!!({rc})))", T("logical")
            if e.op in ("==", "!=", "<", "<=", ">", ">="):
                if lt.base == "char" or rt.base == "char":
                    return f"(f_cmp({self.as_fstr(lc, lt)}, {self.as_fstr(rc, rt)}) {e.op} 0)", T("logical")
                return f"(({lc}) {e.op} ({rc}))", T("logical")
            if e.op == "//":
                lc = self.as_fstr(lc, lt)
                rc = self.as_fstr(rc, rt)
                return f"f_cat({lc}, {rc})", T("char", 1, 0, None, 0)
            if e.op == "**":
                if rt.base == "int":
                    if lt.base == "int":
                        return f"(({lt.ctype()})f_ipow({lc}, {rc}))", lt
                    fn = "__builtin_powi" if lt.kind == 8 else "__builtin_powif"
                    return f"{fn}({lc}, {rc})", lt
                pt = self.promote(lt, rt)
                fn = "pow" if pt.kind == 8 else "powf"
                return f"{fn}({lc}, {rc})", pt
            pt = self.promote(lt, rt)
            return f"(({lc}) {e.op} ({rc}))", pt
        if isinstance(e, ArrCons):
            if ivar is None:
                self.err("array constructor in scalar context")
            items = [self.emit(x, None) for x in e.items]
            t = items[0][1]
            for _, it in items[1:]:
                t = self.promote(t, it)
            arr = ", ".join(c for c, _ in items)
            return f"(({t.ctype()}[]){{{arr}}})[{ivar}]", t
        if isinstance(e, Ref):
            return self.emit_ref(e, ivar)
        self.err("unsupported expression node")

    def as_fstr(self, c, t):
        if t.clen is None or t.clen <= 0 or c.startswith(("f_lit(", "f_cat(", "f_trim(", "f_var(")):
            return c
        return f"f_var({c}, {t.clen})"

    def emit_ref(self, e, ivar):
        name, args = e.parts[0]
        r = self.resolve_ref(e, ivar)
        if r is not None:
            c, t, _ = r
            if t.rank > 0:
                self.err(f"whole array {name} in scalar context")
            return c, t
        if len(e.parts) != 1 or args is None:
            self.err(f"unknown name {name}")
        # user function
        pr = self.lookup_proc(name)
        if pr is not None:
            if pr.kind != "function":
                self.err(f"{name} is a subroutine")
            return f"{pr.cname}({self.call_args(pr, args, ivar)})", pr.rtype
        # intrinsics
        if name in INTRINSIC_ELEMENTAL:
            c, t = self.emit(args[0], ivar)
            if t.base != "real":
                self.err(f"{name} of a non-real")
            return f"{INTRINSIC_ELEMENTAL[name]}{'f' if t.kind == 4 else ''}({c})", t
        if name == "abs":
            c, t = self.emit(args[0], ivar)
            if t.base == "real":
                return f"{'fabs' if t.kind == 8 else 'fabsf'}({c})", t
            return f"{'llabs' if t.kind == 8 else 'abs'}({c})", t
        if name in ("max", "min"):
            parts = [self.emit(a, ivar) for a in args]
            t = parts[0][1]
            for _, pt in parts[1:]:
                t = self.promote(t, pt)
            mac = "F_MAX2" if name == "max" else "F_MIN2"
            c = parts[0][0]
            for pc, _ in parts[1:]:
                c = f"{mac}({t.ctype()}, {c}, {pc})"
            return c, t
        if name == "merge":
            tc, tt = self.emit(args[0], ivar)
            fc, ft = self.emit(args[1], ivar)
            mc, _ = self.emit(args[2], ivar)
            t = tt
            if t.base == "char":
                return f"(({mc}) ? {self.as_fstr(tc, tt)} : {self.as_fstr(fc, ft)})", T("char", 1, 0, None, 0)
            return f"(({mc}) ? ({t.ctype()})({tc}) : ({t.ctype()})({fc}))", t
        if name == "sum":
            arr = args[0]
            n = self.extent_of(arr)
            if n is None:
                self.err("sum of a scalar")
            iv = self.newtmp("_i")
            c, t = self.emit(arr, iv)
            acc = self.newtmp("_s")
            zero = "0"
            return (f"({{ {t.ctype()} {acc} = {zero}; for (int {iv} = 0; {iv} < ({n}); {iv}++) "
                    f"{acc} += ({c}); {acc}; }})"), t
        if name == "size":
            if len(args) == 2:
                self.want_whole = True
                c, t, _ = self.resolve_ref(args[0], None)
                self.want_whole = False
                sym = self.last_sym
                return f"((int)({self.sym_extent(sym, c, self.const_int(args[1], None) - 1)}))", T("int", 4)
            self.want_whole = True
            n = self.extent_of(args[0])
            self.want_whole = False
            if n is None:
                self.err("size of a scalar")
            return f"((int)({n}))", T("int", 4)
        if name in ("allocated", "associated"):
            self.want_whole = True
            c, t, _ = self.resolve_ref(args[0], None)
            self.want_whole = False
            return f"(({c}).p != 0)", T("logical")
        if name == "c_associated":
            c, t = self.emit(args[0], ivar)
            return f"(({c}) != 0)", T("logical")
        if name == "c_loc":
            self.want_whole = True
            c, t, _ = self.resolve_ref(args[0], None)
            self.want_whole = False
            sym = self.last_sym
            if t.rank > 0 and sym.alloc:
                return f"((void *)({c}).p)", T("cptr", 8)
            if t.rank > 0:
                return f"((void *)({c}))", T("cptr", 8)
            return f"((void *)&({c}))", T("cptr", 8)
        if name == "trim":
            c, t = self.emit(args[0], ivar)
            return f"f_trim({self.as_fstr(c, t)})", T("char", 1, 0, None, 0)
        if name == "real":
            c, t = self.emit(args[0], ivar)
            kind = self.const_int(args[1], None) if len(args) > 1 else 4
            rt = T("real", kind)
            return f"(({rt.ctype()})({c}))", rt
        if name == "int":
            c, t = self.emit(args[0], ivar)
            rt_ = T("int", self.const_int(args[1], None) if len(args) > 1 else 4)
            return f"(({rt_.ctype()})({c}))", rt_
        if name == "mod":
            ac, at = self.emit(args[0], ivar)
            bc, bt = self.emit(args[1], ivar)
            pt = self.promote(at, bt)
            if pt.base == "int":
                return f"(({ac}) % ({bc}))", pt
            return f"fmod({ac}, {bc})", pt
        self.err(f"unknown function or intrinsic {name}")

    def call_args(self, pr, args, ivar=None):
        if len(args) != len(pr.args):
            self.err(f"{pr.name}: {len(args)} actual arguments for {len(pr.args)} dummies")
        out = []
        ordered = [None] * len(pr.args)
        for pos, a in enumerate(args):
            if isinstance(a, Kw):
                if a.name not in pr.args:
                    self.err(f"{pr.name} has no dummy argument {a.name}")
                pos, a = pr.args.index(a.name), a.e
            if ordered[pos] is not None:
                self.err("argument given twice")
            ordered[pos] = a
        for a, dn in zip(ordered, pr.args):
            ds = pr.syms[dn]
            if ds.typ.base == "char" and ds.typ.rank == 0:
                c, t = self.emit(a, ivar)
                out.append(self.as_fstr(c, t))
                continue
            if ds.value:
                c, t = self.emit(a, ivar)
                out.append(f"({ds.typ.ctype()})({c})")
                continue
            if ds.typ.rank > 0:
                # whole-array actual
                if not isinstance(a, Ref):
                    self.err("array actual argument must be a name")
                self.want_whole = True
                r = self.resolve_ref(a, None)
                self.want_whole = False
                c, t, _ = r
                out.append(c if not ds.alloc else f"&({c})")
                continue
            lv = None
            if isinstance(a, Ref):
                r = self.resolve_ref(a, ivar)
                if r is not None and r[2] and r[1].rank == 0:
                    lv = r
            if lv is not None and (lv[1].base in ("cptr",) and ds.typ.base == "cptr" or
                                   (lv[1].base, lv[1].kind, lv[1].tname) ==
                                   (ds.typ.base, ds.typ.kind, ds.typ.tname)):
                out.append(f"&({lv[0]})")
            else:
                if ds.typ.base == "type":
                    self.err("derived-type actual must be a variable")
                if ds.intent in ("out", "inout"):
                    self.err(f"non-variable actual for intent({ds.intent}) dummy {dn}")
                c, t = self.emit(a, ivar)
                out.append(f"&({ds.typ.ctype()}){{{c}}}")
        return ", ".join(out)

    # ---------------------------------------------------------------- statements
    def w(self, s):
        self.out.append("  " * self.ind + s)

    def stmt(self, toks):
        label = None
        if (len(toks) > 2 and toks[0][0] == "id" and toks[1] == ("op", ":")
                and toks[2][0] == "id" and toks[2][1] in ("do", "if", "select")):
            label, toks = toks[0][1], toks[2:]
        if (toks[0][0] == "id" and toks[0][1] in ("enddo", "endif", "endselect") and len(toks) == 2
                and toks[1][0] == "id"):
            toks = toks[:1]
        if (toks[0] == ("id", "end") and len(toks) == 3 and toks[2][0] == "id"
                and toks[1][1] in ("do", "if", "select")):
            toks = toks[:2]
        p = Parser(toks, f"{self.cur_line.file}:{self.cur_line.no}")
        k0, w0 = toks[0]
        w1 = toks[1][1] if len(toks) > 1 else None
        # --- block ends
        if k0 == "id" and w0 in ("end", "enddo", "endif", "endselect"):
            what = w1 if w0 == "end" else w0[3:]
            if what == "select":
                kind = self.blocks.pop()
                if kind == "case":
                    self.ind -= 1
                    self.w("}")
                else:
                    assert kind == "select0"
                self.selvar.pop()
                self.ind -= 1
                self.w("}")
                return
            kind = self.blocks.pop()
            if (what, kind) not in (("do", "do"), ("if", "if")):
                self.err(f"END {what} closes a {kind} block")
            if kind == "do":
                lab = self.loopnames.pop()
                if lab and lab[1]:
                    self.w(f"_cyc_{lab[0]}: ;")
            self.ind -= 1
            self.w("}")
            if kind == "do":
                self.ind -= 1
                self.w("}")
                if lab and lab[2]:
                    self.w(f"_brk_{lab[0]}: ;")
            return
        if k0 == "id" and w0 == "else" and (len(toks) == 1):
            self.ind -= 1
            self.w("} else {")
            self.ind += 1
            return
        if k0 == "id" and (w0 == "elseif" or (w0 == "else" and w1 == "if")):
            p.next()
            if w0 == "else":
                p.next()
            p.expect("(")
            c, _ = self.emit(p.expr())
            p.expect(")")
            self.ind -= 1
            self.w(f"}} else if ({c}) {{")
            self.ind += 1
            return
        if k0 == "id" and w0 == "if" and w1 == "(":
            p.next()
            p.next()
            cond = p.expr()
            p.expect(")")
            c, _ = self.emit(cond)
            if p.peek() == ("id", "then") and p.i == len(toks) - 1:
                self.w(f"if ({c}) {{")
                self.ind += 1
                self.blocks.append("if")
                return
            self.w(f"if ({c}) {{")
            self.ind += 1
            self.stmt(toks[p.i:])
            self.ind -= 1
            self.w("}")
            return
        if k0 == "id" and w0 == "do" and (len(toks) == 1 or (toks[1][0] == "id" and w1 != "while"
                                                            and len(toks) > 2 and toks[2][1] == "=")):
            self.w("{")
            self.ind += 1
            if len(toks) == 1:
                self.w("for (;;) {")
            else:
                p.next()
                var = p.next()[1]
                p.expect("=")
                lo = p.expr()
                p.expect(",")
                hi = p.expr()
                step = None
                if p.accept(","):
                    step = p.expr()
                vs = self.lookup(var)
                vc = f"(*{vs.cname})" if vs.dummy else vs.cname
                loc, _ = self.emit(lo)
                hic, _ = self.emit(hi)
                e_ = self.newtmp("_e")
                self.w(f"const int {e_} = {hic};")
                if step is None:
                    self.w(f"for ({vc} = {loc}; {vc} <= {e_}; {vc}++) {{")
                else:
                    sc = self.const_int(step, None)
                    cmp_ = "<=" if sc > 0 else ">="
                    self.w(f"for ({vc} = {loc}; {vc} {cmp_} {e_}; {vc} += ({sc})) {{")
            self.ind += 1
            self.blocks.append("do")
            self.tmp += 1
            self.loopnames.append([f"{label}_{self.tmp}", False, False, label] if label else None)
            return
        if k0 == "id" and w0 == "do" and w1 == "while":
            p.next()
            p.next()
            p.expect("(")
            c, _ = self.emit(p.expr())
            p.expect(")")
            self.w("{")
            self.ind += 1
            self.w(f"while ({c}) {{")
            self.ind += 1
            self.blocks.append("do")
            self.loopnames.append(None)
            return
        if k0 == "id" and (w0 == "stop" or (w0 == "error" and w1 == "stop")):
            msg = next((v[1:-1] for k, v in toks if k == "str"), "STOP")
            esc = msg.replace("\\", "\\\\").replace('"', '\\"')
            self.w(f'fputs("{esc}\\n", stderr); abort();')
            return
        if k0 == "id" and w0 in ("read", "write") and w1 == "(" and self._ctl_items(toks) == 1:
            # unformatted stream transfer: read(u) a, b, ...   /   write(u) a, b, ...
            ip = Parser(toks[2:], p.where)
            unit, _ = self.emit(ip.arglist()[0])
            while not ip.at_end():
                a = ip.expr()
                if not isinstance(a, Ref):
                    self.err("I/O list items must be variables")
                self.want_whole = True
                c, t, _ = self.resolve_ref(a, None)
                self.want_whole = False
                sym = self.last_sym
                esz = f"sizeof({t.scalar().ctype()})" + (f" * {t.clen}" if t.base == "char" else "")
                if t.rank > 0:
                    ptr = f"({c}).p" if sym.alloc else f"({c})"
                    n = self.sym_total_size(sym, c)
                    self.w(f"f_io({unit}, {ptr}, {esz} * (size_t)({n}), {int(w0 == 'write')});")
                else:
                    self.w(f"f_io({unit}, &({c}), {esz}, {int(w0 == 'write')});")
                ip.accept(",")
            return
        if k0 == "id" and w0 == "open" and w1 == "(":
            items = Parser(toks[2:], p.where).arglist()
            unit, _ = self.emit(items[0])
            kw = {x.name: x.e for x in items if isinstance(x, Kw)}
            fc, ft = self.emit(kw["file"])
            old = isinstance(kw.get("status"), Str) and kw["status"].val.lower() == "old"
            self.w(f"f_open({unit}, {self.as_fstr(fc, ft)}, {int(old)});")
            return
        if k0 == "id" and w0 == "close" and w1 == "(":
            unit, _ = self.emit(Parser(toks[2:], p.where).arglist()[0])
            self.w(f"fclose(f_units[{unit}]); f_units[{unit}] = 0;")
            return
        if k0 == "id" and w0 == "call" and w1 == "get_command_argument":
            items = Parser(toks[3:], p.where).arglist()
            ic, _ = self.emit(items[0])
            vc, vt, _ = self.resolve_ref(items[1], None)
            self.w(f"f_assign({vc}, {vt.clen}, ({ic}) < f_argc ? f_lit(f_argv[{ic}], (int)strlen(f_argv[{ic}])) "
                   f": f_lit(\"\", 0));")
            return
        if k0 == "id" and w0 == "write" and w1 == "(":
            # unit and format are ignored: every item goes to stderr in list order (error paths only)
            depth, j = 0, 1
            while True:
                if toks[j][1] == "(" and toks[j][0] == "op":
                    depth += 1
                elif toks[j][1] == ")" and toks[j][0] == "op":
                    depth -= 1
                    if depth == 0:
                        break
                j += 1
            ip = Parser(toks[j + 1:], p.where)
            while not ip.at_end():
                a = ip.arg()
                if isinstance(a, Ref) and a.parts[-1][1] and isinstance(a.parts[-1][1][0], Colon) \
                        and a.parts[-1][1][0].hi is not None:
                    sec = a.parts[-1][1][0]
                    base = Ref(a.parts[:-1] + [(a.parts[-1][0], [sec.lo or Num("1")])])
                    bc, bt, _ = self.resolve_ref(base, None)
                    hi, _ = self.emit(sec.hi)
                    lo, _ = self.emit(sec.lo or Num("1"))
                    self.w(f"fwrite((const char *)({bc}), 1, (size_t)(({hi}) - ({lo}) + 1), stderr);")
                else:
                    c, t = self.emit(a)
                    if t.base == "char":
                        tv = self.newtmp("_w")
                        self.w(f"{{ fstr_t {tv} = {self.as_fstr(c, t)}; fwrite({tv}.s, 1, {tv}.n, stderr); }}")
                    elif t.base == "real":
                        self.w(f'fprintf(stderr, " %.17g", (double)({c}));')
                    else:
                        self.w(f'fprintf(stderr, " %lld", (long long)({c}));')
                ip.accept(",")
            self.w('fputc(10, stderr);')
            return
        if k0 == "id" and w0 == "read" and w1 == "(":
            # only: read(<character variable>, *, iostat=<int>) <integer variable>
            ip = Parser(toks[2:], p.where)
            items = ip.arglist()
            src, st_ = items[0], next((x.e for x in items if isinstance(x, Kw) and x.name == "iostat"), None)
            tgt = ip.expr()
            sc, stt = self.emit(src)
            tc, tt, _ = self.resolve_ref(tgt, None)
            if stt.base != "char" or tt.base != "int":
                self.err("unsupported READ")
            tv = self.newtmp("_r")
            self.w(f"{{ fstr_t {tv} = {self.as_fstr(sc, stt)}; {tv}.s[{tv}.n] = 0; long long _v; "
                   f"int _ok = sscanf({tv}.s, \"%lld\", &_v) == 1; if (_ok) {tc} = _v;")
            if st_ is not None:
                self.w(f"  {self.resolve_ref(st_, None)[0]} = _ok ? 0 : 1;")
            self.w("}")
            return
        if k0 == "id" and w0 == "call" and w1 == "get_environment_variable":
            ip = Parser(toks[3:], p.where)
            items = ip.arglist()
            nm, _t = self.emit(items[0])
            val = self.resolve_ref(items[1], None)
            st_ = next((x.e for x in items if isinstance(x, Kw) and x.name == "status"), None)
            tv = self.newtmp("_g")
            self.w(f"{{ fstr_t {tv} = {self.as_fstr(nm, _t)}; {tv}.s[{tv}.n] = 0; const char *_e = getenv({tv}.s);")
            self.w(f"  f_assign({val[0]}, {val[1].clen}, _e ? f_lit(_e, (int)strlen(_e)) : f_lit(\"\", 0));")
            if st_ is not None:
                self.w(f"  {self.resolve_ref(st_, None)[0]} = _e ? 0 : 1;")
            self.w("}")
            return
        if k0 == "id" and w0 == "call" and w1 == "c_f_pointer":
            ip = Parser(toks[3:], p.where)
            items = ip.arglist()
            pc, _t = self.emit(items[0])
            self.want_whole = True
            fc, ft, _ = self.resolve_ref(items[1], None)
            self.want_whole = False
            self.w(f"({fc}).p = {pc};")
            if len(items) > 2 and isinstance(items[2], ArrCons):
                for dnum, d in enumerate(items[2].items):
                    self.w(f"({fc}).n{dnum + 1} = {self.emit(d)[0]};")
            return
        if k0 == "id" and w0 == "select" and w1 == "case":
            p.next()
            p.next()
            p.expect("(")
            c, t = self.emit(p.expr())
            p.expect(")")
            sv = self.newtmp("_sel")
            self.w("{")
            self.ind += 1
            self.w(f"const {t.ctype()} {sv} = {c};")
            self.blocks.append("select0")
            self.selvar.append(sv)
            return
        if k0 == "id" and w0 == "case":
            first = self.blocks[-1] == "select0"
            sv = self.selvar[-1]
            if w1 == "default":
                head = "{" if first else "} else {"
            else:
                p.next()
                p.expect("(")
                conds = []
                while True:
                    a = p.arg()
                    if isinstance(a, Colon):
                        cc = []
                        if a.lo is not None:
                            cc.append(f"{sv} >= ({self.emit(a.lo)[0]})")
                        if a.hi is not None:
                            cc.append(f"{sv} <= ({self.emit(a.hi)[0]})")
                        conds.append("(" + " && ".join(cc) + ")")
                    else:
                        conds.append(f"({sv} == ({self.emit(a)[0]}))")
                    if not p.accept(","):
                        break
                p.expect(")")
                head = ("if (" if first else "} else if (") + " || ".join(conds) + ") {"
            if first:
                self.blocks[-1] = "case"
            else:
                self.ind -= 1
            self.w(head)
            self.ind += 1
            return
        if k0 == "id" and w0 in ("exit", "cycle") and len(toks) == 1:
            self.w("break;" if w0 == "exit" else "continue;")
            return
        if k0 == "id" and w0 in ("exit", "cycle") and len(toks) == 2 and toks[1][0] == "id":
            lab = next((x for x in reversed(self.loopnames) if x and x[3] == w1), None)
            if lab is None:
                self.err(f"no enclosing loop named {w1}")
            if w0 == "cycle":
                lab[1] = True
                self.w(f"goto _cyc_{lab[0]};")
            else:
                lab[2] = True
                self.w(f"goto _brk_{lab[0]};")
            return
        if k0 == "id" and w0 == "return" and len(toks) == 1:
            self.w(f"goto {self.retlabel};")
            self.uses_ret = True
            return
        if k0 == "id" and w0 == "call":
            p.next()
            name = p.next()[1]
            args = []
            if p.accept("("):
                args = p.arglist()
            pr = self.lookup_proc(name)
            if pr is None:
                self.err(f"unknown subroutine {name}")
            self.w(f"{pr.cname}({self.call_args(pr, args)});")
            return
        if k0 == "id" and w0 in ("allocate", "deallocate") and w1 == "(":
            p.next()
            p.next()
            items = p.arglist()
            for it in items:
                if not isinstance(it, Ref):
                    self.err("bad allocate item")
                if w0 == "allocate":
                    shape = it.parts[-1][1]
                    base = Ref(it.parts[:-1] + [(it.parts[-1][0], None)])
                    self.want_whole = True
                    c, t, _ = self.resolve_ref(base, None)
                    self.want_whole = False
                    dims = [self.emit(d)[0] for d in shape]
                    et = t.scalar()
                    esz = f"sizeof({et.ctype()})" + (f" * {et.clen}" if et.base == "char" else "")
                    for dnum, d in enumerate(dims):
                        self.w(f"({c}).n{dnum + 1} = {d};")
                    prod = " * ".join(f"(size_t)({c}).n{dnum + 1}" for dnum in range(len(dims)))
                    self.w(f"({c}).p = f_alloc({esz} * {prod});")
                else:
                    self.want_whole = True
                    c, t, _ = self.resolve_ref(it, None)
                    self.want_whole = False
                    self.w(f"free(({c}).p); ({c}).p = 0;")
            return
        # --- assignment
        depth, eq = 0, None
        for i, (k, v) in enumerate(toks):
            if k == "op" and v in ("(", "(/"):
                depth += 1
            elif k == "op" and v in (")", "/)"):
                depth -= 1
            elif k == "op" and v == "=" and depth == 0:
                eq = i
                break
        if eq is None:
            self.err("unrecognised statement")
        lp = Parser(toks[:eq], p.where)
        lhs = lp.expr()
        if not lp.at_end():
            self.err("unsupported statement")
        rp = Parser(toks[eq + 1:], p.where)
        rhs = rp.expr()
        if not rp.at_end():
            self.err("trailing tokens after expression")
        if not isinstance(lhs, Ref):
            self.err("bad assignment target")
        if self.lookup(lhs.parts[0][0]) is None:
            self.err(f"assignment to an undeclared name {lhs.parts[0][0]}")
        n = self.extent_of(lhs)
        if n is not None:
            iv = self.newtmp("_i")
            lc, lt, lv = self.resolve_ref(lhs, iv)
            rc, rt = self.emit(rhs, iv)
            if lt.base == "char":
                self.w(f"for (int {iv} = 0; {iv} < ({n}); {iv}++) "
                       f"f_assign({lc}, {lt.clen}, {self.as_fstr(rc, rt)});")
            else:
                self.w(f"for (int {iv} = 0; {iv} < ({n}); {iv}++) {lc} = {rc};")
            return
        lc, lt, lv = self.resolve_ref(lhs, None)
        if not lv:
            self.err("assignment to a constant")
        if self.extent_of(rhs) is not None:
            self.err("array expression assigned to a scalar")
        rc, rt = self.emit(rhs, None)
        if lt.base == "char":
            self.w(f"f_assign({lc}, {lt.clen}, {self.as_fstr(rc, rt)});")
        else:
            self.w(f"{lc} = {rc};")

    @staticmethod
    def _ctl_items(toks):
        """number of items in the parenthesised control list that follows READ / WRITE"""
        depth, n = 0, 1
        for k, v in toks[1:]:
            if k == "op" and v == "(":
                depth += 1
            elif k == "op" and v == ")":
                depth -= 1
                if depth == 0:
                    return n
            elif k == "op" and v == "," and depth == 1:
                n += 1
        return n

    # ---------------------------------------------------------------- driver
    def cdecl(self, s, name=None, static_tls=False):
        """C declaration of a variable / component."""
        name = name or s.cname
        t = s.typ
        pre = "REF_TLS " if static_tls else ""
        if s.alloc:
            return f"{pre}fa_t {name};"
        dim = ""
        if t.rank > 0:
            n = 1
            for d in s.dims:
                n *= self.const_int(d, None)
            dim = f"[{n}]"
        if t.base == "char":
            dim += f"[{t.clen}]"
        return f"{pre}{t.ctype()} {name}{dim};"

    def has_default_init(self, t):
        return t.base == "type" and any(f.init is not None for f in self.type_by_cname(t.tname).fields)

    def proto(self, pr):
        args = []
        for a in pr.args:
            s = pr.syms[a]
            if s.alloc:
                args.append(f"fa_t *{s.cname}")
            elif s.typ.base == "char" and s.typ.rank == 0:
                args.append(f"fstr_t {s.cname}")       # character(len=*) dummy: by value
            elif s.value:
                args.append(f"{s.typ.ctype()} {s.cname}")
            else:
                args.append(f"{s.typ.ctype()} *{s.cname}")
        r = pr.rtype.ctype() if pr.kind == "function" else "void"
        return f"{r} {pr.cname}({', '.join(args) if args else 'void'})"

    def prepare_proc(self, pr):
        """Split declarations from executable statements; fill the symbol table."""
        self.scope_mod, self.scope_proc = pr.module, pr
        body = []
        decl_heads = ("real", "integer", "logical", "character", "double")
        for ln, toks in pr.body:
            self.cur_line = ln
            w0 = toks[0][1]
            if toks[0][0] == "id":
                if w0 in ("implicit", "use", "save", "import"):
                    if w0 == "use":
                        self.err("procedure-level USE is not supported")
                    continue
                is_decl = (w0 in decl_heads) or (w0 == "type" and toks[1][1] == "(")
                if is_decl:
                    self.decl(toks, proc=pr)
                    continue
            body.append((ln, toks))
        pr.body = body
        self.cur_line = pr.line
        for a in pr.args:
            if a not in pr.syms:
                self.err(f"dummy argument {a} is not declared")
            pr.syms[a].dummy = True
        if pr.kind == "function":
            if pr.result not in pr.syms:
                if pr.prefix_type is None:
                    self.err("function result is not declared")
                rs = Sym(pr.result, pr.prefix_type)
                rs.cname = pr.result + "_"
                pr.syms[pr.result] = rs
            pr.syms[pr.result].is_result = True
            pr.rtype = pr.syms[pr.result].typ
        for s in pr.syms.values():
            if s.param:
                s.local_const = True
        for ch in pr.children:
            self.prepare_proc(ch)

    def emit_proc(self, pr, nested=False):
        saved = None
        if nested:      # an internal procedure becomes a GCC nested function of its host
            saved = (self.scope_proc, self.blocks, self.selvar, self.ind, self.uses_ret, self.loopnames)
        base_ind = self.ind if nested else 0
        self.scope_mod, self.scope_proc = pr.module, pr
        self.blocks, self.selvar, self.uses_ret, self.loopnames = [], [], False, []
        self.ind = base_ind
        self.cur_line = pr.line
        self.w(f"/* {pr.line.file}:{pr.line.no} {pr.kind} {pr.name} */")
        self.w(("auto " if nested else "") + self.proto(pr) + " {")
        self.ind = base_ind + 1
        pid = len(self.prof_names)
        self.prof_names.append(pr.cname)
        self.w(f"PROF_ENTER({pid})")
        allocs = []
        for s in pr.syms.values():
            if s.dummy:
                continue
            if s.param:
                c, t = self.emit(s.init)
                self.w(f"const {s.typ.ctype()} {s.cname} = {c};")
                continue
            if s.init is not None:
                self.err(f"initialised (SAVE) local {s.name} is not supported")
            d = self.cdecl(s)
            if pr.is_program:
                self.w("static " + d)      # a main program's variables are SAVEd: zero-initialised
                continue
            if s.alloc:
                d = d[:-1] + " = {0, 0, 0, 0};"
                if not s.pointer:
                    allocs.append(s)
            elif s.typ.rank == 0 and self.has_default_init(s.typ):
                d = d[:-1] + " = {0};"       # the only default initialiser in use is c_null_ptr
            self.w(d)
        retlabel = "_ret" if not nested else f"_ret_{pr.name}"
        for ch in pr.children:
            self.emit_proc(ch, nested=True)
            self.scope_mod, self.scope_proc = pr.module, pr
        self.retlabel = retlabel
        for ln, toks in pr.body:
            self.cur_line = ln
            mark = len(self.out)
            self.stmt(toks)
            self.out.insert(mark, "  " * self.ind + f"/* {ln.file}:{ln.no} */")
        if self.blocks:
            self.cur_line = pr.line
            self.err("unterminated block")
        self.ind = base_ind + 1
        if self.uses_ret:
            self.w(f"{retlabel}: ;")
        for s in allocs:       # gfortran frees allocatable locals on exit
            self.w(f"if ({s.cname}.p) free({s.cname}.p);")
        self.w(f"PROF_EXIT({pid})")
        if pr.kind == "function":
            self.w(f"return {pr.syms[pr.result].cname};")
        self.ind = base_ind
        self.w("}")
        if pr.is_program:
            self.w(f"int main(int argc, char **argv) {{ f_argc = argc; f_argv = argv; {pr.cname}(); return 0; }}")
        if not nested:
            self.w("")
        else:
            self.scope_proc, self.blocks, self.selvar, self.ind, self.uses_ret, self.loopnames = saved

    def translate(self):
        self.ind = 0
        self.out.append(PRELUDE)
        # structs
        for m in self.order:
            self.scope_mod, self.scope_proc = m, None
            for d in m.types.values():
                self.out.append(f"struct {d.cname} {{")
                for f in d.fields:
                    self.out.append("  " + self.cdecl(f))
                self.out.append("};")
        # module variables / named constants
        for m in self.order:
            self.scope_mod, self.scope_proc = m, None
            for s in m.syms.values():
                if s.param:
                    if s.typ.rank > 0:
                        self.err(f"array named constant {s.name} is not supported")
                    if s.typ.base == "int" and s.typ.kind == 4:
                        c = str(self.const_int(s.init, None))
                    else:
                        c, t = self.emit(s.init)
                    self.out.append(f"#define {s.cname} (({s.typ.ctype()})({c}))")
                    self.out.append(f"{s.typ.ctype()} ref_const__{s.cname}(void) {{ return {s.cname}; }}")
                else:
                    if s.init is not None:
                        c, t = self.emit(s.init)
                        self.out.append(self.cdecl(s, static_tls=True)[:-1] + f" = {c};")
                    else:
                        self.out.append(self.cdecl(s, static_tls=True))
                    self.out.append(f"void *ref_addr__{s.cname}(void) {{ return (void *)&{s.cname}; }}")
        # procedures
        for m in self.order:
            for pr in m.procs.values():
                self.prepare_proc(pr)
        for m in self.order:
            for pr in m.procs.values():
                self.scope_mod, self.scope_proc = m, pr
                self.out.append(("extern " if pr.external else "") + self.proto(pr) + ";")
        self.prof_names = []
        for m in self.order:
            for pr in m.procs.values():
                if not pr.external:
                    self.emit_proc(pr)
        names = ", ".join(f'"{n}"' for n in self.prof_names)
        self.out.append("#ifdef REF_PROFILE")
        self.out.append(f"static const char *ref_prof_names[] = {{{names}}};")
        self.out.append(f"int ref_prof_count(void) {{ return {len(self.prof_names)}; }}")
        self.out.append("const char *ref_prof_name(int i) { return ref_prof_names[i]; }")
        self.out.append("unsigned long long ref_prof_cycles(int i) { return ref_prof_cyc[i]; }")
        self.out.append("unsigned long long ref_prof_ncalls(int i) { return ref_prof_calls[i]; }")
        self.out.append("void ref_prof_reset(void) { memset(ref_prof_cyc, 0, sizeof ref_prof_cyc); "
                        "memset(ref_prof_calls, 0, sizeof ref_prof_calls); }")
        self.out.append("#endif")
        return "\n".join(self.out) + "\n"

    def meta(self):
        def fld(s):
            dims = None
            if s.typ.rank > 0 and not s.alloc and not getattr(s, "assumed", False):
                self.scope_proc = None
                dims = [self.const_int(d, None) for d in s.dims]
            return {"name": s.name, "cname": s.cname, "type": s.typ.code(), "rank": s.typ.rank,
                    "alloc": s.alloc, "dims": dims, "intent": s.intent, "value": s.value}
        out = {"types": {}, "procs": {}, "vars": {}, "consts": {}}
        for m in self.order:
            self.scope_mod = m
            for d in m.types.values():
                out["types"][d.cname] = [fld(f) for f in d.fields]
            for s in m.syms.values():
                if not s.param:
                    out["vars"][s.cname] = fld(s)
                else:
                    out["consts"][s.cname] = s.typ.code()
            for pr in m.procs.values():
                out["procs"][pr.cname] = {
                    "kind": pr.kind, "args": [fld(pr.syms[a]) for a in pr.args],
                    "result": pr.rtype.code() if pr.rtype else None, "external": pr.external,
                    "where": f"{pr.line.file}:{pr.line.no}"}
        return out


def main(argv):
    out_c = out_meta = None
    files = []
    it = iter(argv)
    for a in it:
        if a == "-o":
            out_c = next(it)
        elif a == "-m":
            out_meta = next(it)
        else:
            files.append(a)
    tr = Translator()
    for f in files:
        tr.load(f)
    c = tr.translate()
    with open(out_c, "w") as fh:
        fh.write(c)
    if out_meta:
        with open(out_meta, "w") as fh:
            json.dump(tr.meta(), fh, indent=1)
    print(f"f90c: {len(files)} files -> {out_c} ({c.count(chr(10))} lines)")


if __name__ == "__main__":
    main(sys.argv[1:])
