"""f90run -- a small tree-walking interpreter for the Fortran 2008 subset LightKrylov's hot path is written in.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): nothing in lightkrylov_b200/ imports it.

Why it exists: neither this image nor the GPU box has a Fortran compiler (profiles/r02_probe_result.txt), so the reference
cannot be built.  This module EXECUTES THE REFERENCE'S OWN SOURCE TEXT instead: it reads the pre-expanded .f90 files where they
lie under /root/reference (src/Constants.f90, src/AbstractTypes/*.f90, src/Krylov/*.f90, ...), parses modules, submodules,
derived types with type-bound procedures, generic interfaces and procedure bodies, and runs them statement by statement on
numpy scalars / arrays.  No algorithm is restated here: the control flow, the operation order and every constant come from the
reference's text.  What IS supplied natively (the NATIVES / INTRINSICS tables at the end of this file) is what the reference itself takes from outside: Fortran
intrinsics, fortran-stdlib's optval / BLAS / LAPACK-backed linear algebra, and its logging / timing / error-reporting helpers
(no-ops, `stop_error` raises).

tests/golden/make_ref_golden.py uses it (in the container, where /root/reference exists) to produce the committed fixtures
tests/golden/ref_*.npz that pin oracle/lk_oracle.c to outputs of the reference's code.

Supported: free-form source, continuation lines, modules / submodules (`module procedure` bodies take their dummy declarations
from the parent's interface block), parameters, module variables, derived types (extends, allocatable components, default
initialisation, type-bound procedures with pass(name), deferred bindings), generic interfaces resolved on type / kind / rank of
the actual arguments, optional and keyword arguments, intent(out) semantics for allocatable components, automatic arrays,
allocatable polymorphic scalars and arrays (`allocate(.., source=/mold=)`), array sections (views), vector subscripts,
elemental subroutines, if / do / do while / select type / select case / block / associate, named loops with exit / cycle,
internal writes (ignored).  Anything else raises FortranError with the offending line -- it never guesses.
"""
import copy
import re

import numpy as np


class FortranError(Exception):
    pass


class StopError(FortranError):
    """stop_error / error stop reached in the interpreted program."""


class _Absent:
    def __repr__(self):
        return "ABSENT"


ABSENT = _Absent()


# ------------------------------------------------------------------------------------------------ source -> logical lines
def _strip_comment(line):
    q = None
    for i, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            return line[:i]
    return line


def _split_semicolons(line):
    out, q, cur = [], None, []
    for ch in line:
        if q:
            if ch == q:
                q = None
            cur.append(ch)
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == ";":
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur))
    return [s.strip() for s in out if s.strip()]


def logical_lines(text):
    """[(line_no, statement)] with comments removed, continuations joined, `;` split."""
    out, buf, start = [], "", 0
    for no, raw in enumerate(text.splitlines(), 1):
        s = _strip_comment(raw).strip()
        if not s or s.startswith("#"):
            continue
        if buf:
            if s.startswith("&"):
                s = s[1:]
        else:
            start = no
        if s.endswith("&"):
            buf += s[:-1]
            continue
        buf += s
        for st in _split_semicolons(buf):
            out.append((start, st))
        buf = ""
    return out


# ------------------------------------------------------------------------------------------------ tokens
_TOKEN = re.compile(r"""\s*(?:
  (?P<num>(?:\d+\.\d*(?![A-Za-z]+\.)|\.\d+|\d+)(?:[edED][+-]?\d+)?(?:_\w+)?)
 |(?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
 |(?P<dot>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.)
 |(?P<name>[A-Za-z_]\w*)
 |(?P<op>\*\*|//|==|/=|<=|>=|=>|::|\(/|/\)|[-+*/(),=<>%:\[\]])
)""", re.X | re.I)

_DOTMAP = {".eq.": "==", ".ne.": "/=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}


def tokenize(s):
    toks, pos = [], 0
    s = s.rstrip()
    while pos < len(s):
        m = _TOKEN.match(s, pos)
        if not m or m.end() == pos:
            raise FortranError(f"cannot tokenize {s[pos:]!r} in {s!r}")
        pos = m.end()
        k = m.lastgroup
        v = m.group(k)
        if k == "name":
            toks.append(("name", v.lower()))
        elif k == "dot":
            v = v.lower()
            if v in (".true.", ".false."):
                toks.append(("log", v == ".true."))
            else:
                toks.append(("op", _DOTMAP.get(v, v)))
        elif k == "str":
            q = v[0]
            toks.append(("str", v[1:-1].replace(q + q, q)))
        elif k == "num":
            toks.append(("num", v.lower()))
        else:
            toks.append(("op", v))
    return toks


def _parse_number(v):
    kind = None
    if "_" in v:
        v, kind = v.split("_", 1)
    if re.fullmatch(r"\d+", v):
        return int(v)                      # integer literals of any kind (1_c_int64_t, ...)
    if "d" in v:
        return np.float64(v.replace("d", "e"))
    if kind in ("dp", "c_double", "real64"):
        return np.float64(v)
    if kind in (None, "sp", "c_float", "real32"):
        return np.float32(v)           # default real is single precision
    raise FortranError(f"unknown kind suffix _{kind}")


# ------------------------------------------------------------------------------------------------ expression parser
class _P:
    def __init__(self, toks, src=""):
        self.t, self.i, self.src = toks, 0, src

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else (None, None)

    def at(self, v):
        return self.peek() == ("op", v)

    def eat(self, v=None):
        tok = self.peek()
        if v is not None and tok != ("op", v):
            raise FortranError(f"expected {v!r}, got {tok!r} in {self.src!r}")
        self.i += 1
        return tok

    def done(self):
        return self.i >= len(self.t)

    # precedence climbing
    def expr(self):
        a = self.p_or()
        while self.peek() in (("op", ".eqv."), ("op", ".neqv.")):
            op = self.eat()[1]
            a = ("bin", op, a, self.p_or())
        return a

    def p_or(self):
        a = self.p_and()
        while self.at(".or."):
            self.eat()
            a = ("bin", ".or.", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.at(".and."):
            self.eat()
            a = ("bin", ".and.", a, self.p_not())
        return a

    def p_not(self):
        if self.at(".not."):
            self.eat()
            return ("un", ".not.", self.p_not())
        return self.p_rel()

    def p_rel(self):
        a = self.p_cat()
        if self.peek()[0] == "op" and self.peek()[1] in ("==", "/=", "<", "<=", ">", ">="):
            op = self.eat()[1]
            a = ("bin", op, a, self.p_cat())
        return a

    def p_cat(self):
        a = self.p_add()
        while self.at("//"):
            self.eat()
            a = ("bin", "//", a, self.p_add())
        return a

    def p_add(self):
        if self.at("-") or self.at("+"):
            op = self.eat()[1]
            a = ("un", op, self.p_mul())
        else:
            a = self.p_mul()
        while self.at("+") or self.at("-"):
            op = self.eat()[1]
            a = ("bin", op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.at("*") or self.at("/"):
            op = self.eat()[1]
            a = ("bin", op, a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_primary()
        if self.at("**"):
            self.eat()
            if self.at("-") or self.at("+"):          # a ** -b (extension the reference does not use, but harmless)
                op = self.eat()[1]
                b = ("un", op, self.p_pow())
            else:
                b = self.p_pow()
            a = ("bin", "**", a, b)
        return a

    def p_primary(self):
        k, v = self.peek()
        if k == "num":
            self.eat()
            return ("lit", _parse_number(v))
        if k == "str":
            self.eat()
            return ("lit", v)
        if k == "log":
            self.eat()
            return ("lit", bool(v))
        if k == "op" and v == "(":
            self.eat()
            a = self.expr()
            if self.at(","):                          # complex literal (re, im)
                self.eat()
                b = self.expr()
                self.eat(")")
                return ("cplx", a, b)
            self.eat(")")
            return self.postfix(("paren", a))
        if k == "op" and v in ("[", "(/"):
            close = "]" if v == "[" else "/)"
            self.eat()
            items = []
            while not self.at(close):
                items.append(self.ac_item())
                if self.at(","):
                    self.eat()
            self.eat(close)
            return ("arr", items)
        if k == "name":
            self.eat()
            return self.postfix(("name", v))
        raise FortranError(f"unexpected token {self.peek()!r} in {self.src!r}")

    def ac_item(self):
        # implied do: ( expr, i = a, b )
        if self.at("("):
            save = self.i
            try:
                self.eat()
                e = self.expr()
                if self.at(","):
                    self.eat()
                    if self.peek()[0] == "name" and self.peek(1) == ("op", "="):
                        var = self.eat()[1]
                        self.eat("=")
                        lo = self.expr()
                        self.eat(",")
                        hi = self.expr()
                        st = None
                        if self.at(","):
                            self.eat()
                            st = self.expr()
                        self.eat(")")
                        return ("implied", e, var, lo, hi, st)
                raise FortranError("not an implied do")
            except FortranError:
                self.i = save
        return self.expr()

    def postfix(self, a):
        while True:
            if self.at("("):
                self.eat()
                args = []
                while not self.at(")"):
                    args.append(self.arg())
                    if self.at(","):
                        self.eat()
                self.eat(")")
                a = ("call", a, args)
            elif self.at("%"):
                self.eat()
                k, v = self.eat()
                if k != "name":
                    raise FortranError(f"component name expected in {self.src!r}")
                a = ("comp", a, v)
            else:
                return a

    def arg(self):
        """(keyword|None, ast) where ast may be ('slice', lo, hi, step)"""
        if self.peek()[0] == "name" and self.peek(1) == ("op", "=") :
            kw = self.eat()[1]
            self.eat("=")
            return (kw, self.expr())
        lo = None
        if not self.at(":"):
            lo = self.expr()
            if not self.at(":"):
                return (None, lo)
        self.eat(":")
        hi = st = None
        if not (self.at(",") or self.at(")") or self.at(":")):
            hi = self.expr()
        if self.at(":"):
            self.eat()
            st = self.expr()
        return (None, ("slice", lo, hi, st))


def parse_expr(s):
    p = _P(tokenize(s), s)
    e = p.expr()
    if not p.done():
        raise FortranError(f"trailing tokens in expression {s!r}")
    return e


# ------------------------------------------------------------------------------------------------ declarations
class TypeSpec:
    __slots__ = ("base", "kind", "tname", "charlen")

    def __init__(self, base, kind=None, tname=None):
        self.base, self.kind, self.tname = base, kind, tname

    def __repr__(self):
        return f"{self.base}({self.kind or self.tname or ''})"

    def dtype(self):
        if self.base == "integer":
            return np.int64
        if self.base == "real":
            return np.float64 if self.kind == "dp" else np.float32
        if self.base == "complex":
            return np.complex128 if self.kind == "dp" else np.complex64
        if self.base == "logical":
            return np.bool_
        return object


class Decl:
    """one declared entity"""
    __slots__ = ("name", "ts", "dims", "attrs", "init", "intent", "optional", "allocatable", "parameter")

    def __init__(self, name, ts, dims, attrs, init):
        self.name, self.ts, self.dims, self.attrs, self.init = name, ts, dims, attrs, init
        self.intent = attrs.get("intent")
        self.optional = "optional" in attrs
        self.allocatable = "allocatable" in attrs or "pointer" in attrs
        self.parameter = "parameter" in attrs

    @property
    def rank(self):
        return len(self.dims) if self.dims else 0


_TYPE_KW = ("integer", "real", "complex", "logical", "character", "type", "class", "double", "procedure")


def _split_top(toks, sep=","):
    """split a token list at top-level separators"""
    out, cur, depth = [], [], 0
    for t in toks:
        if t[0] == "op" and t[1] in ("(", "[", "(/"):
            depth += 1
        elif t[0] == "op" and t[1] in (")", "]", "/)"):
            depth -= 1
        if depth == 0 and t == ("op", sep):
            out.append(cur)
            cur = []
        else:
            cur.append(t)
    out.append(cur)
    return out


def _find_top(toks, v):
    depth = 0
    for i, t in enumerate(toks):
        if t[0] == "op" and t[1] in ("(", "[", "(/"):
            depth += 1
        elif t[0] == "op" and t[1] in (")", "]", "/)"):
            depth -= 1
        elif depth == 0 and t == ("op", v):
            return i
    return -1


def _paren_group(toks, i):
    """toks[i] == '(' -> (inner tokens, index after the closing paren)"""
    assert toks[i] == ("op", "(")
    depth = 0
    for j in range(i, len(toks)):
        if toks[j] == ("op", "("):
            depth += 1
        elif toks[j] == ("op", ")"):
            depth -= 1
            if depth == 0:
                return toks[i + 1:j], j + 1
    raise FortranError("unbalanced parentheses")


def _parse_typespec(toks, src):
    """leading type-spec of a declaration -> (TypeSpec, rest tokens)"""
    base = toks[0][1]
    i = 1
    if base == "double":          # double precision
        return TypeSpec("real", "dp"), toks[2:]
    kind = tname = None
    if i < len(toks) and toks[i] == ("op", "("):
        inner, i = _paren_group(toks, i)
        if base in ("type", "class"):
            tname = "*" if inner == [("op", "*")] else inner[0][1]
        elif base == "procedure":
            tname = inner[0][1] if inner else None
        elif base == "character":
            kind = None
        else:
            # (dp) | (kind=dp)
            ks = [t for t in inner if t[0] == "name" and t[1] != "kind"]
            kind = ks[0][1] if ks else None
            kind = {"c_double": "dp", "c_float": "sp", "real64": "dp", "real32": "sp", "c_double_complex": "dp",
                    "c_float_complex": "sp"}.get(kind, kind)
            if kind is None and inner and inner[0][0] == "num":
                kind = {"4": "sp", "8": "dp"}.get(inner[0][1])
    ts = TypeSpec("class" if base == "class" else base, kind, tname)
    return ts, toks[i:]


def parse_declaration(toks, src):
    """type-spec [, attr]... :: entity [, entity]...   -> [Decl]"""
    ts, rest = _parse_typespec(toks, src)
    k = _find_top(rest, "::")
    if k < 0:
        raise FortranError(f"declaration without '::' not supported: {src!r}")
    attr_toks, ent_toks = rest[:k], rest[k + 1:]
    attrs = {}
    for a in _split_top(attr_toks):
        if not a:
            continue
        nm = a[0][1]
        if nm == "intent":
            attrs["intent"] = "".join(t[1] for t in a[2:-1])
        elif nm == "dimension":
            inner, _ = _paren_group(a, 1)
            attrs["dimension"] = [_parse_dim(d, src) for d in _split_top(inner)]
        else:
            attrs[nm] = True
    decls = []
    for e in _split_top(ent_toks):
        name = e[0][1]
        dims, init, j = attrs.get("dimension"), None, 1
        if j < len(e) and e[j] == ("op", "("):
            inner, j = _paren_group(e, j)
            dims = [_parse_dim(d, src) for d in _split_top(inner)]
        if j < len(e) and e[j][0] == "op" and e[j][1] in ("=", "=>"):
            p = _P(e[j + 1:], src)
            init = p.expr()
        decls.append(Decl(name, ts, dims, attrs, init))
    return decls


def _parse_dim(toks, src):
    """':' -> None (deferred/assumed) ; 'n' -> (None, ast) ; 'lo:hi' -> (ast, ast)"""
    if toks == [("op", ":")]:
        return None
    k = _find_top(toks, ":")
    if k >= 0:
        lo = _P(toks[:k], src).expr() if toks[:k] else None
        hi_t = toks[k + 1:]
        if not hi_t or hi_t == [("op", "*")]:
            return None if lo is None else ("lb", lo)
        return (lo, _P(hi_t, src).expr())
    if toks == [("op", "*")]:
        return None
    return (None, _P(toks, src).expr())


# ------------------------------------------------------------------------------------------------ program structure
class TypeDef:
    def __init__(self, name, parent):
        self.name, self.parent = name, parent
        self.components = {}          # name -> Decl
        self.bindings = {}            # binding -> (procname|None, passname|None, deferred, nopass)
        self.generics = {}            # generic binding -> [binding names]


class Proc:
    def __init__(self, name, kind, args, result, prefixes, module):
        self.name, self.kind, self.args, self.result = name, kind, args, result
        self.prefixes, self.module = prefixes, module
        self.body = []                # statement nodes (declarations included, executed in order)
        self.decls = {}               # name -> Decl (dummies + locals), filled at parse time
        self.result_ts = None
        self.elemental = "elemental" in prefixes
        self.file = None
        self.internals = {}           # internal procedure name -> global (mangled) name

    def dummy(self, name):
        return self.decls.get(name)


class Inst:
    """instance of a derived type"""
    __slots__ = ("tname", "f")

    def __init__(self, tname, f=None):
        self.tname, self.f = tname, f if f is not None else {}

    def __repr__(self):
        return f"<{self.tname} {list(self.f)}>"


class CPtr:
    """type(c_ptr) / type(c_funptr): `obj` is None (c_null_ptr), a numpy array (c_loc of an array: shared memory), a ScalarRef
    (c_loc of a scalar variable) or any Python object a native handed out as an opaque handle.  Copying the Fortran value
    copies the POINTER (shallow), also inside deep copies of derived-type values."""
    __slots__ = ("obj",)

    def __init__(self, obj=None):
        self.obj = obj

    def __deepcopy__(self, memo):
        return CPtr(self.obj)

    def __repr__(self):
        return f"CPtr({type(self.obj).__name__})"


class ScalarRef:
    """c_loc(scalar variable): reads and writes go to the variable's slot in its scope"""
    __slots__ = ("interp", "sc", "ast")

    def __init__(self, interp, sc, ast):
        self.interp, self.sc, self.ast = interp, sc, ast

    def get(self):
        return self.interp.ev(self.ast, self.sc)

    def set(self, v):
        self.interp.raw_store(self.ast, v, self.sc)


class _Exit(Exception):
    def __init__(self, label):
        self.label = label


class _Cycle(Exception):
    def __init__(self, label):
        self.label = label


class _Return(Exception):
    pass


_END_RE = re.compile(r"^end\s*(module|submodule|subroutine|function|procedure|type|interface|program)?\b")


class Program:
    """parsed modules: global tables shared by every loaded file"""

    def __init__(self):
        self.procs = {}               # name -> Proc (with body)
        self.signatures = {}          # name -> Proc (interface-only: dummies of module procedures / abstract interfaces)
        self.generics = {}            # generic name -> [specific names]
        self.types = {}               # name -> TypeDef
        self.globals = {}             # parameters and module variables
        self.global_decls = {}
        self.natives = {}             # name -> python callable(interp, args, kwargs)
        self.pending_globals = []     # (Decl, src) evaluated after loading, in order

    # ---- loading
    def load(self, path):
        with open(path) as fh:
            lines = logical_lines(fh.read())
        self._file = path
        self._parse_unit(lines)

    def _parse_unit(self, lines):
        i = 0
        n = len(lines)
        module = None
        while i < n:
            no, s = lines[i]
            low = s.lower()
            toks = tokenize(s)
            t0 = toks[0][1] if toks else ""
            if t0 == "module" and len(toks) == 2:
                module = toks[1][1]
                i += 1
            elif t0 == "submodule":
                module = toks[-1][1]
                i += 1
            elif t0 == "program":
                raise FortranError("program units are not supported")
            elif _END_RE.match(low) and re.match(r"^end\s*(module|submodule)\b", low) or low == "end":
                i += 1
            elif t0 in ("use", "implicit", "private", "public", "contains", "save", "external", "intrinsic", "import"):
                i += 1
            elif t0 in ("interface", "abstract"):
                i = self._parse_interface(lines, i, module)
            elif t0 == "type" and len(toks) > 1 and toks[1] != ("op", "("):
                i = self._parse_type(lines, i)
            elif self._is_proc_header(toks):
                i = self._parse_proc(lines, i, module, interface_only=False)
            elif t0 in _TYPE_KW and _find_top(toks, "::") >= 0:
                for d in parse_declaration(toks, s):
                    self.pending_globals.append((d, s))
                i += 1
            else:
                raise FortranError(f"{self._file}:{no}: unsupported specification statement {s!r}")

    @staticmethod
    def _is_proc_header(toks):
        """[prefix]... subroutine|function name ...  |  module procedure name   (prefixes: pure, elemental, a type-spec, ...)"""
        depth, prev = 0, None
        for k, t in enumerate(toks):
            if t[0] == "op" and t[1] == "(":
                depth += 1
                continue
            if t[0] == "op" and t[1] == ")":
                depth -= 1
                continue
            if depth:
                continue
            if t[0] != "name":
                return False
            nm = t[1]
            if nm in ("subroutine", "function"):
                return k + 1 < len(toks) and toks[k + 1][0] == "name"
            if nm == "procedure" and prev == "module" and len(toks) <= 4:
                return True
            if nm not in ("module", "pure", "impure", "elemental", "recursive", "integer", "real", "logical", "complex",
                          "type", "class", "character", "double", "precision"):
                return False
            prev = nm
        return False

    def _parse_interface(self, lines, i, module):
        no, s = lines[i]
        toks = tokenize(s)
        gname = None
        if toks[0][1] == "interface" and len(toks) > 1:
            if toks[1][0] == "name" and toks[1][1] in ("operator", "assignment"):
                gname = "".join(str(t[1]) for t in toks[1:])
            else:
                gname = toks[1][1]
        i += 1
        specifics = []
        while True:
            no, s = lines[i]
            low = s.lower()
            toks = tokenize(s)
            if re.match(r"^end\s*interface\b", low):
                i += 1
                break
            names = [t[1] for t in toks if t[0] == "name"]
            if names[:2] == ["module", "procedure"] or names[:1] == ["procedure"]:
                start = 2 if names[0] == "module" else 1
                specifics += [t[1] for t in toks[start:] if t[0] == "name"]
                i += 1
            elif self._is_proc_header(toks):
                j = self._parse_proc(lines, i, module, interface_only=True)
                specifics.append(self._last_proc.name)
                i = j
            else:
                raise FortranError(f"{self._file}:{no}: unsupported statement in interface: {s!r}")
        if gname:
            lst = self.generics.setdefault(gname, [])
            for sp in specifics:
                if sp not in lst:
                    lst.append(sp)
        return i

    def _parse_type(self, lines, i):
        no, s = lines[i]
        toks = tokenize(s)
        k = _find_top(toks, "::")
        name = toks[k + 1][1] if k >= 0 else toks[1][1]
        parent = None
        for a in _split_top(toks[1:k if k >= 0 else 1]):
            if a and a[0][1] == "extends":
                parent = a[2][1]
        td = TypeDef(name, parent)
        i += 1
        in_contains = False
        while True:
            no, s = lines[i]
            low = s.lower()
            toks = tokenize(s)
            if re.match(r"^end\s*type\b", low):
                i += 1
                break
            t0 = toks[0][1]
            if t0 == "contains":
                in_contains = True
            elif t0 in ("private", "public", "sequence") and len(toks) == 1:
                pass
            elif in_contains:
                self._parse_binding(td, toks, s)
            else:
                for d in parse_declaration(toks, s):
                    td.components[d.name] = d
            i += 1
        self.types[name] = td
        return i

    def _parse_binding(self, td, toks, src):
        t0 = toks[0][1]
        k = _find_top(toks, "::")
        if t0 == "generic":
            ents = toks[k + 1:]
            arrow = _find_top(ents, "=>")
            gname = "".join(str(t[1]) for t in ents[:arrow])
            td.generics.setdefault(gname, []).extend(t[1] for t in ents[arrow + 1:] if t[0] == "name")
            return
        if t0 == "final":
            return
        if t0 != "procedure":
            raise FortranError(f"unsupported type-bound statement {src!r}")
        passname, deferred, nopass = None, False, False
        j = 1
        if toks[j] == ("op", "("):
            _, j = _paren_group(toks, j)          # procedure(interface): deferred binding
        for a in _split_top(toks[j:k] if k >= 0 else []):
            if not a:
                continue
            if a[0][1] == "pass" and len(a) > 1:
                passname = a[2][1]
            elif a[0][1] == "deferred":
                deferred = True
            elif a[0][1] == "nopass":
                nopass = True
        ents = toks[k + 1:] if k >= 0 else toks[j:]
        for e in _split_top(ents):
            bname = e[0][1]
            target = e[2][1] if len(e) >= 3 and e[1] == ("op", "=>") else (None if deferred else bname)
            td.bindings[bname] = (target, passname, deferred, nopass)

    def _parse_proc(self, lines, i, module, interface_only, host=None):
        no, s = lines[i]
        toks = tokenize(s)
        prefixes, result_ts = [], None
        j = 0
        while toks[j][1] not in ("subroutine", "function", "procedure"):
            if toks[j][1] in ("integer", "real", "logical", "complex", "type", "class", "character"):
                result_ts, rest = _parse_typespec(toks[j:], s)
                j = len(toks) - len(rest)
                continue
            prefixes.append(toks[j][1])
            j += 1
        kind = toks[j][1]
        name = toks[j + 1][1]
        args, result = [], None
        j += 2
        if kind == "procedure":
            kind = "modproc"
        else:
            if j < len(toks) and toks[j] == ("op", "("):
                inner, j = _paren_group(toks, j)
                args = [t[1] for t in inner if t[0] == "name"]
            if kind == "function":
                result = name
            while j < len(toks) and toks[j][0] == "name" and toks[j][1] in ("result", "bind"):
                inner, j2 = _paren_group(toks, j + 1)
                if toks[j][1] == "result":
                    result = inner[0][1]
                j = j2
        if host is not None:
            host.internals[name] = host.name + "::" + name
            name = host.name + "::" + name
        proc = Proc(name, kind, args, result, prefixes, module)
        proc.result_ts = result_ts
        proc.file = self._file
        i += 1
        body, i = self._parse_block(lines, i, ("end_proc",), proc)
        proc.body = body
        self._last_proc = proc
        if interface_only:
            self.signatures[name] = proc
        else:
            if kind == "modproc":
                sig = self.signatures.get(name)
                if sig is None:
                    raise FortranError(f"module procedure {name}: interface not loaded (load the parent module first)")
                proc.args, proc.result, proc.prefixes = sig.args, sig.result, sig.prefixes
                proc.elemental, proc.result_ts = sig.elemental, sig.result_ts
                proc.body = [st for st in sig.body if st[0] == "decl"] + proc.body
                merged = dict(sig.decls)
                merged.update(proc.decls)
                proc.decls = merged
            self.procs[name] = proc
        return i

    # ---- executable / declaration statements of a procedure body
    def _parse_block(self, lines, i, terminators, proc):
        """returns (stmts, next_i); the terminator line is consumed, self._term holds (kind, tokens, src)"""
        stmts = []
        while True:
            if i >= len(lines):
                raise FortranError(f"{self._file}: unexpected end of file (open construct)")
            no, s = lines[i]
            low = re.sub(r"\s+", " ", s.lower())
            toks = tokenize(s)
            label = None
            if len(toks) > 2 and toks[0][0] == "name" and toks[1] == ("op", ":") and toks[2][0] == "name" \
                    and toks[2][1] in ("do", "if", "block", "associate", "select"):
                label = toks[0][1]
                toks = toks[2:]
                low = low.split(":", 1)[1].strip()
            t0 = toks[0][1] if toks[0][0] == "name" else None
            where = (self._file, no, s)
            # ---- terminators
            term = None
            if re.match(r"^end ?(subroutine|function|procedure)\b", low) or low == "end":
                term = "end_proc"
            elif re.match(r"^end ?if\b", low):
                term = "end_if"
            elif re.match(r"^end ?do\b", low):
                term = "end_do"
            elif re.match(r"^end ?select\b", low):
                term = "end_select"
            elif re.match(r"^end ?block\b", low):
                term = "end_block"
            elif re.match(r"^end ?associate\b", low):
                term = "end_associate"
            elif re.match(r"^else ?if\b", low) and toks[-1] == ("name", "then"):
                term = "else_if"
            elif t0 == "else" and len(toks) <= 2:
                term = "else"
            elif (t0 in ("type", "class") and len(toks) > 1 and toks[1] == ("name", "is")) or \
                    (t0 == "class" and toks[1:2] == [("name", "default")]) or \
                    (t0 == "case" and "select_case" in terminators):
                term = "guard"
            if term is not None:
                if term not in terminators:
                    raise FortranError(f"{self._file}:{no}: unexpected {s!r} (expected one of {terminators})")
                self._term = (term, toks, s)
                return stmts, i + 1
            i += 1
            # ---- declarations / ignorable
            if t0 in ("use", "implicit", "import", "external", "intrinsic", "save"):
                continue
            if t0 == "contains":
                # internal procedures: registered as "<host>::<name>"; host association is NOT provided (they may only use
                # their own dummies / locals and module entities -- enough for the selector functions the reference passes around)
                while True:
                    no2, s2 = lines[i]
                    low2 = re.sub(r"\s+", " ", s2.lower())
                    if re.match(r"^end ?(subroutine|function|procedure)\b", low2) or low2 == "end":
                        self._term = ("end_proc", tokenize(s2), s2)
                        return stmts, i + 1
                    i = self._parse_proc(lines, i, proc.module, interface_only=False, host=proc)
            if t0 in _TYPE_KW and _find_top(toks, "::") >= 0 and not (t0 == "type" and toks[1] != ("op", "(")):
                ds = parse_declaration(toks, s)
                for d in ds:
                    proc.decls[d.name] = d
                stmts.append(("decl", ds, where))
                continue
            # ---- constructs
            if t0 == "if" and toks[-1] == ("name", "then"):
                cond_t, _ = _paren_group(toks, 1)
                branches, orelse = [], None
                cond = _P(cond_t, s).expr()
                while True:
                    blk, i = self._parse_block(lines, i, ("else_if", "else", "end_if"), proc)
                    term, ttoks, tsrc = self._term
                    if cond is not None:
                        branches.append((cond, blk))
                    else:
                        orelse = blk
                    if term == "end_if":
                        break
                    if term == "else_if":
                        k = next(ix for ix, t in enumerate(ttoks) if t == ("op", "("))
                        cond_t, _ = _paren_group(ttoks, k)
                        cond = _P(cond_t, tsrc).expr()
                    else:
                        cond = None
                stmts.append(("if", branches, orelse, where))
                continue
            if t0 == "if":
                cond_t, j = _paren_group(toks, 1)
                inner = self._simple_stmt(toks[j:], s, where)
                stmts.append(("if", [(_P(cond_t, s).expr(), [inner])], None, where))
                continue
            if t0 == "do":
                if len(toks) == 1:
                    hdr = ("forever",)
                elif toks[1] == ("name", "while"):
                    cond_t, _ = _paren_group(toks, 2)
                    hdr = ("while", _P(cond_t, s).expr())
                else:
                    var = toks[1][1]
                    parts = _split_top(toks[3:])
                    hdr = ("count", var, _P(parts[0], s).expr(), _P(parts[1], s).expr(),
                           _P(parts[2], s).expr() if len(parts) > 2 else None)
                blk, i = self._parse_block(lines, i, ("end_do",), proc)
                stmts.append(("do", hdr, blk, label, where))
                continue
            if t0 == "select" and toks[1] == ("name", "type"):
                inner, _ = _paren_group(toks, 2)
                arrow = _find_top(inner, "=>")
                if arrow >= 0:
                    assoc, sel = inner[0][1], _P(inner[arrow + 1:], s).expr()
                else:
                    sel = _P(inner, s).expr()
                    assoc = inner[0][1] if len(inner) == 1 else None
                guards = []
                blk, i = self._parse_block(lines, i, ("guard", "end_select"), proc)      # nothing before the first guard
                while self._term[0] == "guard":
                    gt = self._term[1]
                    if gt[1] == ("name", "default"):
                        g = ("default", None)
                    else:
                        inner_g, _ = _paren_group(gt, 2)
                        g = (gt[0][1], inner_g[0][1])            # ('type'|'class', name)
                    blk, i = self._parse_block(lines, i, ("guard", "end_select"), proc)
                    guards.append((g, blk))
                stmts.append(("select_type", sel, assoc, guards, where))
                continue
            if t0 == "select" and toks[1] == ("name", "case"):
                inner, _ = _paren_group(toks, 2)
                sel = _P(inner, s).expr()
                cases = []
                blk, i = self._parse_block(lines, i, ("guard", "end_select", "select_case"), proc)
                while self._term[0] == "guard":
                    gt = self._term[1]
                    if gt[1] == ("name", "default"):
                        vals = None
                    else:
                        inner_g, _ = _paren_group(gt, 1)
                        vals = [_P(v, s).arg()[1] for v in _split_top(inner_g)]
                    blk, i = self._parse_block(lines, i, ("guard", "end_select", "select_case"), proc)
                    cases.append((vals, blk))
                stmts.append(("select_case", sel, cases, where))
                continue
            if t0 == "block" and len(toks) == 1:
                blk, i = self._parse_block(lines, i, ("end_block",), proc)
                stmts.append(("block", blk, where))
                continue
            if t0 == "associate":
                inner, _ = _paren_group(toks, 1)
                pairs = []
                for a in _split_top(inner):
                    pairs.append((a[0][1], _P(a[2:], s).expr()))
                blk, i = self._parse_block(lines, i, ("end_associate",), proc)
                stmts.append(("associate", pairs, blk, where))
                continue
            stmts.append(self._simple_stmt(toks, s, where))

    def _simple_stmt(self, toks, s, where):
        t0 = toks[0][1] if toks[0][0] == "name" else None
        eq = _find_top(toks, "=")
        if eq > 0 and t0 != "call":
            lhs = _P(toks[:eq], s)
            target = lhs.p_primary()
            if lhs.done():
                return ("assign", target, _P(toks[eq + 1:], s).expr(), where)
        arrow = _find_top(toks, "=>")
        if arrow > 0:
            lhs = _P(toks[:arrow], s)
            target = lhs.p_primary()
            return ("ptr_assign", target, _P(toks[arrow + 1:], s).expr(), where)
        if t0 == "call":
            p = _P(toks[1:], s)
            e = p.p_primary()
            if e[0] != "call":
                e = ("call", e, [])
            return ("callsub", e, where)
        if t0 == "allocate":
            inner, _ = _paren_group(toks, 1)
            items, opts = [], {}
            for a in _split_top(inner):
                if len(a) > 1 and a[0][0] == "name" and a[1] == ("op", "="):
                    opts[a[0][1]] = _P(a[2:], s).expr()
                else:
                    items.append(_P(a, s).p_primary())
            return ("allocate", items, opts, where)
        if t0 == "deallocate":
            inner, _ = _paren_group(toks, 1)
            items = []
            for a in _split_top(inner):
                if len(a) > 1 and a[0][0] == "name" and a[1] == ("op", "="):
                    continue
                items.append(_P(a, s).p_primary())
            return ("deallocate", items, where)
        if t0 in ("write", "print", "flush", "read", "open", "close", "rewind"):
            return ("io", s, where)
        if t0 == "return":
            return ("return", where)
        if t0 == "exit":
            return ("exit", toks[1][1] if len(toks) > 1 else None, where)
        if t0 == "cycle":
            return ("cycle", toks[1][1] if len(toks) > 1 else None, where)
        if t0 == "continue":
            return ("nop", where)
        if t0 == "stop" or (t0 == "error" and toks[1:2] == [("name", "stop")]):
            return ("stop", s, where)
        if t0 == "nullify":
            return ("nop", where)
        raise FortranError(f"{where[0]}:{where[1]}: unsupported statement {s!r}")


# ------------------------------------------------------------------------------------------------ interpreter
class Scope:
    __slots__ = ("vars", "decls", "proc")

    def __init__(self, proc):
        self.vars, self.decls, self.proc = {}, dict(proc.decls) if proc else {}, proc


_MISSING = object()
_REAL_TYPES = (float, np.floating)
_CPLX_TYPES = (complex, np.complexfloating)
_INT_TYPES = (int, np.integer)


def _is_int(v):
    return isinstance(v, _INT_TYPES) and not isinstance(v, (bool, np.bool_))


def _kind_of_value(v):
    """('integer'|'real'|'complex'|'logical'|'character'|'derived', kind|tname, rank)"""
    if isinstance(v, np.ndarray):
        rank = v.ndim
        if v.dtype == object:
            first = next((x for x in v.ravel() if x is not None), None)
            return ("derived", first.tname if isinstance(first, Inst) else None, rank)
        dt = v.dtype
        if dt == np.bool_:
            return ("logical", None, rank)
        if np.issubdtype(dt, np.integer):
            return ("integer", None, rank)
        if np.issubdtype(dt, np.complexfloating):
            return ("complex", "sp" if dt == np.complex64 else "dp", rank)
        if np.issubdtype(dt, np.floating):
            return ("real", "sp" if dt == np.float32 else "dp", rank)
        return ("character", None, rank)
    if isinstance(v, Inst):
        return ("derived", v.tname, 0)
    if isinstance(v, (bool, np.bool_)):
        return ("logical", None, 0)
    if _is_int(v):
        return ("integer", None, 0)
    if isinstance(v, _CPLX_TYPES):
        return ("complex", "sp" if isinstance(v, np.complex64) else "dp", 0)
    if isinstance(v, _REAL_TYPES):
        return ("real", "sp" if isinstance(v, np.float32) else "dp", 0)
    if isinstance(v, str):
        return ("character", None, 0)
    return (None, None, 0)


class Interp:
    def __init__(self, program):
        self.p = program
        self.intrinsics = dict(INTRINSICS)
        self.natives = dict(NATIVES)
        self.natives.update(program.natives)
        self.trace = False
        self.call_depth = 0
        self.hooks = {}               # reference procedure name -> logical function tried first (see run_proc_scalar)
        self.hook_hits = {}
        self.called = set()           # names of the interpreted procedures that ran (coverage of a test session)
        gsc = Scope(None)
        gsc.vars = self.p.globals
        self.gscope = gsc
        for name, val in (("c_null_ptr", CPtr(None)), ("c_null_funptr", CPtr(None)), ("c_double", "dp"), ("c_float", "sp"),
                          ("c_int", "c_int"), ("c_int32_t", "c_int32_t"), ("c_int64_t", "c_int64_t"), ("c_size_t", "c_size_t"),
                          ("c_bool", "c_bool"), ("c_char", "c_char"), ("ilp", "ilp"), ("c_null_char", "\x00")):
            self.p.globals.setdefault(name, val)
        for d, src in self.p.pending_globals:
            self.p.global_decls[d.name] = d
            self._declare(d, gsc)
        self.p.pending_globals = []

    # ---------------------------------------------------------------- types
    def type_chain(self, tname):
        chain = []
        while tname is not None and tname in self.p.types:
            td = self.p.types[tname]
            chain.append(td)
            tname = td.parent
        return chain                         # most derived first

    def isa(self, tname, ancestor):
        if tname == ancestor:
            return True
        return any(td.name == ancestor for td in self.type_chain(tname))

    def component_decl(self, tname, comp):
        for td in self.type_chain(tname):
            if comp in td.components:
                return td.components[comp]
        return None

    def new_inst(self, tname):
        inst = Inst(tname)
        for td in reversed(self.type_chain(tname)):
            for d in td.components.values():
                inst.f[d.name] = self._default_value(d, self.gscope)
        return inst

    def find_binding(self, tname, bname):
        for td in self.type_chain(tname):
            if bname in td.generics:
                return ("generic", td.generics[bname])
            if bname in td.bindings and not td.bindings[bname][2]:
                return ("specific", td.bindings[bname])
        return None

    # ---------------------------------------------------------------- declarations
    def _shape_of(self, dims, sc):
        shape = []
        for dm in dims:
            if dm is None or dm[0] == "lb":
                return None
            lo, hi = dm
            lo_v = 1 if lo is None else int(self.ev(lo, sc))
            if lo_v != 1:
                raise FortranError("array lower bounds other than 1 are not supported")
            shape.append(max(0, int(self.ev(hi, sc))))
        return tuple(shape)

    def _undefined_scalar(self, ts):
        if ts.base == "integer":
            return 0
        if ts.base == "real":
            return ts.dtype()(np.nan)
        if ts.base == "complex":
            return ts.dtype()(complex(np.nan, np.nan))
        if ts.base == "logical":
            return False
        if ts.base == "character":
            return ""
        return None

    def _new_array(self, ts, shape, fill=None):
        dt = ts.dtype()
        if dt is object:
            a = np.empty(shape, dtype=object, order="F")
            if ts.tname in self.p.types:      # class(T) elements allocated without source= have the declared type as dynamic type
                for ix in np.ndindex(*shape):
                    a[ix] = self.new_inst(ts.tname)
            return a
        if fill is None:
            fill = self._undefined_scalar(ts)
        return np.full(shape, fill, dtype=dt, order="F")

    def _default_value(self, d, sc):
        if d.allocatable:
            return None
        if d.dims:
            shape = self._shape_of(d.dims, sc)
            if shape is None:
                return None
            a = self._new_array(d.ts, shape)
            if d.init is not None:
                v = self.ev(d.init, sc)
                if isinstance(v, np.ndarray) and v.shape != a.shape:
                    v = v.reshape(a.shape, order="F")
                a[...] = v
            return a
        if d.init is not None:
            return self.coerce(d.ts, self.ev(d.init, sc))
        if d.ts.base == "type" and d.ts.tname in ("c_ptr", "c_funptr"):
            return CPtr(None)
        if d.ts.base == "type":
            return self.new_inst(d.ts.tname) if d.ts.tname in self.p.types else Inst(d.ts.tname)
        if d.ts.base in ("class", "procedure"):
            return None
        return self._undefined_scalar(d.ts)

    def _declare(self, d, sc):
        sc.decls[d.name] = d
        if d.name in sc.vars and sc.proc is not None and d.name in sc.proc.args:
            return
        sc.vars[d.name] = self._default_value(d, sc)

    def coerce(self, ts, v):
        if v is None or v is ABSENT or isinstance(v, np.ndarray):
            return v
        b = ts.base
        if isinstance(v, str):                 # kind parameters (sp, dp) are kept symbolic
            return v
        try:
            if b == "integer":
                return int(v.real) if isinstance(v, _CPLX_TYPES) else int(v)
            if b == "real":
                return ts.dtype()(v.real if isinstance(v, _CPLX_TYPES) else v)
            if b == "complex":
                return ts.dtype()(v)
            if b == "logical":
                return bool(v)
        except (TypeError, ValueError) as exc:
            raise FortranError(f"cannot convert {v!r} to {ts!r}: {exc}")
        return v

    # ---------------------------------------------------------------- expressions
    def lookup(self, name, sc):
        v = sc.vars.get(name, _MISSING)
        if v is _MISSING:
            v = self.p.globals.get(name, _MISSING)
        if v is _MISSING:
            raise FortranError(f"undefined name {name!r}")
        return v

    def has_var(self, name, sc):
        return name in sc.vars or name in self.p.globals

    def ev(self, e, sc):
        k = e[0]
        if k == "lit":
            return e[1]
        if k == "name":
            nm = e[1]
            if sc.proc is not None and nm in sc.proc.internals and nm not in sc.vars:
                return ("procref", sc.proc.internals[nm])
            if not self.has_var(nm, sc) and (nm in self.p.procs or nm in self.natives or nm in self.p.generics):
                return ("procref", nm)                  # procedure passed as an actual argument
            return self.lookup(nm, sc)
        if k == "paren":
            v = self.ev(e[1], sc)
            return v.copy() if isinstance(v, np.ndarray) else v
        if k == "cplx":
            a, b = self.ev(e[1], sc), self.ev(e[2], sc)
            dp = isinstance(a, np.float64) or isinstance(b, np.float64)
            return (np.complex128 if dp else np.complex64)(complex(a, b))
        if k == "un":
            a = self.ev(e[2], sc)
            if e[1] == "-":
                return -a
            if e[1] == "+":
                return a
            return np.logical_not(a) if isinstance(a, np.ndarray) else (not a)
        if k == "bin":
            return self.binop(e[1], self.ev(e[2], sc), self.ev(e[3], sc))
        if k == "arr":
            return self.array_constructor(e[1], sc)
        if k == "comp":
            base = self.ev(e[1], sc)
            if isinstance(base, Inst):
                if e[2] in base.f:
                    return base.f[e[2]]
                if self.find_binding(base.tname, e[2]) is not None:      # parameterless type-bound function without ()
                    return self.call_bound(base, e[1], e[2], [], sc, want_result=True)
                raise FortranError(f"type {base.tname} has no component {e[2]!r}")
            if isinstance(base, _CPLX_TYPES) or (isinstance(base, np.ndarray) and np.iscomplexobj(base)):
                if e[2] == "re":
                    return base.real
                if e[2] == "im":
                    return base.imag
            if isinstance(base, np.ndarray) and base.dtype == object:
                out = [x.f[e[2]] for x in base.ravel(order="F")]
                return np.array(out).reshape(base.shape, order="F")
            raise FortranError(f"component {e[2]!r} of a non-derived value {type(base).__name__}")
        if k == "call":
            return self.ev_call(e, sc)
        raise FortranError(f"cannot evaluate {e!r}")

    def binop(self, op, a, b):
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            if (_is_int(a) or (isinstance(a, np.ndarray) and np.issubdtype(a.dtype, np.integer))) and \
               (_is_int(b) or (isinstance(b, np.ndarray) and np.issubdtype(b.dtype, np.integer))):
                q = np.trunc(np.true_divide(a, b))
                return q.astype(np.int64) if isinstance(q, np.ndarray) else int(q)
            return a / b
        if op == "**":
            if _is_int(a) and _is_int(b):
                return int(a) ** int(b) if b >= 0 else int(int(a) ** int(b))
            if _is_int(b) and not isinstance(a, np.ndarray):
                return type(a)(a ** int(b)) if isinstance(a, (np.floating, np.complexfloating)) else a ** int(b)
            return a ** b
        if op == "==":
            return a == b
        if op == "/=":
            return a != b
        if op == "<":
            return a < b
        if op == "<=":
            return a <= b
        if op == ">":
            return a > b
        if op == ">=":
            return a >= b
        if op == ".and.":
            return np.logical_and(a, b) if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) else (bool(a) and bool(b))
        if op == ".or.":
            return np.logical_or(a, b) if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) else (bool(a) or bool(b))
        if op == ".eqv.":
            return bool(a) == bool(b)
        if op == ".neqv.":
            return bool(a) != bool(b)
        if op == "//":
            return str(a) + str(b)
        raise FortranError(f"operator {op!r}")

    def array_constructor(self, items, sc):
        out = []

        def add(v):
            if isinstance(v, np.ndarray):
                out.extend(v.ravel(order="F").tolist() if v.dtype == object else list(v.ravel(order="F")))
            else:
                out.append(v)
        for it in items:
            if it[0] == "implied":
                _, e, var, lo, hi, st = it
                lo_v, hi_v = int(self.ev(lo, sc)), int(self.ev(hi, sc))
                st_v = int(self.ev(st, sc)) if st is not None else 1
                saved = sc.vars.get(var, _MISSING)
                for i in range(lo_v, hi_v + (1 if st_v > 0 else -1), st_v):
                    sc.vars[var] = i
                    add(self.ev(e, sc))
                if saved is _MISSING:
                    sc.vars.pop(var, None)
                else:
                    sc.vars[var] = saved
            else:
                add(self.ev(it, sc))
        if out and isinstance(out[0], Inst):
            a = np.empty(len(out), dtype=object)
            for i, v in enumerate(out):
                a[i] = v
            return a
        if out and all(_is_int(v) for v in out):
            return np.array(out, dtype=np.int64)
        return np.array(out)

    # ---------------------------------------------------------------- indexing
    def make_index(self, args, sc, shape):
        idx, n_arr = [], 0
        for axis, (kw, a) in enumerate(args):
            if kw is not None:
                raise FortranError("keyword in an array subscript")
            if a[0] == "slice":
                lo = None if a[1] is None else int(self.ev(a[1], sc))
                hi = None if a[2] is None else int(self.ev(a[2], sc))
                st = 1 if a[3] is None else int(self.ev(a[3], sc))
                if st > 0:
                    if (lo is not None and lo < 1) or (hi is not None and hi > shape[axis]):
                        if not (hi is not None and lo is not None and hi < lo):
                            raise FortranError(f"section {lo}:{hi} out of bounds for extent {shape[axis]}")
                    idx.append(slice(None if lo is None else max(lo - 1, 0), None if hi is None else max(hi, 0), st))
                else:
                    start = shape[axis] - 1 if lo is None else lo - 1
                    stop = 0 if hi is None else hi - 1
                    idx.append(slice(start, stop - 1 if stop - 1 >= 0 else None, st))
            else:
                v = self.ev(a, sc)
                if isinstance(v, np.ndarray):
                    n_arr += 1
                    if v.dtype == np.bool_:
                        raise FortranError("logical subscripts are not Fortran")
                    if v.size and (v.min() < 1 or v.max() > shape[axis]):
                        raise FortranError(f"vector subscript out of bounds for extent {shape[axis]}")
                    idx.append(v.astype(np.int64) - 1)
                else:
                    if not _is_int(v):
                        raise FortranError(f"non-integer subscript {v!r}")
                    if v < 1 or v > shape[axis]:
                        raise FortranError(f"subscript {v} out of bounds for extent {shape[axis]}")
                    idx.append(int(v) - 1)
        if len(idx) != len(shape):
            raise FortranError(f"rank mismatch: {len(idx)} subscripts for rank {len(shape)}")
        if n_arr >= 2 or (n_arr == 1 and any(isinstance(i, slice) for i in idx) and len(idx) > 1 and
                          not isinstance(idx[0], np.ndarray) and any(isinstance(i, int) for i in idx)):
            # general case: open mesh (Fortran semantics = outer product of the subscripts)
            full, squeeze = [], []
            for axis, i in enumerate(idx):
                if isinstance(i, slice):
                    full.append(np.arange(shape[axis])[i])
                elif isinstance(i, np.ndarray):
                    full.append(i)
                else:
                    full.append(np.array([i]))
                    squeeze.append(axis)
            return ("mesh", np.ix_(*full), tuple(squeeze))
        return ("plain", tuple(idx), ())

    def index(self, arr, args, sc):
        if isinstance(arr, str):
            (_, a), = args
            lo = 1 if a[1] is None else int(self.ev(a[1], sc))
            hi = len(arr) if a[2] is None else int(self.ev(a[2], sc))
            return arr[lo - 1:hi]
        if not isinstance(arr, np.ndarray):
            raise FortranError(f"subscripted value is not an array ({type(arr).__name__})")
        mode, ix, squeeze = self.make_index(args, sc, arr.shape)
        v = arr[ix]
        if mode == "mesh" and squeeze:
            v = np.squeeze(v, axis=squeeze)
        return v

    # ---------------------------------------------------------------- assignment
    def copy_value(self, v):
        if isinstance(v, Inst):
            return copy.deepcopy(v)
        if isinstance(v, np.ndarray):
            return copy.deepcopy(v) if v.dtype == object else np.array(v, order="F")
        return v

    def inst_assign(self, dst, src):
        if dst is src:
            return
        found = self.find_binding(dst.tname, "assignment(=)")
        if found is not None and found[0] == "generic":              # defined assignment: generic :: assignment(=) => proc
            for b in found[1]:
                f2 = self.find_binding(dst.tname, b)
                proc = self.p.procs.get(f2[1][0]) if f2 and f2[0] == "specific" else None
                if proc is None:
                    continue
                bound = self.bind_args(proc, [(None, dst, None), (None, src, None)])
                if bound is not None:
                    self.run_proc(proc, bound, self.gscope)
                    return
        dst.tname = src.tname
        dst.f = copy.deepcopy(src.f)

    def store_into_array(self, arr, ix, value):
        if arr.dtype == object:
            cur = arr[ix]
            if isinstance(cur, Inst) and isinstance(value, Inst):
                self.inst_assign(cur, value)
            elif isinstance(cur, np.ndarray):
                vals = np.broadcast_to(value, cur.shape) if isinstance(value, np.ndarray) else None
                for j in np.ndindex(*cur.shape):
                    src = vals[j] if vals is not None else value
                    if isinstance(cur[j], Inst) and isinstance(src, Inst):
                        self.inst_assign(cur[j], src)
                    else:
                        cur[j] = self.copy_value(src)
            else:
                arr[ix] = self.copy_value(value)
        else:
            if isinstance(value, np.ndarray) and np.iscomplexobj(value) and not np.iscomplexobj(arr):
                value = value.real
            elif isinstance(value, _CPLX_TYPES) and not np.iscomplexobj(arr):
                value = value.real
            arr[ix] = value

    def assign_slot(self, cur, value, decl, setter):
        """Fortran intrinsic assignment into a variable / component whose current value is `cur`."""
        if isinstance(cur, np.ndarray):
            if isinstance(value, np.ndarray) and value.shape != cur.shape:
                if decl is not None and decl.allocatable:
                    setter(self._cast_array(value, decl))
                    return
                raise FortranError(f"shape mismatch in assignment: {cur.shape} <- {value.shape}")
            self.store_into_array(cur, Ellipsis, value)
            return
        if isinstance(cur, Inst) and isinstance(value, Inst):
            self.inst_assign(cur, value)
            return
        if isinstance(value, np.ndarray):
            if decl is not None and not decl.allocatable and decl.rank == 0:
                raise FortranError("array assigned to a scalar")
            setter(self._cast_array(value, decl))
            return
        if isinstance(value, Inst):
            setter(copy.deepcopy(value))
            return
        if decl is not None and decl.rank > 0 and cur is None:
            raise FortranError(f"scalar assigned to the unallocated array {decl.name}")
        setter(self.coerce(decl.ts, value) if decl is not None else value)

    def _cast_array(self, value, decl):
        if value.dtype == object:
            return copy.deepcopy(value)
        dt = decl.ts.dtype() if decl is not None else value.dtype
        if dt is object:
            dt = value.dtype
        if np.iscomplexobj(value) and not np.issubdtype(dt, np.complexfloating):
            value = value.real
        return np.array(value, dtype=dt, order="F")

    def assign(self, target, value, sc):
        k = target[0]
        if k == "name":
            nm = target[1]
            if nm in sc.vars:
                holder, decl = sc.vars, sc.decls.get(nm)
            elif nm in self.p.globals:
                holder, decl = self.p.globals, self.p.global_decls.get(nm)
            else:
                raise FortranError(f"assignment to the undeclared name {nm!r}")
            if holder[nm] is ABSENT:
                raise FortranError(f"assignment to the absent optional argument {nm!r}")
            self.assign_slot(holder[nm], value, decl, lambda v: holder.__setitem__(nm, v))
        elif k == "comp":
            obj = self.ev(target[1], sc)
            if isinstance(obj, np.ndarray) and np.iscomplexobj(obj) and target[2] in ("re", "im"):
                (obj.real if target[2] == "re" else obj.imag)[...] = value          # z%re = ... on a complex array (a view)
                return
            if not isinstance(obj, Inst):
                raise FortranError("component assignment on a non-derived value")
            decl = self.component_decl(obj.tname, target[2])
            if target[2] not in obj.f and decl is None:
                raise FortranError(f"type {obj.tname} has no component {target[2]!r}")
            self.assign_slot(obj.f.get(target[2]), value, decl, lambda v: obj.f.__setitem__(target[2], v))
        elif k == "call":
            arr = self.ev(target[1], sc)
            if isinstance(arr, np.ndarray):
                mode, ix, _ = self.make_index(target[2], sc, arr.shape)
                if mode == "mesh":
                    if isinstance(value, np.ndarray):
                        value = value.reshape(arr[ix].shape)
                self.store_into_array(arr, ix, value)
            else:
                raise FortranError(f"assignment to a subscripted non-array ({type(arr).__name__})")
        else:
            raise FortranError(f"bad assignment target {target!r}")

    def raw_store(self, target, value, sc):
        """rebind (no copy): used for argument copy-out and allocate"""
        k = target[0]
        if k == "name":
            nm = target[1]
            holder = sc.vars if nm in sc.vars else self.p.globals
            d = sc.decls.get(nm) if nm in sc.vars else self.p.global_decls.get(nm)
            if d is not None and d.rank == 0 and not isinstance(value, (np.ndarray, Inst)):
                value = self.coerce(d.ts, value)
            holder[nm] = value
        elif k == "comp":
            obj = self.ev(target[1], sc)
            if isinstance(obj, _CPLX_TYPES) and target[2] in ("re", "im"):
                new = type(obj)(complex(value, obj.imag) if target[2] == "re" else complex(obj.real, value))
                self.raw_store(target[1], new, sc)
                return
            obj.f[target[2]] = value
        elif k == "call":
            arr = self.ev(target[1], sc)
            mode, ix, _ = self.make_index(target[2], sc, arr.shape)
            if arr.dtype == object:
                arr[ix] = value
            else:
                self.store_into_array(arr, ix, value)
        elif k == "paren":
            return
        else:
            raise FortranError(f"bad store target {target!r}")

    @staticmethod
    def is_lvalue(ast):
        return ast[0] in ("name", "comp") or (ast[0] == "call" and ast[1][0] in ("name", "comp"))

    # ---------------------------------------------------------------- calls
    def ev_call(self, e, sc, want_result=True):
        base, args = e[1], e[2]
        if base[0] == "name":
            nm = base[1]
            if self.has_var(nm, sc):
                v = self.lookup(nm, sc)
                if isinstance(v, tuple) and v and v[0] == "procref":
                    return self.call_named(v[1], args, sc, want_result)
                return self.index(v, args, sc)
            return self.call_named(nm, args, sc, want_result)
        if base[0] == "comp":
            obj = self.ev(base[1], sc)
            if isinstance(obj, Inst):
                if base[2] in obj.f:
                    return self.index(obj.f[base[2]], args, sc)
                return self.call_bound(obj, base[1], base[2], args, sc, want_result)
            if obj is None:
                raise FortranError(f"type-bound reference {base[2]!r} on an unallocated object")
            raise FortranError(f"'%{base[2]}(...)' on a non-derived value")
        return self.index(self.ev(base, sc), args, sc)

    def eval_actuals(self, args, sc):
        out = []
        for kw, a in args:
            if a[0] == "slice":
                raise FortranError("array section syntax in a procedure reference")
            out.append((kw, self.ev(a, sc), a if self.is_lvalue(a) else None))
        return out

    def call_named(self, nm, args, sc, want_result):
        if sc.proc is not None and nm in sc.proc.internals:
            nm = sc.proc.internals[nm]
        if nm == "c_loc":
            v = self.ev(args[0][1], sc)
            return CPtr(v) if isinstance(v, np.ndarray) else CPtr(ScalarRef(self, sc, args[0][1]))
        if nm == "c_funloc":
            return CPtr(self.ev(args[0][1], sc))
        if nm == "present":
            return self.ev(args[0][1], sc) is not ABSENT
        if nm == "allocated":
            return self.ev(args[0][1], sc) is not None
        actuals = self.eval_actuals(args, sc)
        if nm in self.natives:
            return self.call_native(nm, self.natives[nm], actuals, sc)
        if nm in self.p.generics:
            try:
                proc = self.resolve_generic(nm, self.p.generics[nm], actuals)
            except FortranError:
                if nm not in self.p.types:
                    raise
                proc = None                     # generic named like a type: no specific matches -> the structure constructor
        if nm in self.p.generics and proc is not None:
            if isinstance(proc, str):
                return self.call_native(proc, self.natives[proc], actuals, sc)
            return self.call_proc(proc, actuals, sc)
        if nm in self.p.procs:
            return self.call_proc(self.p.procs[nm], actuals, sc)
        if nm in self.p.types:                          # structure constructor
            inst = self.new_inst(nm)
            names = [d for td in reversed(self.type_chain(nm)) for d in td.components]
            pos = 0
            for kw, v, _ in actuals:
                key = kw if kw is not None else names[pos]
                if kw is None:
                    pos += 1
                d = self.component_decl(nm, key)
                inst.f[key] = self._cast_array(v, d) if isinstance(v, np.ndarray) else self.copy_value(self.coerce(d.ts, v))
            return inst
        if nm in self.intrinsics:
            vals = [v for kw, v, _ in actuals if kw is None]
            kws = {kw: v for kw, v, _ in actuals if kw is not None}
            return self.intrinsics[nm](*vals, **kws)
        raise FortranError(f"unknown procedure {nm!r} (load its file or register a native)")

    def call_native(self, nm, fn, actuals, sc):
        vals = [v for kw, v, _ in actuals if kw is None]
        kws = {kw: v for kw, v, _ in actuals if kw is not None}
        res = fn(self, *vals, **kws)
        if isinstance(res, tuple) and res and res[0] == "__out__":
            pos = [a for a in actuals if a[0] is None]
            for key, val in res[1].items():
                a = pos[key] if isinstance(key, int) else next((x for x in actuals if x[0] == key), None)
                if a is not None and a[2] is not None:
                    self.raw_store(a[2], val, sc)
            return res[2] if len(res) > 2 else None
        return res

    def proc_by_name(self, nm):
        return self.p.procs.get(nm) or self.p.signatures.get(nm)

    def resolve_generic(self, gname, specifics, actuals, skip_dummy=None):
        matches = []
        for sp in specifics:
            if sp in self.natives and sp not in self.p.procs:
                continue
            proc = self.proc_by_name(sp)
            if proc is None:
                continue
            if self.bind_args(proc, actuals, skip_dummy, check_only=True) is not None:
                matches.append(proc)
        if len(matches) == 1:
            m = matches[0]
            if m.name not in self.p.procs:
                if m.name in self.natives:
                    return m.name
                raise FortranError(f"{gname}: specific {m.name} has an interface but no body loaded")
            return m
        if not matches:
            desc = [(kw, _kind_of_value(v)) for kw, v, _ in actuals]
            raise FortranError(f"no specific procedure of generic {gname!r} matches {desc}")
        raise FortranError(f"ambiguous generic {gname!r}: {[m.name for m in matches]}")

    def value_matches(self, d, v, elemental):
        if v is ABSENT or v is None:
            return True
        if isinstance(v, tuple) and v and v[0] == "procref":
            return d.ts.base == "procedure"
        if isinstance(v, CPtr):
            return d.ts.base == "type" and d.ts.tname in ("c_ptr", "c_funptr") and d.rank == 0
        base, kind, rank = _kind_of_value(v)
        if rank != d.rank and not (elemental and d.rank == 0):
            return False
        b = d.ts.base
        if b in ("integer", "logical", "character"):
            return base == b
        if b in ("real", "complex"):
            return base == b and kind == (d.ts.kind or "sp")
        if b in ("type", "class"):
            if base != "derived":
                return False
            if kind is None or d.ts.tname == "*":
                return True
            return kind == d.ts.tname if b == "type" else self.isa(kind, d.ts.tname)
        return False

    def bind_args(self, proc, actuals, skip_dummy=None, check_only=False):
        """-> {dummy: (value, lvalue_ast)} or None when the actual arguments do not fit the interface"""
        dummies = [a for a in proc.args if a != skip_dummy]
        bound = {}
        pos = 0
        for kw, v, lv in actuals:
            if kw is None:
                if pos >= len(dummies):
                    return None
                name = dummies[pos]
                pos += 1
            else:
                if kw not in dummies:
                    return None
                name = kw
            if name in bound:
                return None
            d = proc.decls.get(name)
            if d is None:
                return None
            if not self.value_matches(d, v, proc.elemental):
                return None
            bound[name] = (v, lv)
        for name in dummies:
            if name not in bound:
                d = proc.decls.get(name)
                if d is None or not d.optional:
                    return None
                bound[name] = (ABSENT, None)
        return bound

    def call_bound(self, obj, obj_ast, bname, args, sc, want_result):
        if obj.tname not in self.p.types:
            return None                                  # type from a module that is not loaded (timers): no-op
        actuals = self.eval_actuals(args, sc)
        found = self.find_binding(obj.tname, bname)
        if found is None:
            raise FortranError(f"type {obj.tname} has no binding {bname!r}")
        obj_lv = obj_ast if self.is_lvalue(obj_ast) else None
        if found[0] == "generic":
            cands = []
            for b in found[1]:
                f2 = self.find_binding(obj.tname, b)
                if f2 and f2[0] == "specific":
                    cands.append(f2[1])
        else:
            cands = [found[1]]
        errors = []
        for target, passname, _, nopass in cands:
            proc = self.p.procs.get(target)
            if proc is None:
                if target in self.natives:
                    return self.call_native(target, self.natives[target], [(None, obj, obj_lv)] + actuals, sc)
                errors.append(f"{target}: body not loaded")
                continue
            if nopass:
                bound = self.bind_args(proc, actuals)
            else:
                pname = passname or proc.args[0]
                bound = self.bind_args(proc, actuals, skip_dummy=pname)
                if bound is not None:
                    bound[pname] = (obj, obj_lv)
            if bound is not None:
                return self.run_proc(proc, bound, sc)
            errors.append(f"{target}: arguments do not match")
        raise FortranError(f"{obj.tname}%{bname}: {errors}")

    def call_proc(self, proc, actuals, sc):
        bound = self.bind_args(proc, actuals)
        if bound is None:
            desc = [(kw, _kind_of_value(v)) for kw, v, _ in actuals]
            raise FortranError(f"{proc.name}: actual arguments {desc} do not match dummies {proc.args}")
        return self.run_proc(proc, bound, sc)

    def reset_intent_out(self, v):
        if isinstance(v, Inst):
            for td in self.type_chain(v.tname):
                for d in td.components.values():
                    if d.allocatable:
                        v.f[d.name] = None
                    elif d.init is not None and not d.dims:
                        v.f[d.name] = self.coerce(d.ts, self.ev(d.init, self.gscope))
                    elif d.init is not None and isinstance(v.f.get(d.name), np.ndarray):
                        v.f[d.name][...] = self.ev(d.init, self.gscope)          # default-initialised array component
        elif isinstance(v, np.ndarray) and v.dtype == object:
            for x in v.ravel():
                self.reset_intent_out(x)

    def run_proc(self, proc, bound, caller_sc):
        # elemental reference with array actuals
        if proc.elemental:
            arrs = [(n, v) for n, (v, lv) in bound.items()
                    if isinstance(v, np.ndarray) and proc.decls[n].rank == 0]
            if arrs:
                shape = arrs[0][1].shape
                if proc.kind == "function" or proc.result:
                    res = None
                for ix in np.ndindex(*shape):
                    b2 = {}
                    for n, (v, lv) in bound.items():
                        if isinstance(v, np.ndarray) and proc.decls[n].rank == 0:
                            if v.shape != shape:
                                raise FortranError(f"{proc.name}: elemental arguments of different shapes")
                            b2[n] = (v[ix], None)
                        else:
                            b2[n] = (v, None)
                    r = self.run_proc_scalar(proc, b2, caller_sc)
                    if proc.result:
                        if res is None:
                            res = np.empty(shape, dtype=np.asarray(r).dtype, order="F")
                        res[ix] = r
                    else:
                        for n, (v, lv) in bound.items():
                            if isinstance(v, np.ndarray) and proc.decls[n].rank == 0 and v.dtype != object \
                                    and proc.decls[n].intent != "in":
                                v[ix] = self._last_scope.vars[n]
                return res if proc.result else None
        return self.run_proc_scalar(proc, bound, caller_sc)

    def run_proc_scalar(self, proc, bound, caller_sc):
        hook = self.hooks.get(proc.name)
        if hook is not None:
            # emulates a one-line patch at the top of the procedure:   if (<hook>(<same dummies>)) return
            tproc = self.p.procs[hook]
            if len(tproc.args) != len(proc.args):
                raise FortranError(f"hook {hook}: {len(tproc.args)} dummies for the {len(proc.args)} of {proc.name}")
            if bool(self.run_proc(tproc, {tn: bound[rn] for tn, rn in zip(tproc.args, proc.args)}, caller_sc)):
                self.hook_hits[proc.name] = self.hook_hits.get(proc.name, 0) + 1
                return None
        self.called.add(proc.name)
        sc = Scope(proc)
        for n, (v, lv) in bound.items():
            d = proc.decls[n]
            if d.intent == "out" and v is not ABSENT:
                if d.allocatable:
                    v = None
                else:
                    self.reset_intent_out(v)
            elif d.intent == "in" and d.rank == 0 and not isinstance(v, (np.ndarray, Inst)) and v is not ABSENT and v is not None:
                v = self.coerce(d.ts, v) if d.ts.base in ("integer", "real", "complex") and _kind_of_value(v)[0] == d.ts.base else v
            sc.vars[n] = v
        if proc.result and proc.result_ts is not None and proc.result not in proc.decls:
            sc.vars[proc.result] = self._undefined_scalar(proc.result_ts)
            sc.decls[proc.result] = Decl(proc.result, proc.result_ts, None, {}, None)
        self.call_depth += 1
        if self.call_depth > 200:
            raise FortranError("call depth > 200")
        if self.trace:
            print("  " * self.call_depth + proc.name)
        try:
            self.exec_block(proc.body, sc)
        except _Return:
            pass
        finally:
            self.call_depth -= 1
        self._last_scope = sc
        # copy-out
        for n, (v, lv) in bound.items():
            if lv is None or v is ABSENT:
                continue
            d = proc.decls[n]
            if d.intent == "in":
                continue
            new = sc.vars[n]
            if new is v and isinstance(new, (np.ndarray, Inst)):
                continue
            if new is v and new is None:
                continue
            if isinstance(v, np.ndarray) and isinstance(new, np.ndarray) and not d.allocatable:
                continue
            self.raw_store(lv, new, caller_sc)
        if proc.result:
            return sc.vars[proc.result]
        return None

    # ---------------------------------------------------------------- statements
    def exec_block(self, stmts, sc):
        for st in stmts:
            try:
                self.exec_stmt(st, sc)
            except (_Exit, _Cycle, _Return, StopError):
                raise
            except FortranError as exc:
                w = st[-1]
                if not getattr(exc, "located", False):
                    exc.located = True
                    exc.args = (f"{exc.args[0]}\n    at {w[0]}:{w[1]}: {w[2]}",) + exc.args[1:]
                raise
            except Exception as exc:     # numpy errors etc.
                w = st[-1]
                err = FortranError(f"{type(exc).__name__}: {exc}\n    at {w[0]}:{w[1]}: {w[2]}")
                err.located = True
                raise err from exc

    def exec_stmt(self, st, sc):
        k = st[0]
        if k == "assign":
            self.assign(st[1], self.ev(st[2], sc), sc)
        elif k == "callsub":
            self.ev_call(st[1], sc, want_result=False)
        elif k == "decl":
            for d in st[1]:
                self._declare(d, sc)
        elif k == "if":
            for cond, blk in st[1]:
                if bool(self.ev(cond, sc)):
                    self.exec_block(blk, sc)
                    return
            if st[2] is not None:
                self.exec_block(st[2], sc)
        elif k == "do":
            self.exec_do(st, sc)
        elif k == "select_type":
            _, sel, assoc, guards, _w = st
            v = self.ev(sel, sc)
            if assoc is not None and sel != ("name", assoc):
                sc.vars[assoc] = v
            chosen = None
            tn = v.tname if isinstance(v, Inst) else None
            if isinstance(v, np.ndarray) and v.dtype == object:           # select type on an array: the type of its elements
                first = next((x for x in v.ravel() if isinstance(x, Inst)), None)
                tn = first.tname if first is not None else None
            for (gk, gname), blk in guards:
                if gk == "type" and tn == gname:
                    chosen = blk
                    break
            if chosen is None:
                best = None
                for (gk, gname), blk in guards:
                    if gk == "class" and tn is not None and self.isa(tn, gname):
                        depth = [td.name for td in self.type_chain(tn)].index(gname)
                        if best is None or depth < best[0]:
                            best = (depth, blk)
                if best:
                    chosen = best[1]
            if chosen is None:
                for (gk, gname), blk in guards:
                    if gk == "default":
                        chosen = blk
            if chosen is not None:
                self.exec_block(chosen, sc)
        elif k == "select_case":
            _, sel, cases, _w = st
            v = self.ev(sel, sc)
            default = None
            for vals, blk in cases:
                if vals is None:
                    default = blk
                    continue
                for c in vals:
                    if c[0] == "slice":
                        lo = self.ev(c[1], sc) if c[1] is not None else None
                        hi = self.ev(c[2], sc) if c[2] is not None else None
                        hit = (lo is None or v >= lo) and (hi is None or v <= hi)
                    else:
                        cv = self.ev(c, sc)
                        hit = (v.strip() == cv.strip()) if isinstance(v, str) else v == cv
                    if hit:
                        self.exec_block(blk, sc)
                        return
            if default is not None:
                self.exec_block(default, sc)
        elif k == "block":
            self.exec_block(st[1], sc)
        elif k == "associate":
            for nm, e in st[1]:
                sc.vars[nm] = self.ev(e, sc)
            self.exec_block(st[2], sc)
        elif k == "allocate":
            self.exec_allocate(st, sc)
        elif k == "deallocate":
            for it in st[1]:
                self.raw_store(it, None, sc)
        elif k in ("io", "nop"):
            pass
        elif k == "ptr_assign":
            tgt, val = st[1], self.ev(st[2], sc)
            if tgt[0] == "call" and tgt[1][0] == "name" and all(a[0] == "slice" for _, a in tgt[2]):
                # bounds-remapping pointer assignment  p(1:k, 1:1) => e(:k): a reshaped VIEW of the target
                shape = []
                for _, a in tgt[2]:
                    lo, hi = int(self.ev(a[1], sc)), int(self.ev(a[2], sc))
                    if lo != 1:
                        raise FortranError("pointer bounds remapping with a lower bound other than 1")
                    shape.append(hi)
                view = val.reshape(tuple(shape), order="F")
                if view.size and not np.shares_memory(view, val):
                    raise FortranError("pointer bounds remapping of a non-contiguous target")
                sc.vars[tgt[1][1]] = view
            else:
                self.raw_store(tgt, val, sc)
        elif k == "return":
            raise _Return()
        elif k == "exit":
            raise _Exit(st[1])
        elif k == "cycle":
            raise _Cycle(st[1])
        elif k == "stop":
            raise StopError(st[1])
        else:
            raise FortranError(f"statement kind {k!r}")

    def exec_do(self, st, sc):
        _, hdr, blk, label, _w = st

        def body():
            """-> True to leave the loop"""
            try:
                self.exec_block(blk, sc)
            except _Cycle as c:
                if c.label not in (None, label):
                    raise
            except _Exit as x:
                if x.label not in (None, label):
                    raise
                return True
            return False
        if hdr[0] == "count":
            _, var, lo, hi, step = hdr
            lo_v, hi_v = int(self.ev(lo, sc)), int(self.ev(hi, sc))
            st_v = 1 if step is None else int(self.ev(step, sc))
            n = max(0, (hi_v - lo_v + st_v) // st_v)
            i = lo_v
            sc.vars[var] = i
            for _ in range(n):
                sc.vars[var] = i
                if body():
                    return
                i += st_v
                sc.vars[var] = i             # as Fortran: after normal completion the index holds the first value beyond the range
        elif hdr[0] == "while":
            while bool(self.ev(hdr[1], sc)):
                if body():
                    return
        else:
            while True:
                if body():
                    return

    def exec_allocate(self, st, sc):
        _, items, opts, _w = st
        source = self.ev(opts["source"], sc) if "source" in opts else None
        mold = self.ev(opts["mold"], sc) if "mold" in opts else None
        for it in items:
            if it[0] == "call":
                target = it[1]
                shape = []
                for kw, a in it[2]:
                    if a[0] == "slice":
                        lo, hi = int(self.ev(a[1], sc)), int(self.ev(a[2], sc))
                        if lo != 1:
                            raise FortranError("allocate with a lower bound other than 1")
                        shape.append(max(0, hi))
                    else:
                        shape.append(max(0, int(self.ev(a, sc))))
                shape = tuple(shape)
                decl = self.decl_of(target, sc)
                src = source if source is not None else mold
                if isinstance(src, Inst):
                    arr = np.empty(shape, dtype=object, order="F")
                    for ix in np.ndindex(*shape):
                        arr[ix] = copy.deepcopy(src)
                elif isinstance(src, np.ndarray) and src.dtype == object:
                    arr = np.empty(shape, dtype=object, order="F")
                    flat = src.ravel(order="F")
                    for j, ix in enumerate(np.ndindex(*shape[::-1])):
                        arr[ix[::-1]] = copy.deepcopy(flat[j if source is not None else 0])
                elif decl is None:
                    raise FortranError("allocate: unknown declaration")
                elif source is not None:
                    arr = self._new_array(decl.ts, shape, fill=0)
                    arr[...] = source
                else:
                    arr = self._new_array(decl.ts, shape)
                self.raw_store(target, arr, sc)
            else:
                src = source if source is not None else mold
                if src is None:
                    decl = self.decl_of(it, sc)
                    if decl is not None and decl.ts.base in ("type", "class") and decl.ts.tname in self.p.types:
                        self.raw_store(it, self.new_inst(decl.ts.tname), sc)
                        continue
                    raise FortranError("allocate of a scalar without source= / mold=")
                self.raw_store(it, self.copy_value(src), sc)
        if "stat" in opts:
            self.assign(opts["stat"], 0, sc)

    def decl_of(self, target, sc):
        if target[0] == "name":
            return sc.decls.get(target[1]) or self.p.global_decls.get(target[1])
        if target[0] == "comp":
            obj = self.ev(target[1], sc)
            return self.component_decl(obj.tname, target[2])
        return None

    # ---------------------------------------------------------------- public helpers
    def call(self, name, *args, **kwargs):
        """call a reference procedure (specific or generic) from Python; arrays / Inst are passed by reference.
        Returns the function result, or for subroutines a dict of the final values of scalar dummies."""
        sc = Scope(None)
        actuals = []
        for i, a in enumerate(args):
            sc.vars[f"__a{i}"] = a
            actuals.append((None, ("name", f"__a{i}")))
        for k, a in kwargs.items():
            sc.vars[f"__k{k}"] = a
            actuals.append((k.lower(), ("name", f"__k{k}")))
        res = self.ev_call(("call", ("name", name.lower()), actuals), sc)
        outs = {i: sc.vars[f"__a{i}"] for i in range(len(args))}
        outs.update({k: sc.vars[f"__k{k}"] for k in kwargs})
        return res, outs


# ------------------------------------------------------------------------------------------------ intrinsics
def _kind_cast(v, kind, cplx):
    if kind is None:
        return v
    if cplx:
        return (np.complex128 if kind == "dp" else np.complex64)(v) if not isinstance(v, np.ndarray) \
            else v.astype(np.complex128 if kind == "dp" else np.complex64)
    return (np.float64 if kind == "dp" else np.float32)(v) if not isinstance(v, np.ndarray) \
        else v.astype(np.float64 if kind == "dp" else np.float32)


def _i_real(a, kind=None):
    if isinstance(a, np.ndarray):
        r = a.real if np.iscomplexobj(a) else a
        if kind is None:
            return r.astype(np.float32) if np.issubdtype(r.dtype, np.integer) else np.array(r)
        return _kind_cast(r, kind, False)
    r = a.real if isinstance(a, _CPLX_TYPES) else a
    if kind is None:
        if _is_int(r):
            return np.float32(r)
        return r if isinstance(r, np.floating) else np.float64(r)
    return _kind_cast(r, kind, False)


def _i_cmplx(a, b=0, kind=None):
    k = kind or "sp"                       # cmplx without kind= returns default (single) complex -- as in Fortran
    if isinstance(a, _CPLX_TYPES) or (isinstance(a, np.ndarray) and np.iscomplexobj(a)):
        return _kind_cast(a, k, True)
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return _kind_cast(np.asarray(a) + 1j * np.asarray(b), k, True)
    return _kind_cast(complex(float(a), float(b)), k, True)


def _i_size(a, dim=None):
    if not isinstance(a, np.ndarray):
        raise FortranError(f"size() of a non-array ({type(a).__name__})")
    return int(a.size) if dim is None else int(a.shape[int(dim) - 1])


def _i_abs(a):
    return np.abs(a)


def _i_sqrt(a):
    with np.errstate(invalid="ignore"):
        return np.sqrt(a)


def _reduce(fn):
    def f(a, dim=None, mask=None):
        if mask is not None:
            a = a[mask]
        if dim is not None:
            return fn(a, axis=int(dim) - 1)
        return fn(a)
    return f


def _i_minmax(fn):
    def f(*a):
        r = a[0]
        for x in a[1:]:
            r = fn(r, x)
        return r
    return f


def _i_loc(fn):
    def f(a, dim=None, mask=None):
        if a.ndim == 1:
            if mask is not None:
                a = np.where(mask, a, -np.inf if fn is np.argmax else np.inf)
            r = int(fn(a)) + 1
            return r if dim is not None else np.array([r], dtype=np.int64)
        ix = np.unravel_index(fn(a), a.shape)
        return np.array([i + 1 for i in ix], dtype=np.int64)
    return f


def _seq_sum(v):
    """left-to-right summation in the precision of the array, as a compiled loop without -ffast-math does it (numpy's own sum
    is pairwise); add.accumulate is sequential by definition"""
    v = np.asarray(v)
    if v.ndim != 1 or v.size == 0:
        return np.sum(v)
    return np.add.accumulate(v)[-1]


def _i_dot_product(a, b):
    if np.iscomplexobj(a):
        return _seq_sum(np.conj(a) * b)
    return _seq_sum(a * b)


def _i_merge(t, f, mask):
    if isinstance(mask, np.ndarray):
        return np.where(mask, t, f)
    return t if mask else f


def _i_sign(a, b):
    return np.copysign(np.abs(a), b) if not _is_int(a) else (abs(a) if b >= 0 else -abs(a))


def _i_epsilon(x):
    return np.finfo(np.asarray(x).dtype).eps.astype(np.asarray(x).dtype)


def _i_huge(x):
    if _is_int(x):
        return 2 ** 31 - 1
    return np.finfo(np.asarray(x).dtype).max


def _i_tiny(x):
    return np.finfo(np.asarray(x).dtype).tiny


def _i_precision(x):
    return 15 if np.asarray(x).dtype in (np.float64, np.complex128) else 6


def _i_reshape(a, shape, order=None, pad=None):
    return np.reshape(np.asarray(a), tuple(int(s) for s in shape), order="F")


def _i_int(a, kind=None):
    if isinstance(a, np.ndarray):
        return np.trunc(a.real).astype(np.int64)
    return int(a.real) if isinstance(a, _CPLX_TYPES) else int(a)


def _i_mod(a, p):
    if _is_int(a) and _is_int(p):
        return int(np.fmod(a, p))
    return np.fmod(a, p)


def _i_matmul(a, b):
    r = a @ b
    return np.asfortranarray(r) if isinstance(r, np.ndarray) and r.ndim == 2 else r


def _i_transfer(source, mold, size=None):
    """transfer(character string, character array mold): the one use the shim makes of it"""
    if isinstance(source, str) and isinstance(mold, np.ndarray):
        n = mold.size if size is None else int(size)
        a = np.empty(n, dtype=object)
        for i in range(n):
            a[i] = source[i] if i < len(source) else " "
        return a
    raise FortranError("transfer: only string -> character array is provided")


def _i_kind(x):
    return "dp" if np.asarray(x).dtype in (np.float64, np.complex128) else "sp"


INTRINSICS = {
    "size": _i_size,
    "shape": lambda a: np.array(a.shape, dtype=np.int64),
    "lbound": lambda a, dim=None: 1 if dim is not None else np.ones(a.ndim, dtype=np.int64),
    "ubound": lambda a, dim=None: int(a.shape[int(dim) - 1]) if dim is not None else np.array(a.shape, dtype=np.int64),
    "abs": _i_abs, "sqrt": _i_sqrt,
    "real": _i_real, "dble": lambda a: _i_real(a, "dp"), "aimag": lambda a: a.imag, "conjg": np.conj, "cmplx": _i_cmplx,
    "int": _i_int, "nint": lambda a, kind=None: int(np.rint(a)), "floor": lambda a: int(np.floor(a)),
    "ceiling": lambda a: int(np.ceil(a)), "mod": _i_mod, "modulo": lambda a, p: a % p,
    "min": _i_minmax(lambda x, y: np.minimum(x, y) if isinstance(x, np.ndarray) or isinstance(y, np.ndarray) else (x if x <= y else y)),
    "max": _i_minmax(lambda x, y: np.maximum(x, y) if isinstance(x, np.ndarray) or isinstance(y, np.ndarray) else (x if x >= y else y)),
    "minval": _reduce(np.min), "maxval": _reduce(np.max), "product": _reduce(np.prod),
    "sum": lambda a, dim=None, mask=None: _seq_sum(a) if dim is None and mask is None and np.ndim(a) == 1 else _reduce(np.sum)(a, dim, mask),
    "any": _reduce(np.any), "all": _reduce(np.all), "count": lambda a, dim=None: int(np.count_nonzero(a)),
    "maxloc": _i_loc(np.argmax), "minloc": _i_loc(np.argmin),
    "matmul": _i_matmul, "transpose": lambda a: np.asfortranarray(a.T), "dot_product": _i_dot_product,
    "norm2": lambda a: np.sqrt(np.sum(np.abs(a) ** 2)),
    "isnan": np.isnan, "ieee_is_nan": np.isnan,
    "exp": np.exp, "log": np.log, "log10": np.log10, "sin": np.sin, "cos": np.cos, "tan": np.tan, "atan": np.arctan,
    "atan2": np.arctan2, "acos": np.arccos, "asin": np.arcsin, "tanh": np.tanh, "sinh": np.sinh, "cosh": np.cosh,
    "sign": _i_sign, "merge": _i_merge, "epsilon": _i_epsilon, "huge": _i_huge, "tiny": _i_tiny, "precision": _i_precision,
    "selected_real_kind": lambda p=6, r=37: "sp" if p <= 6 else "dp",
    "kind": _i_kind, "reshape": _i_reshape, "transfer": _i_transfer,
    "trim": lambda s: s.rstrip(), "adjustl": lambda s: s.lstrip(), "len": len, "len_trim": lambda s: len(s.rstrip()),
    "to_lower": lambda s: s.lower(),
    "spread": lambda a, dim, ncopies: np.repeat(np.expand_dims(np.asarray(a), int(dim) - 1), int(ncopies), axis=int(dim) - 1),
}


# ------------------------------------------------------------------------------------------------ natives
# What the reference takes from OUTSIDE its own sources: fortran-stdlib (optval, BLAS wrappers, dense linear algebra on small
# host matrices) and its logging / timing / error helpers.  Signature: fn(interp, *positional, **keywords).
def _n_noop(interp, *a, **k):
    return None


def _n_stop_error(interp, msg="", *a, **k):
    raise StopError(str(msg))


def _n_check_info(interp, info, origin="", *a, **k):
    if info < 0:
        raise StopError(f"check_info: {origin} returned info = {info}")


def _n_optval(interp, x, default):
    return default if x is ABSENT else x


def _n_assert_shape(interp, a, shp, *rest, **k):
    if tuple(int(s) for s in np.atleast_1d(shp)) != a.shape:
        raise StopError(f"assert_shape: {a.shape} vs {tuple(shp)}")


def _n_scal(interp, n, a, x, incx):
    x[:n] *= a


def _n_axpy(interp, n, a, x, incx, y, incy):
    y[:n] += a * x[:n]


def _n_dot(interp, n, x, incx, y, incy):
    return x.dtype.type(_seq_sum(x[:n] * y[:n]))             # reference BLAS xDOT: one running sum


def _n_dotc(interp, n, x, incx, y, incy):
    return x.dtype.type(_seq_sum(np.conj(x[:n]) * y[:n]))


def _n_nrm2(interp, n, x, incx):
    return np.sqrt(np.sum(np.abs(x[:n]) ** 2))


def _n_gemv(interp, trans, m, n, alpha, a, lda, x, incx, beta, y, incy):
    t = trans.upper()
    A = a[:m, :n]
    op = A if t == "N" else (A.T if t == "T" else A.conj().T)
    ylen = m if t == "N" else n
    xlen = n if t == "N" else m
    prod = op @ x[:xlen]
    if beta == 0:
        y[:ylen] = alpha * prod
    else:
        y[:ylen] = alpha * prod + beta * y[:ylen]


def _n_eye(interp, n, m=None, mold=None):
    dt = np.asarray(mold).dtype if mold is not None else np.float64
    return np.asfortranarray(np.eye(int(n), int(m) if m is not None else int(n), dtype=dt))


def _n_mnorm(interp, a, order="fro", *rest, **k):
    o = str(order).strip().lower()
    if o in ("fro", "euclidean", "f"):
        return np.abs(a).dtype.type(np.sqrt(np.sum(np.abs(a) ** 2)))
    if o in ("1",):
        return np.linalg.norm(a, 1)
    if o in ("inf",):
        return np.linalg.norm(a, np.inf)
    if o in ("2",):
        return np.linalg.norm(a, 2)
    raise FortranError(f"mnorm order {order!r}")


def _n_norm(interp, a, order=2, *rest, **k):
    if isinstance(order, str):
        return _n_mnorm(interp, a, order)
    if int(order) == 2:
        return np.abs(a).dtype.type(np.sqrt(np.sum(np.abs(a) ** 2)))
    return np.linalg.norm(a.ravel(), int(order))


def _n_random_number(interp, x):
    rng = interp.rng
    if isinstance(x, np.ndarray):
        x[...] = rng.random(x.shape)
        return None
    return ("__out__", {0: rng.random()})


def _n_hermitian(interp, a):
    return np.asfortranarray(np.conj(a).T)


def _n_normal(interp, loc, scale):
    """stdlib_stats_distribution_normal rvs_normal(loc, scale): only used to (re)fill vectors with random numbers, so the
    stream is implementation-defined anyway (seeded numpy generator interp.rng)"""
    rng = interp.rng
    shape = np.shape(loc)
    dt = np.asarray(loc).dtype
    if np.issubdtype(dt, np.complexfloating):
        r = np.real(loc) + np.real(scale) * rng.standard_normal(shape) + 1j * (np.imag(loc) + np.imag(scale) * rng.standard_normal(shape))
    else:
        r = loc + scale * rng.standard_normal(shape)
    return np.asarray(r, dtype=dt) if shape else dt.type(r)


def _lapack(name, *arrays):
    from scipy.linalg import get_lapack_funcs
    return get_lapack_funcs(name, arrays)


def _n_lasr(interp, side, pivot, direct, m, n, c, s, a, lda):
    """LAPACK xLASR, the one variant the reference uses: SIDE = L, PIVOT = V, DIRECT = F (restated from the LAPACK source)"""
    if (side.upper(), pivot.upper(), direct.upper()) != ("L", "V", "F"):
        raise FortranError("lasr: only side=L, pivot=V, direct=F is provided")
    for j in range(m - 1):
        ct, st = c[j], s[j]
        if ct != 1 or st != 0:
            for i in range(n):
                temp = a[j + 1, i]
                a[j + 1, i] = ct * temp - st * a[j, i]
                a[j, i] = st * temp + ct * a[j, i]


def _n_lartg(interp, f, g, c, s, r):
    fn = _lapack("lartg", np.asarray(f))
    cs, sn, rr = fn(f, g)
    t = np.asarray(f).dtype.type
    return ("__out__", {2: t(cs), 3: t(sn), 4: t(rr)})


def _n_trtrs(interp, uplo, trans, diag, n, nrhs, a, lda, b, ldb, info):
    fn = _lapack("trtrs", a)
    x, inf = fn(a[:n, :n], b[:n, :nrhs], lower=int(uplo.upper() == "L"),
                trans={"N": 0, "T": 1, "C": 2}[trans.upper()], unitdiag=int(diag.upper() == "U"))
    b[:n, :nrhs] = x
    return ("__out__", {9: int(inf)})


def _n_geev(interp, jobvl, jobvr, n, a, lda, *rest):
    fn = _lapack("geev", a)
    if np.iscomplexobj(a):
        w, vl, ldvl, vr, ldvr, work, lwork, rwork, info = rest
        wv, vlv, vrv, inf = fn(a[:n, :n], compute_vl=int(jobvl.upper() == "V"), compute_vr=int(jobvr.upper() == "V"))
        w[:n] = wv
        pos_info = 13
    else:
        wr, wi, vl, ldvl, vr, ldvr, work, lwork, info = rest
        wrv, wiv, vlv, vrv, inf = fn(a[:n, :n], compute_vl=int(jobvl.upper() == "V"), compute_vr=int(jobvr.upper() == "V"))
        wr[:n], wi[:n] = wrv, wiv
        pos_info = 13
    if jobvr.upper() == "V":
        vr[:n, :n] = vrv
    if jobvl.upper() == "V":
        vl[:n, :n] = vlv
    return ("__out__", {pos_info: int(inf)})


def _n_eigh(interp, a, lambda_, vectors=None, upper_a=None, overwrite_a=None, err=None, **k):
    """stdlib_linalg eigh(A, lambda, vectors): xSYEV / xHEEV, lower triangle by default"""
    fn = _lapack("heev" if np.iscomplexobj(a) else "syev", a)
    lower = 0 if (upper_a is not None and upper_a is not ABSENT and upper_a) else 1
    w, v, inf = fn(a, compute_v=1, lower=lower)
    if inf != 0:
        raise StopError(f"eigh: info = {inf}")
    lambda_[...] = w
    if vectors is not None and vectors is not ABSENT:
        vectors[...] = v


def _n_svd(interp, a, s, u=None, vt=None, **k):
    """stdlib_linalg svd(A, s, u, vt): xGESDD"""
    fn = _lapack("gesdd", a)
    uu, ss, vvt, inf = fn(a, compute_uv=1, full_matrices=1)
    if inf != 0:
        raise StopError(f"svd: info = {inf}")
    s[...] = ss[:s.shape[0]]
    if u is not None and u is not ABSENT:
        u[...] = uu[:u.shape[0], :u.shape[1]]
    if vt is not None and vt is not ABSENT:
        vt[...] = vvt[:vt.shape[0], :vt.shape[1]]


def _n_schur(interp, a, t, z=None, eigvals=None, **k):
    """stdlib_linalg schur(A, T, Z, eigvals): xGEES (real kinds: real quasi-triangular form)"""
    import scipy.linalg as sla
    tt, zz = sla.schur(np.array(a), output="complex" if np.iscomplexobj(a) else "real")
    t[...] = tt
    if z is not None and z is not ABSENT:
        z[...] = zz
    if eigvals is not None and eigvals is not ABSENT:
        n = tt.shape[0]
        if np.iscomplexobj(tt):
            eigvals[...] = np.diag(tt)
        else:                                   # eigenvalues of the 1x1 / 2x2 diagonal blocks, as xGEES returns (wr, wi)
            w = np.zeros(n, dtype=np.complex128)
            i = 0
            while i < n:
                if i + 1 < n and tt[i + 1, i] != 0:
                    a11, a12, a21, a22 = tt[i, i], tt[i, i + 1], tt[i + 1, i], tt[i + 1, i + 1]
                    im = np.sqrt(abs(a12)) * np.sqrt(abs(a21))              # standardised block: a11 == a22, a12 * a21 < 0
                    w[i], w[i + 1] = complex(a11, im), complex(a22, -im)
                    i += 2
                else:
                    w[i] = tt[i, i]
                    i += 1
            eigvals[...] = w


def _n_trsen(interp, job, compq, select, n, t, ldt, q, ldq, *rest):
    fn = _lapack("trsen", t)
    sel = np.asarray(select, dtype=np.int32)
    if np.iscomplexobj(t):
        w, m, s, sep, work, lwork, info = rest
        ts, qs, wv, mm, ss, sp_, inf = fn(sel, np.array(t[:n, :n]), np.array(q[:n, :n]), job=job.upper(), wantq=int(compq.upper() == "V"))
        w[:n] = wv
        out = {9: int(mm), 10: ss, 11: sp_, 14: int(inf)}
    else:
        wr, wi, m, s, sep, work, lwork, iwork, liwork, info = rest
        ts, qs, wrv, wiv, mm, ss, sp_, inf = fn(sel, np.array(t[:n, :n]), np.array(q[:n, :n]), job=job.upper(), wantq=int(compq.upper() == "V"))
        wr[:n], wi[:n] = wrv, wiv
        out = {10: int(mm), 11: ss, 12: sp_, 17: int(inf)}
    t[:n, :n] = ts
    q[:n, :n] = qs
    return ("__out__", out)


def _n_expm(interp, a, *rest, **k):
    """stdlib_linalg expm (third-party, not in the reference tree): Pade approximant with scaling and squaring"""
    import scipy.linalg as sla
    return np.asfortranarray(sla.expm(np.asarray(a, dtype=np.complex128 if np.iscomplexobj(a) else np.float64)).astype(a.dtype))


def _n_sort_index(interp, array, index, work=None, iwork=None, reverse=None):
    """stdlib_sorting sort_index: STABLE sort of `array` in place, `index` (1-based) = the permutation applied"""
    rev = reverse is not None and reverse is not ABSENT and bool(reverse)
    key = np.array(array)
    if rev:
        # a stable DEcreasing sort keeps ties in their original order
        order = np.array(sorted(range(len(key)), key=lambda i: -key[i]), dtype=np.int64)
    else:
        order = np.argsort(key, kind="stable")
    array[...] = key[order]
    index[...] = order + 1


def _n_median(interp, a, *rest, **k):
    return np.asarray(a).dtype.type(np.median(a))


def _n_diag(interp, v, k=0):
    return np.asfortranarray(np.diag(v, int(k)))


def _n_inv(interp, a, *rest, **k):
    return np.asfortranarray(np.linalg.inv(a).astype(a.dtype))


def _n_is_close(interp, a, b, rel_tol=None, abs_tol=None, equal_nan=None):
    """stdlib_math is_close: |a - b| <= max(rel_tol * max(|a|, |b|), abs_tol), rel_tol = sqrt(epsilon) by default"""
    eps = np.finfo(np.asarray(a).real.dtype if np.asarray(a).dtype.kind in "fc" else np.float64).eps
    rt = np.sqrt(eps) if rel_tol is None or rel_tol is ABSENT else rel_tol
    at = 0.0 if abs_tol is None or abs_tol is ABSENT else abs_tol
    return np.abs(a - b) <= np.maximum(rt * np.maximum(np.abs(a), np.abs(b)), at)


def _n_c_associated(interp, p, q=None):
    if q is None or q is ABSENT:
        return p.obj is not None
    return p.obj is not None and p.obj is q.obj


def _n_c_f_pointer(interp, cptr, fptr, shape=None):
    arr = cptr.obj
    if isinstance(arr, ScalarRef):                     # c_loc of a scalar (derived-type) variable: the object itself
        return ("__out__", {1: arr.get()})
    if shape is not None and shape is not ABSENT:
        arr = arr.ravel(order="F")[:int(np.prod(shape))].reshape(tuple(int(x) for x in shape), order="F")
    return ("__out__", {1: arr})


def _n_type_error(interp, *a, **k):
    raise StopError(f"type_error{a}")


NATIVES = {
    # logging / timing / error helpers (src/Utilities/Logger.f90, Timer.f90): side effects only
    "log_message": _n_noop, "log_information": _n_noop, "log_warning": _n_noop, "log_debug": _n_noop, "log_error": _n_noop,
    "time_lightkrylov": lambda interp: False,
    "lightkrylov_timer": lambda interp, *a, **k: Inst("lightkrylov_timer"),      # Timer_Utils.f90 is not loaded: its methods are no-ops
    "stop_error": _n_stop_error, "type_error": _n_type_error, "check_info": _n_check_info,
    "check_allocation": _n_noop, "assert_shape": _n_assert_shape,
    # fortran-stdlib
    "optval": _n_optval, "eye": _n_eye, "mnorm": _n_mnorm, "norm": _n_norm,
    "scal": _n_scal, "axpy": _n_axpy, "dot": _n_dot, "dotc": _n_dotc, "nrm2": _n_nrm2, "gemv": _n_gemv,
    "random_number": _n_random_number, "hermitian": _n_hermitian, "normal": _n_normal,
    "diag": _n_diag, "inv": _n_inv, "median": _n_median, "sort_index": _n_sort_index, "expm": _n_expm,
    "eigh": _n_eigh, "svd": _n_svd,
    # LAPACK (stdlib_linalg_lapack generic names)
    "lasr": _n_lasr, "lartg": _n_lartg, "trtrs": _n_trtrs, "geev": _n_geev, "trsen": _n_trsen, "schur": _n_schur,
    "is_close": _n_is_close, "save_npy": _n_noop, "c_associated": _n_c_associated, "c_f_pointer": _n_c_f_pointer, "padr": lambda interp, s, n, *a: str(s).ljust(int(n)),
}
