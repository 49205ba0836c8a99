"""TEST INFRASTRUCTURE: an interpreter for the small Fortran subset the reference's compute kernels are written in.

Why.  The reference is Fortran and this image has no Fortran compiler, so `oracle/_ref` cannot be built and the C
restatement in `oracle/plbm_oracle_impl.h` could only be pinned on the reference's two golden files (Bardow FVM + BGK, DUGKS).
This module removes the human from the loop for the rest: it reads the reference's OWN SOURCE TEXT where it lies
(`/root/reference/src/*.f90|F90`), preprocesses it (cpp conditionals, `!$omp` lines, continuations), parses the procedures and
EXECUTES their statements one by one with Fortran's typing rules on IEEE scalars (numpy float32 / float64: every add, multiply and
divide individually rounded, no contraction -- what gfortran emits for baseline x86-64).  `oracle/make_refsrc_golden.py` runs the
reference kernels through it on seeded inputs and commits inputs and outputs as fixtures (`tests/golden/refsrc_*.npz`);
`tests/test_oracle_refsrc.py` requires the C oracle to reproduce them BIT FOR BIT, and regenerates them wherever
`/root/reference` exists.  It is neither fast nor general: a kernel call costs milliseconds per node, and only what the kernels
use is implemented -- anything else raises, nothing is guessed.

Supported: modules with `parameter` constants (scalar and `[real(wp) :: ...]` arrays), subroutines / functions (internal ones
included, `result(...)`), declarations of integer / real(wp) / logical scalars and explicit- or assumed-shape arrays with lower
bounds, `do`, block and one-line `if`, `call` with positional or keyword arguments, assignments to scalars, elements, whole arrays
and sections, the operators + - * / ** (integer exponent 2 only) and the relational / logical ones, the intrinsics mod, size,
real, sqrt, abs, min, max, hypot, merge.  Literals carry their kind: `1.0_wp`, `1.0d0`, `1.0` (default real = float32, as in
Fortran), integers; mixed-kind arithmetic promotes like Fortran (integer -> real of the other operand's kind; float32 -> float64).
"""
import ctypes
import ctypes.util
import os
import re

import numpy as np

# the transcendental intrinsics go to the C library a gfortran binary on this machine would call (libm's sin / sinf ...), not to
# numpy's own vectorised implementations, whose last bit may differ
_LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_LIBM_FUNCS = {}
for _n in ("sin", "cos", "tan", "exp", "log", "atan", "asin", "acos", "tanh", "sinh", "cosh"):
    _d, _f = getattr(_LIBM, _n), getattr(_LIBM, _n + "f")
    _d.restype, _d.argtypes = ctypes.c_double, [ctypes.c_double]
    _f.restype, _f.argtypes = ctypes.c_float, [ctypes.c_float]
    _LIBM_FUNCS[_n] = (_d, _f)

REF_SRC = "/root/reference/src"


class FortranError(RuntimeError):
    pass


# ---------------------------------------------------------------------------------------------------------------- source
def preprocess(text, macros):
    """cpp conditionals (#if NAME / #ifdef / #ifndef / #elif / #else / #endif, names only), comments, !$omp lines, continuations."""
    out, stack = [], []
    for raw in text.splitlines():
        line = raw.rstrip()
        s = line.strip()
        if s.startswith("#"):
            if re.match(r"#\s*(warning|error|pragma)\b", s):
                if all(b[0] for b in stack) and re.match(r"#\s*error", s):
                    raise FortranError(f"live #error: {s}")
                continue
            m = re.match(r"#\s*(if|ifdef|ifndef|else|endif|elif|define|include)\b\s*(.*)", s)
            if not m:
                raise FortranError(f"preprocessor line not understood: {s}")
            kw, rest = m.group(1), m.group(2).strip()
            def cond(rest):
                """#if expressions of the reference: names, integers, defined(X), !, &&, ||, parentheses"""
                toks = re.findall(r"defined\s*\(\s*\w+\s*\)|defined\s+\w+|\w+|\|\||&&|!|\(|\)|\S", rest)
                py = []
                for t_ in toks:
                    if t_.startswith("defined"):
                        py.append(str(bool(macros.get(re.findall(r"\w+", t_)[1], 0))))
                    elif t_ in ("||", "&&", "!", "(", ")"):
                        py.append({"||": "or", "&&": "and", "!": "not"}.get(t_, t_))
                    elif re.fullmatch(r"\d+", t_):
                        py.append(str(int(t_) != 0))
                    elif re.fullmatch(r"\w+", t_):
                        py.append(str(bool(macros.get(t_, 0))))
                    else:
                        raise FortranError(f"preprocessor condition not understood: {s}")
                expr = " ".join(py)
                return bool(eval(expr, {"__builtins__": {}}, {}))

            if kw == "ifdef":
                v = bool(macros.get(rest, 0))
                stack.append([v, v])
            elif kw == "if":
                v = cond(rest)
                stack.append([v, v])  # [this branch is live, some branch of the chain has been live]
            elif kw == "ifndef":
                v = not macros.get(rest, 0)
                stack.append([v, v])
            elif kw == "elif":
                v = (not stack[-1][1]) and cond(rest)
                stack[-1] = [v, stack[-1][1] or v]
            elif kw == "else":
                stack[-1] = [not stack[-1][1], True]
            elif kw == "endif":
                stack.pop()
            elif kw == "define" and all(b[0] for b in stack):
                mm = re.match(r"(\w+)\s*(\S*)", rest)
                macros[mm.group(1)] = int(mm.group(2)) if re.fullmatch(r"-?\d+", mm.group(2) or "") else 1
            elif kw == "define":
                pass
            else:
                raise FortranError(f"preprocessor directive not supported: {s}")
            continue
        if not all(b[0] for b in stack):
            continue
        # strip comments (no string literals with '!' in the kernels' statements we execute; guard anyway)
        if "!" in line:
            q = None
            for i, ch in enumerate(line):
                if q:
                    if ch == q:
                        q = None
                elif ch in "'\"":
                    q = ch
                elif ch == "!":
                    line = line[:i]
                    break
        out.append(line.rstrip())
    # continuations
    joined, cur = [], ""
    for line in out:
        s = line.strip()
        if not s:
            continue
        if s.startswith("&"):
            s = s[1:].lstrip()
        if s.endswith("&"):
            cur += s[:-1] + " "
            continue
        joined.append((cur + s).strip())
        cur = ""
    if cur:
        joined.append(cur.strip())
    stmts = []
    for line in joined:
        low = line.lower()
        stmts.extend(p.strip() for p in split_top(low, ";") if p.strip())
    return stmts


def split_top(s, sep=","):
    """split at separators that are not inside parentheses / brackets / quotes"""
    out, depth, cur, q = [], 0, "", None
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
            cur += ch
        elif ch in "([":
            depth += 1
            cur += ch
        elif ch in ")]":
            depth -= 1
            cur += ch
        elif ch == sep and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return out


# ------------------------------------------------------------------------------------------------------------ expressions
TOKEN = re.compile(r"""
    (?P<num>(\d+\.\d*|\.\d+|\d+)([ed][+-]?\d+)?(_\w+)?)
  | (?P<dotop>\.(and|or|not|eq|ne|lt|le|gt|ge|true|false)\.)
  | (?P<name>[a-z_]\w*(\s*%\s*[a-z_]\w*)*)
  | (?P<op>\*\*|==|/=|<=|>=|=>|::|[-+*/(),:<>=%\[\]])
  | (?P<ws>\s+)
""", re.X)


def tokenize(s):
    toks, i = [], 0
    while i < len(s):
        m = TOKEN.match(s, i)
        if not m:
            raise FortranError(f"cannot tokenize {s[i:]!r} in {s!r}")
        i = m.end()
        if m.lastgroup == "ws":
            continue
        # "1.and." style ambiguity does not occur in the kernels; a number followed by a dotop is tokenized greedily otherwise
        v = m.group(m.lastgroup)
        toks.append((m.lastgroup, re.sub(r"\s+", "", v) if m.lastgroup == "name" else v))
    return toks


class Parser:
    """recursive descent over the token list -> nested tuples"""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def take(self, val=None):
        k, v = self.peek()
        if val is not None and v != val:
            raise FortranError(f"expected {val!r}, got {v!r}")
        self.i += 1
        return k, v

    def expr(self):
        return self.p_or()

    def p_or(self):
        a = self.p_and()
        while self.peek()[1] == ".or.":
            self.take()
            a = ("or", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.peek()[1] == ".and.":
            self.take()
            a = ("and", a, self.p_not())
        return a

    def p_not(self):
        if self.peek()[1] == ".not.":
            self.take()
            return ("not", self.p_not())
        return self.p_rel()

    REL = {"==": "eq", ".eq.": "eq", "/=": "ne", ".ne.": "ne", "<": "lt", ".lt.": "lt", "<=": "le", ".le.": "le", ">": "gt", ".gt.": "gt",
           ">=": "ge", ".ge.": "ge"}

    def p_rel(self):
        a = self.p_add()
        if self.peek()[1] in self.REL:
            op = self.REL[self.take()[1]]
            a = (op, a, self.p_add())
        return a

    def p_add(self):
        # a leading sign applies to the first TERM (Fortran: -a*b = -(a*b))
        if self.peek()[1] in ("+", "-"):
            sign = self.take()[1]
            a = self.p_mul()
            if sign == "-":
                a = ("neg", a)
        else:
            a = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.take()[1]
            a = ("add" if op == "+" else "sub", a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.take()[1]
            a = ("mul" if op == "*" else "div", a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_primary()
        if self.peek()[1] == "**":
            self.take()
            # right associative; a signed exponent does not occur
            a = ("pow", a, self.p_pow())
        return a

    def p_primary(self):
        k, v = self.peek()
        if k == "num":
            self.take()
            return ("num", v)
        if k == "dotop" and v in (".true.", ".false."):
            self.take()
            return ("bool", v == ".true.")
        if v == "(":
            self.take()
            e = self.expr()
            self.take(")")
            return ("paren", e)
        if v == "[":
            self.take()
            kind = None
            # [real(wp) :: a, b, ...]
            save = self.i
            if self.peek()[1] in ("real", "integer"):
                tname = self.take()[1]
                if self.peek()[1] == "(":
                    self.take()
                    kind = self.take()[1]
                    self.take(")")
                if self.peek()[1] == "::":
                    self.take()
                    kind = (tname, kind)
                else:
                    self.i, kind = save, None
            items = []
            while self.peek()[1] != "]":
                items.append(self.expr())
                if self.peek()[1] == ",":
                    self.take()
            self.take("]")
            return ("array", kind, items)
        if k == "name":
            self.take()
            if self.peek()[1] == "(":
                self.take()
                args = []
                while self.peek()[1] != ")":
                    args.append(self.p_arg())
                    if self.peek()[1] == ",":
                        self.take()
                self.take(")")
                return ("ref", v, args)
            return ("var", v)
        raise FortranError(f"unexpected token {v!r}")

    def p_arg(self):
        """subscript / actual argument: expr | [lo]:[hi] | name = expr"""
        k, v = self.peek()
        if k == "name" and self.i + 1 < len(self.t) and self.t[self.i + 1][1] == "=" and (self.i + 2 >= len(self.t) or self.t[self.i + 2][1] != "="):
            self.take()
            self.take("=")
            return ("kw", v, self.expr())
        lo = None
        if self.peek()[1] != ":":
            lo = self.expr()
            if self.peek()[1] != ":":
                return lo
        self.take(":")
        hi = None
        if self.peek()[1] not in (",", ")"):
            hi = self.expr()
        return ("slice", lo, hi)


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if p.i != len(p.t):
        raise FortranError(f"trailing tokens in expression {s!r}")
    return e


# ------------------------------------------------------------------------------------------------------------- run time
class FArray:
    """numpy storage + Fortran lower bounds"""

    def __init__(self, data, lower=None):
        self.a = data
        self.lb = list(lower) if lower is not None else [1] * data.ndim

    def index(self, subs):
        if len(subs) != self.a.ndim:
            raise FortranError("rank mismatch in array reference")
        idx = []
        for d, s in enumerate(subs):
            if isinstance(s, slice):
                lo = self.lb[d] if s.start is None else s.start
                hi = self.lb[d] + self.a.shape[d] - 1 if s.stop is None else s.stop
                if lo < self.lb[d] or hi > self.lb[d] + self.a.shape[d] - 1:
                    raise FortranError("array section out of bounds")
                idx.append(slice(lo - self.lb[d], hi - self.lb[d] + 1))
            else:
                i = int(s) - self.lb[d]
                if not 0 <= i < self.a.shape[d]:
                    raise FortranError(f"subscript {s} of dimension {d + 1} out of bounds [{self.lb[d]}, {self.lb[d] + self.a.shape[d] - 1}]")
                idx.append(i)
        return tuple(idx)


ABSENT = object()


class Proc:
    def __init__(self, kind, name, args, result, parent, module):
        self.kind, self.name, self.args, self.result, self.parent, self.module = kind, name, args, result, parent, module
        self.decls, self.body, self.contains = {}, [], {}


class Decl:
    def __init__(self, typ, kind, dims, param_expr, line):
        self.typ, self.kind, self.dims, self.param_expr, self.line = typ, kind, dims, param_expr, line


class Module:
    def __init__(self, name):
        self.name, self.decls, self.procs, self.uses, self.values, self.renames = name, {}, {}, [], {}, {}


DECL_RE = re.compile(r"^(integer|real|logical|double precision)\b\s*(\(([^)]*)\))?(.*?)::(.*)$")
SKIP_DECL = re.compile(r"^(type\s*\(|class\s*\(|procedure\b|character\b|external\b|save\b|import\b|abstract\b|interface\b|end interface|"
                       r"public\b|private\b|implicit\b|use\b|generic\b|module procedure\b|intrinsic\b)")


class Interp:
    def __init__(self, wp="f64", macros=None, src=REF_SRC):
        self.wp = np.float64 if wp == "f64" else np.float32
        self.macros = dict(macros or {})
        self.src = src
        self.modules = {}
        self.nstmt = 0

    # ---- parsing --------------------------------------------------------------------------------------------------
    def _global(self):
        return self.modules.setdefault("_global", Module("_global"))

    def load(self, filename):
        path = os.path.join(self.src, filename)
        with open(path) as fh:
            return self.load_text(fh.read())

    def load_text(self, text):
        stmts = preprocess(text, self.macros)
        mod, stack, skip_until = None, [], None
        for st in stmts:
            if skip_until:
                if re.match(skip_until, st):
                    skip_until = None
                continue
            m = re.match(r"^module\s+(\w+)$", st)
            if m and not st.startswith("module procedure"):
                mod = Module(m.group(1))
                self.modules[mod.name] = mod
                continue
            if re.match(r"^end\s*module", st):
                mod = None
                continue
            if re.match(r"^(abstract\s+)?interface\b", st):
                skip_until = r"^end\s*interface"
                continue
            if re.match(r"^type\b(?!\s*\()", st):  # a derived-type definition (not a `type(name) ::` declaration)
                skip_until = r"^end\s*type"
                continue
            if st == "contains":
                continue
            m = re.match(r"^(?:(?:pure|elemental|recursive|impure)\s+)*(?:(real\s*\(\s*\w+\s*\)|integer|logical)\s+)?(subroutine|function)\s+(\w+)\s*(\(([^)]*)\))?\s*(result\s*\(\s*(\w+)\s*\))?", st)
            if m and not re.match(r"^end\b", st):
                rtype, kind, name, args, res = m.group(1), m.group(2), m.group(3), m.group(5), m.group(7)
                parent = stack[-1] if stack else None
                p = Proc(kind, name, [a.strip() for a in (args or "").split(",") if a.strip()], res or (name if kind == "function" else None), parent, mod)
                if rtype:
                    p.decls[p.result] = Decl("real" if rtype.startswith("real") else rtype, "wp" if rtype.startswith("real") else None, None, None, st)
                if parent is None and mod is None:
                    p.module = self._global()  # an external procedure (the bind(c) wrappers of sim/): parsed, not ours to run
                (parent.contains if parent else p.module.procs)[name] = p
                stack.append(p)
                continue
            if re.match(r"^end\s*(subroutine|function)", st) or (st == "end" and stack):
                stack.pop()
                continue
            target = stack[-1] if stack else mod
            if target is None:
                continue  # program units we do not run
            if re.match(r"^use\b", st):
                m = re.match(r"^use\s*(?:,\s*\w+\s*)?(?:::)?\s*(\w+)", st)
                (mod or self._global()).uses.append(m.group(1))
                for local, remote in re.findall(r"(\w+)\s*=>\s*(\w+)", st):
                    (mod or self._global()).renames[local] = remote  # use m, only: local => remote
                continue
            if SKIP_DECL.match(st):
                continue
            m = DECL_RE.match(st)
            if m and self._looks_like_decl(st):
                self._declare(target, m, st)
                continue
            if stack:
                target.body.append(st)
        return self

    @staticmethod
    def _looks_like_decl(st):
        return "::" in st and re.match(r"^(integer|real|logical|double precision)\b", st) is not None

    def _declare(self, target, m, st):
        typ, kind, attrs, names = m.group(1), (m.group(3) or "").strip(), m.group(4), m.group(5)
        kind = kind.replace("kind", "").replace("=", "").strip() or None
        is_param = "parameter" in attrs
        dim_attr = re.search(r"dimension\s*\((.*)\)", attrs)
        for item in split_top(names):
            item = item.strip()
            if not item:
                continue
            expr = None
            if "=" in item and is_param or re.search(r"[^=/<>]=[^=]", item):
                lhs, expr = item.split("=", 1)
                item, expr = lhs.strip(), expr.strip()
            mm = re.match(r"^(\w+)\s*(\((.*)\))?$", item)
            if not mm:
                raise FortranError(f"declaration not understood: {st}")
            dims = mm.group(3) if mm.group(3) is not None else (dim_attr.group(1) if dim_attr else None)
            target.decls[mm.group(1)] = Decl(typ, kind, dims, expr if is_param or expr else None, st)

    # ---- values -----------------------------------------------------------------------------------------------------
    def real_kind(self, kind):
        if kind in (None, ""):
            return np.float32  # default real
        if kind == "wp":
            return self.wp
        if kind in ("dp", "real64", "8"):
            return np.float64
        if kind in ("sp", "real32", "4"):
            return np.float32
        raise FortranError(f"real kind {kind!r} not known")

    def literal(self, text):
        m = re.fullmatch(r"(\d+\.\d*|\.\d+|\d+)([ed][+-]?\d+)?(?:_(\w+))?", text)
        mant, expo, kind = m.group(1), m.group(2), m.group(3)
        if "." not in mant and expo is None:
            # an integer literal; with a kind suffix (`90_wp` in sim/sim_lw6.F90: an INTEGER of kind 8, since wp = 8) still an integer
            return int(mant)
        if expo and expo[0] == "d":
            if kind:
                raise FortranError("d exponent with a kind suffix")
            return np.float64(mant + "e" + expo[1:])
        # decimal -> binary conversion is correctly rounded in numpy's constructors, like in gfortran
        return self.real_kind(kind)((mant + (expo or "")))

    @staticmethod
    def promote(a, b):
        """Fortran's numeric conversion for a binary operation"""
        ia, ib = isinstance(a, (int, np.integer)) and not isinstance(a, bool), isinstance(b, (int, np.integer)) and not isinstance(b, bool)
        if ia and ib:
            return int(a), int(b)
        ta = a.dtype.type if isinstance(a, (np.ndarray, np.floating)) else None
        tb = b.dtype.type if isinstance(b, (np.ndarray, np.floating)) else None
        if ia:
            return tb(int(a)), b
        if ib:
            return a, ta(int(b))
        if ta is None or tb is None:
            raise FortranError(f"operands {type(a)} / {type(b)} not numeric")
        if ta is tb:
            return a, b
        wide = np.float64
        return (a.astype(wide) if isinstance(a, np.ndarray) else wide(a)), (b.astype(wide) if isinstance(b, np.ndarray) else wide(b))

    # ---- evaluation -------------------------------------------------------------------------------------------------
    def lookup(self, frame, name):
        if "%" in name:  # component of a derived-type object (a python dict bound to the dummy argument)
            base, *fields = name.split("%")
            obj = self.lookup(frame, base)
            for fld in fields:
                if not isinstance(obj, dict) or fld not in obj:
                    raise FortranError(f"{name}: no component {fld}")
                obj = obj[fld]
            return obj
        f = frame
        while f is not None:
            if name in f["vars"]:
                return f["vars"][name]
            proc = f["proc"]
            if name in proc.decls and proc.decls[name].param_expr is not None and proc.decls[name].dims is None:
                v = self.convert(self.eval(parse_expr(proc.decls[name].param_expr), f), proc.decls[name])
                f["vars"][name] = v
                return v
            f = f["host"]
        return self.module_value(frame["proc"].module, name)

    def module_value(self, mod, name, seen=()):
        if name in mod.values:
            return mod.values[name]
        if name in mod.decls:
            d = mod.decls[name]
            if d.param_expr is None:
                raise FortranError(f"module variable {name} is not a parameter")
            fr = {"vars": {}, "proc": Proc("module", mod.name, [], None, None, mod), "host": None}
            v = self.eval(parse_expr(d.param_expr), fr)
            if d.dims is not None:
                lo, shape = self.dims_of(d.dims, fr)
                arr = np.asarray(v.a if isinstance(v, FArray) else v)
                v = FArray(arr.reshape(shape), lo)
            else:
                v = self.convert(v, d)
            mod.values[name] = v
            return v
        for u in mod.uses:
            if u in self.modules and u not in seen:
                try:
                    return self.module_value(self.modules[u], name, seen + (mod.name,))
                except KeyError:
                    pass
        raise KeyError(name)

    def convert(self, v, decl):
        if decl.typ == "integer":
            if isinstance(v, (np.floating, float)):
                return int(np.trunc(v))
            return int(v)
        if decl.typ in ("real", "double precision"):
            t = np.float64 if decl.typ == "double precision" else self.real_kind(decl.kind)
            return t(v)
        if decl.typ == "logical":
            return bool(v)
        raise FortranError("type not supported")

    def dims_of(self, dims, frame):
        lo, shape = [], []
        for d in split_top(dims):
            d = d.strip()
            if d == ":":
                return None, None  # assumed shape
            if ":" in d:
                a, b = d.split(":")
                l, u = int(self.eval(parse_expr(a), frame)), int(self.eval(parse_expr(b), frame))
            else:
                l, u = 1, int(self.eval(parse_expr(d), frame))
            lo.append(l)
            shape.append(u - l + 1)
        return lo, tuple(shape)

    def find_proc(self, frame, name):
        p = frame["proc"]
        while p is not None and p.kind != "module":
            if name in p.contains:
                return p.contains[name]
            p = p.parent
        mod = frame["proc"].module
        name = mod.renames.get(name, name)
        for m in [mod] + [self.modules[u] for u in mod.uses if u in self.modules]:
            if name in m.procs:
                return m.procs[name]
        return None

    def eval(self, e, fr):
        k = e[0]
        if k == "num":
            return self.literal(e[1])
        if k == "bool":
            return e[1]
        if k == "paren":
            return self.eval(e[1], fr)
        if k == "var":
            v = self.lookup(fr, e[1])
            if v is None:
                raise FortranError(f"variable {e[1]} used before it is defined")
            return v.a if isinstance(v, FArray) else v
        if k == "neg":
            return -self.eval(e[1], fr)
        if k in ("add", "sub", "mul", "div"):
            a, b = self.promote(self.eval(e[1], fr), self.eval(e[2], fr))
            if k == "add":
                return a + b
            if k == "sub":
                return a - b
            if k == "mul":
                return a * b
            if isinstance(a, int):
                if b == 0:
                    raise FortranError("integer division by zero")
                q = abs(a) // abs(b)
                return q if (a >= 0) == (b >= 0) else -q
            return a / b
        if k == "pow":
            a, b = self.eval(e[1], fr), self.eval(e[2], fr)
            if isinstance(b, int) and b == 2:
                return a * a  # what gfortran emits for x**2
            if isinstance(a, int) and isinstance(b, int) and b >= 0:
                return a ** b
            raise FortranError("only **2 (and integer ** integer) is supported")
        if k in ("eq", "ne", "lt", "le", "gt", "ge"):
            a, b = self.promote(self.eval(e[1], fr), self.eval(e[2], fr))
            return {"eq": a == b, "ne": a != b, "lt": a < b, "le": a <= b, "gt": a > b, "ge": a >= b}[k]
        if k == "and":
            return bool(self.eval(e[1], fr)) and bool(self.eval(e[2], fr))
        if k == "or":
            return bool(self.eval(e[1], fr)) or bool(self.eval(e[2], fr))
        if k == "not":
            return not self.eval(e[1], fr)
        if k == "array":
            vals = [self.eval(x, fr) for x in e[2]]
            if e[1]:
                t = self.real_kind(e[1][1]) if e[1][0] == "real" else np.int64
                return np.array([t(v) for v in vals], dtype=t)
            return np.array(vals)
        if k == "ref":
            return self.eval_ref(e, fr)
        raise FortranError(f"expression node {k} not supported")

    def eval_ref(self, e, fr):
        name, args = e[1], e[2]
        # variable (array) first: a local array may shadow an intrinsic
        try:
            v = self.lookup(fr, name)
        except KeyError:
            v = None
        if isinstance(v, FArray):
            subs = [self.subscript(a, fr) for a in args]
            return v.a[v.index(subs)]
        if v is not None:
            raise FortranError(f"{name} is a scalar but is referenced with subscripts")
        proc = self.find_proc(fr, name)
        if proc is not None:
            return self.call(proc, args, fr)
        return self.intrinsic(name, args, fr)

    def subscript(self, a, fr):
        if a[0] == "slice":
            lo = None if a[1] is None else int(self.eval(a[1], fr))
            hi = None if a[2] is None else int(self.eval(a[2], fr))
            return slice(lo, hi)
        v = self.eval(a, fr)
        if not isinstance(v, (int, np.integer)):
            raise FortranError("subscript is not an integer")
        return int(v)

    def intrinsic(self, name, args, fr):
        if name == "size":
            arr = self.lookup(fr, args[0][1]) if args[0][0] == "var" else None
            if not isinstance(arr, FArray):
                raise FortranError("size() of something that is not an array variable")
            return int(arr.a.size) if len(args) == 1 else int(arr.a.shape[int(self.eval(args[1], fr)) - 1])
        if name == "present":
            return self.lookup(fr, args[0][1]) is not ABSENT
        if name == "real":  # real(x [, kind]): the kind is a name, not a value
            kind = None
            if len(args) > 1:
                k = args[1][2] if args[1][0] == "kw" else args[1]
                kind = k[1] if k[0] == "var" else str(self.eval(k, fr))
            t = self.real_kind(kind)
            v = self.eval(args[0], fr)
            return v.astype(t) if isinstance(v, np.ndarray) else t(v)
        vals = [self.eval(a[2] if a[0] == "kw" else a, fr) for a in args]
        if name == "mod":
            a, b = vals
            if isinstance(a, int) and isinstance(b, int):
                r = abs(a) % abs(b)
                return r if a >= 0 else -r
            raise FortranError("mod() of reals is not supported")
        if name in _LIBM_FUNCS:
            x = vals[0]
            if isinstance(x, np.float32):
                return np.float32(_LIBM_FUNCS[name][1](float(x)))
            if isinstance(x, np.float64):
                return np.float64(_LIBM_FUNCS[name][0](float(x)))
            raise FortranError(f"{name}() of {type(x).__name__}: scalars of a real kind only")
        if name == "sqrt":
            return np.sqrt(vals[0])
        if name == "abs":
            return abs(vals[0])
        if name in ("min", "max"):
            r = vals[0]
            for v in vals[1:]:
                a, b = self.promote(r, v)
                r = (a if a <= b else b) if name == "min" else (a if a >= b else b)
            return r
        if name == "hypot":
            a, b = self.promote(vals[0], vals[1])
            return np.hypot(a, b)
        if name == "merge":
            return vals[0] if vals[2] else vals[1]
        raise FortranError(f"procedure or intrinsic {name!r} is not supported")

    # ---- calls and statements -----------------------------------------------------------------------------------------
    def call(self, proc, args, caller, actuals=None):
        """args: expression nodes evaluated in `caller` (arrays by reference, sections as numpy views), or `actuals`: python values"""
        fr = {"vars": {}, "proc": proc, "host": None}
        if proc.parent is not None and caller is not None:
            # host association: an internal procedure sees the frame of the host that is executing (or, when called from outside
            # for a test, an empty frame of that host whose parameters are evaluated on demand)
            h = caller
            while h is not None and h["proc"] is not proc.parent:
                h = h["host"] if h["host"] is not None else None
            fr["host"] = h
        if fr["host"] is None and proc.parent is not None:
            fr["host"] = {"vars": {}, "proc": proc.parent, "host": None}
        bound, byref = {}, []
        if actuals is not None:
            for n, v in zip(proc.args, actuals):
                bound[n] = v
        else:
            pos = 0
            for a in args:
                name, node = (a[1], a[2]) if a[0] == "kw" else (proc.args[pos], a)
                pos += a[0] != "kw"
                bound[name] = self.actual(node, caller)
                if node[0] == "var":
                    byref.append((name, node))  # a variable: argument association, a scalar dummy writes through
        # dummies: scalars first (array bounds may depend on them)
        for n in proc.args:
            d = proc.decls.get(n)
            if n not in bound:
                fr["vars"][n] = ABSENT  # an optional argument that was not passed
                continue
            if d is None:
                if not isinstance(bound[n], dict):
                    raise FortranError(f"dummy argument {n} of {proc.name} has no declaration we understand")
                fr["vars"][n] = bound[n]  # a derived-type object
                continue
            if d.dims is None:
                fr["vars"][n] = None if bound[n] is None else self.convert(bound[n], d)  # None: an actual that is not defined yet
        for n in proc.args:
            d = proc.decls.get(n)
            if d is not None and d.dims is not None and n in bound:
                v = bound[n]
                arr = v.a if isinstance(v, FArray) else v
                lo, shape = self.dims_of(d.dims, fr)
                if lo is None:
                    fr["vars"][n] = FArray(arr)
                else:
                    if arr.size < int(np.prod(shape)):
                        raise FortranError(f"actual argument for {n} is smaller than its declared shape")
                    if tuple(arr.shape) != shape:
                        raise FortranError(f"explicit-shape dummy {n}{shape} bound to an actual of shape {arr.shape}: pass arrays of the declared shape")
                    fr["vars"][n] = FArray(arr, lo)
        # locals
        for n, d in proc.decls.items():
            if n in fr["vars"] or n in proc.args:
                continue
            if d.param_expr is not None:
                continue  # evaluated on demand
            if d.dims is not None:
                lo, shape = self.dims_of(d.dims, fr)
                if lo is None:
                    fr["vars"][n] = None  # allocatable / pointer: nothing until move_alloc gives it something
                    continue
                t = int if d.typ == "integer" else (self.real_kind(d.kind) if d.typ == "real" else bool)
                fr["vars"][n] = FArray(np.full(shape, np.nan if d.typ == "real" else 0, dtype=t), lo)
            else:
                fr["vars"][n] = None
        if proc.kind == "function" and proc.result not in proc.decls:
            fr["vars"][proc.result] = {}  # a derived-type result: an object whose components the body defines
        self.exec_block(proc.body, 0, len(proc.body), fr)
        for n, node in byref:
            v = fr["vars"].get(n)
            if v is None or v is ABSENT or isinstance(v, (FArray, dict)):
                continue
            before = bound.get(n)
            if before is None or type(before) is not type(v) or before != v:  # the callee defined / changed the dummy
                self.assign(node, v, caller)
        if proc.kind == "function":
            v = fr["vars"][proc.result]
            return v.a.copy() if isinstance(v, FArray) else v
        return None

    def actual(self, a, fr):
        if a[0] == "paren":
            return self.eval(a, fr)
        if a[0] == "var":
            return self.lookup(fr, a[1])  # may be None: a variable that the callee defines (intent(out))
        if a[0] == "ref":
            try:
                v = self.lookup(fr, a[1])
            except KeyError:
                v = None
            if isinstance(v, FArray) and any(x[0] == "slice" for x in a[2]):
                subs = [self.subscript(x, fr) for x in a[2]]
                return FArray(v.a[v.index(subs)])
        return self.eval(a, fr)

    def assign(self, lhs, value, fr):
        if lhs[0] == "var" and "%" in lhs[1]:
            base, *fields = lhs[1].split("%")
            obj = self.lookup(fr, base)
            for fld in fields[:-1]:
                obj = obj[fld]
            cur = obj.get(fields[-1])
            if isinstance(cur, FArray):
                cur.a[...] = value
            elif isinstance(cur, (np.floating, float)) or (cur is None and isinstance(value, np.floating)):
                obj[fields[-1]] = self.wp(value)  # the real components of lattice_grid are real(wp)
            else:
                obj[fields[-1]] = int(value) if isinstance(value, (int, np.integer)) and not isinstance(value, bool) else value
            return
        if lhs[0] == "var":
            name = lhs[1]
            f = fr
            while f is not None and name not in f["vars"]:
                f = f["host"]
            if f is None:
                raise FortranError(f"assignment to undeclared variable {name}")
            cur = f["vars"][name]
            if isinstance(cur, FArray):
                cur.a[...] = value  # whole-array assignment: numpy converts to the array's type element by element
                return
            p = f["proc"]
            d = p.decls.get(name)
            if d is None:
                raise FortranError(f"no declaration for {name} in {p.name}")
            if isinstance(value, np.ndarray):
                raise FortranError(f"array assigned to scalar {name}")
            f["vars"][name] = self.convert(value, d)
            return
        if lhs[0] == "ref":
            arr = self.lookup(fr, lhs[1])
            if not isinstance(arr, FArray):
                raise FortranError(f"{lhs[1]} is not an array")
            subs = [self.subscript(a, fr) for a in lhs[2]]
            arr.a[arr.index(subs)] = value
            return
        raise FortranError("left-hand side not supported")

    def bind_name(self, name, obj, fr):
        """make a variable or a component NAME the array object `obj` (move_alloc), as opposed to copying values into it"""
        if "%" in name:
            base, *fields = name.split("%")
            o = self.lookup(fr, base)
            for fld in fields[:-1]:
                o = o[fld]
            o[fields[-1]] = obj
            return
        f = fr
        while f is not None and name not in f["vars"]:
            f = f["host"]
        if f is None:
            raise FortranError(f"move_alloc: {name} is not declared")
        f["vars"][name] = obj

    def exec_block(self, body, i, end, fr):
        """executes body[i:end]; returns 'cycle' / 'exit' / 'return' / None"""
        while i < end:
            st = body[i]
            self.nstmt += 1
            m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", st)
            if m:
                j = self.match_end(body, i, r"^do\b", r"^end\s*do$")
                parts = split_top(m.group(2))
                lo, hi = int(self.eval(parse_expr(parts[0]), fr)), int(self.eval(parse_expr(parts[1]), fr))
                step = int(self.eval(parse_expr(parts[2]), fr)) if len(parts) > 2 else 1
                v = lo
                while (v <= hi) if step > 0 else (v >= hi):
                    self.assign(("var", m.group(1)), v, fr)
                    r = self.exec_block(body, i + 1, j, fr)
                    if r == "exit":
                        break
                    if r == "return":
                        return r
                    v += step
                i = j + 1
                continue
            m = re.match(r"^if\s*\((.*)\)\s*then$", st)
            if m:
                j = self.match_end(body, i, r"^if\s*\(.*\)\s*then$", r"^end\s*if$")
                # branches at depth 0
                marks, depth = [(i, m.group(1))], 0
                for k in range(i + 1, j):
                    if re.match(r"^if\s*\(.*\)\s*then$", body[k]):
                        depth += 1
                    elif re.match(r"^end\s*if$", body[k]):
                        depth -= 1
                    elif depth == 0:
                        mm = re.match(r"^else\s*if\s*\((.*)\)\s*then$", body[k])
                        if mm:
                            marks.append((k, mm.group(1)))
                        elif body[k] == "else":
                            marks.append((k, None))
                marks.append((j, None))
                for (a, cond), (b, _) in zip(marks[:-1], marks[1:]):
                    if cond is None or bool(self.eval(parse_expr(cond), fr)):
                        r = self.exec_block(body, a + 1, b, fr)
                        if r:
                            return r
                        break
                i = j + 1
                continue
            m = re.match(r"^if\s*\(", st)
            if m:
                # one-line if: find the matching parenthesis
                depth, k = 0, st.index("(")
                for k in range(st.index("("), len(st)):
                    depth += st[k] == "("
                    depth -= st[k] == ")"
                    if depth == 0:
                        break
                if bool(self.eval(parse_expr(st[st.index("(") + 1:k]), fr)):
                    r = self.exec_block([st[k + 1:].strip()], 0, 1, fr)
                    if r:
                        return r
                i += 1
                continue
            if st in ("cycle", "exit", "return"):
                return st
            if st == "continue" or re.match(r"^(\w+\s*:\s*)?block$|^end\s*block(\s+\w+)?$|^(print|write)\b", st):
                i += 1
                continue
            m = re.match(r"^associate\s*\((.*)\)$", st)
            if m:
                j = self.match_end(body, i, r"^associate\s*\(", r"^end\s*associate$")
                saved = {}
                for pair in split_top(m.group(1)):
                    name, expr = (x.strip() for x in pair.split("=>", 1))
                    saved[name] = fr["vars"].get(name, ABSENT)
                    fr["vars"][name] = self.actual(parse_expr(expr), fr)
                r = self.exec_block(body, i + 1, j, fr)
                for name, old in saved.items():
                    if old is ABSENT:
                        del fr["vars"][name]
                    else:
                        fr["vars"][name] = old
                if r:
                    return r
                i = j + 1
                continue
            m = re.match(r"^call\s+(\w+(?:\s*%\s*\w+)+)\s*\(\s*\)$", st)
            if m:  # type-bound / procedure-pointer component: a python callable stored in the object
                fn = self.lookup(fr, re.sub(r"\s+", "", m.group(1)))
                if not callable(fn):
                    raise FortranError(f"{m.group(1)} is not callable")
                fn()
                i += 1
                continue
            m = re.match(r"^call\s+move_alloc\s*\((.*)\)$", st)
            if m:
                e = parse_expr(f"move_alloc({m.group(1)})")
                kw = {a[1]: a[2] for a in e[2] if a[0] == "kw"}
                pos = [a for a in e[2] if a[0] != "kw"]
                src, dst = kw.get("from", pos[0] if pos else None), kw.get("to", pos[1] if len(pos) > 1 else None)
                obj = self.lookup(fr, src[1])
                if not isinstance(obj, FArray):
                    raise FortranError("move_alloc: from is not an allocated array")
                self.bind_name(dst[1], obj, fr)
                self.bind_name(src[1], None, fr)
                i += 1
                continue
            m = re.match(r"^call\s+(\w+)\s*(\((.*)\))?$", st)
            if m:
                proc = self.find_proc(fr, m.group(1))
                if proc is None:
                    raise FortranError(f"call of unknown procedure {m.group(1)}")
                e = parse_expr(f"{m.group(1)}({m.group(3) or ''})")
                self.call(proc, e[2], fr)
                i += 1
                continue
            # assignment
            parts = self.split_assignment(st)
            if parts is None:
                raise FortranError(f"statement not supported: {st}")
            lhs, rhs = parts
            self.assign(parse_expr(lhs), self.eval(parse_expr(rhs), fr), fr)
            i += 1
        return None

    @staticmethod
    def split_assignment(st):
        depth = 0
        for k, ch in enumerate(st):
            if ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            elif ch == "=" and depth == 0:
                if st[k + 1:k + 2] == "=" or st[k - 1:k] in ("=", "/", "<", ">"):
                    continue
                return st[:k].strip(), st[k + 1:].strip()
        return None

    @staticmethod
    def match_end(body, i, open_re, close_re):
        depth = 0
        for k in range(i, len(body)):
            if re.match(open_re, body[k]):
                depth += 1
            elif re.match(close_re, body[k]):
                depth -= 1
                if depth == 0:
                    return k
        raise FortranError(f"no end for {body[i]!r}")

    # ---- entry point ------------------------------------------------------------------------------------------------
    def run(self, module, path, *actuals):
        """run("collision_trt", "collide_trt/trt_naive", nx, ny, f1, ld, le, ld) -- arrays are numpy arrays of the declared shape,
        modified in place (Fortran's argument association); returns the function result, if any."""
        mod = self.modules[module]
        names = path.split("/")
        proc = mod.procs[names[0]]
        for n in names[1:]:
            proc = proc.contains[n]
        vals = []
        for v in actuals:
            if isinstance(v, np.ndarray):
                vals.append(FArray(v))
            elif isinstance(v, (float, np.floating)):
                vals.append(self.wp(v))
            else:
                vals.append(v)
        return self.call(proc, None, None, actuals=vals)

    def constant(self, module, name):
        v = self.module_value(self.modules[module], name)
        return v.a if isinstance(v, FArray) else v
