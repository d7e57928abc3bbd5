"""Fortran -> Python translator for the subset of Fortran the reference's hot path is written in.

TEST INFRASTRUCTURE ONLY (lives under oracle/; nothing in pfemfort_b200/ may import it).

Purpose: the reference (chennachaos/PFEMFort) is Fortran + PETSc + MPI + METIS and no Fortran compiler exists in this
image, so the reference cannot be built.  This module executes the reference's OWN source files, read where they lie
under /root/reference/src, statement by statement: fixed-form source -> logical statements -> AST -> Python source ->
exec.  Arithmetic uses numpy scalar types with Fortran's typing rules (INTEGER = Python int, default REAL = float32,
DOUBLE PRECISION = float64; `1.0/3.0` is therefore the single-precision quotient exactly as gfortran folds it,
promoted on assignment), integer division truncates, integer powers are repeated multiplication, MATMUL sums in
ascending inner index from zero (gfortran's inline expansion, frontend-passes.c), uninitialised locals are NaN /
sentinel so any dependence on them is visible, initialised locals are SAVEd, scalar arguments are passed by reference
(copy-in / copy-out through `Ref` cells), arrays are numpy arrays in Fortran order with 1-based subscripts.

What is external to the reference (PETSc, MPI, METIS, the VTK writer) is provided by oracle/refrun/mocks.py.
Generated Python is written only under oracle/_ref/ (git-ignored: it is derived from the reference's sources).

Supported: PROGRAM / MODULE / SUBROUTINE units, TYPE definitions with type-bound procedures, declarations (incl. the
PETSc macro types), assignment, IF / ELSE IF / ELSE, DO (counted, infinite, named), SELECT CASE, CALL, STOP, RETURN,
EXIT, CYCLE, ALLOCATE / DEALLOCATE, OPEN / CLOSE / READ / WRITE / INQUIRE (list-directed), the intrinsics the path
uses.  Anything else raises `Unsupported` with the source line: nothing is skipped silently.
"""
from __future__ import annotations

import keyword
import re
from dataclasses import dataclass, field


class Unsupported(Exception):
    pass


# --------------------------------------------------------------------------------------------------------------------
# source -> logical statements
# --------------------------------------------------------------------------------------------------------------------

def _strip_comment(s: str) -> str:
    q = None
    for i, ch in enumerate(s):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == '!':
            return s[:i]
    return s


def _split_semicolons(s: str):
    out, cur, q = [], [], None
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == ';':
            out.append(''.join(cur))
            cur = []
        else:
            cur.append(ch)
    out.append(''.join(cur))
    return [x.strip() for x in out if x.strip()]


def logical_statements(text: str):
    """Fixed-form source (unlimited line length, -cpp): list of (first line number, statement text)."""
    joined = []
    for no, raw in enumerate(text.splitlines(), 1):
        if raw.startswith('#'):
            continue  # preprocessor lines: PETSc includes / module switches, provided by the mocks
        line = raw.rstrip()
        if not line.strip():
            continue
        if line[0] in 'cC*!' or line.lstrip().startswith('!'):
            continue
        if len(line) > 5 and line[:5].strip() == '' and line[5] not in ' 0':
            if not joined:
                raise Unsupported(f"line {no}: continuation without a statement")
            joined[-1][1] += ' ' + _strip_comment(line[6:])
            continue
        joined.append([no, _strip_comment(line[6:] if len(line) > 6 else '')])
    out = []
    for no, s in joined:
        for part in _split_semicolons(s):
            if re.match(r'^CHKERRQ\s*\(', part):
                continue  # PETSc's error-check macro (returns when ierr /= 0); the mocks always return 0
            out.append((no, part))
    return out


def logical_statements_free(text: str):
    """Free-form source (.f90): `&` continuations, `!` comments anywhere, `;` separators."""
    joined, cont = [], False
    for no, raw in enumerate(text.splitlines(), 1):
        if raw.startswith('#'):
            continue
        st = _strip_comment(raw).strip()
        if not st:
            continue
        if cont:
            if st.startswith('&'):
                st = st[1:].lstrip()
            nxt = st.endswith('&')
            joined[-1][1] += ' ' + (st[:-1].rstrip() if nxt else st)
            cont = nxt
            continue
        cont = st.endswith('&')
        joined.append([no, st[:-1].rstrip() if cont else st])
    out = []
    for no, s in joined:
        for part in _split_semicolons(s):
            out.append((no, part))
    return out


# --------------------------------------------------------------------------------------------------------------------
# tokens and expressions
# --------------------------------------------------------------------------------------------------------------------

_TOK = re.compile(r"""
 (?P<ws>\s+)
|(?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
|(?P<dotop>\.(?:and|or|not|true|false|eq|ne|lt|le|gt|ge|eqv|neqv)\.)
|(?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[edED][+-]?\d+)?)
|(?P<name>[A-Za-z][A-Za-z0-9_]*)
|(?P<op>\*\*|//|==|/=|<=|>=|=>|[-+*/()=,:%<>])
""", re.X | re.I)


def tokenize(s: str):
    toks, pos = [], 0
    while pos < len(s):
        m = _TOK.match(s, pos)
        if not m:
            raise Unsupported(f"cannot tokenize {s[pos:pos + 20]!r} in {s!r}")
        pos = m.end()
        k = m.lastgroup
        if k == 'ws':
            continue
        v = m.group(k)
        if k == 'dotop':
            v = v.lower()
            rel = {'.eq.': '==', '.ne.': '/=', '.lt.': '<', '.le.': '<=', '.gt.': '>', '.ge.': '>='}
            if v in rel:
                toks.append(('op', rel[v]))
            elif v in ('.true.', '.false.'):
                toks.append(('log', v == '.true.'))
            else:
                toks.append(('op', v))
        elif k == 'name':
            toks.append(('name', v.lower()))
        elif k == 'str':
            q = v[0]
            toks.append(('str', v[1:-1].replace(q + q, q)))
        else:
            toks.append((k, v))
    return toks


class _P:
    """Precedence-climbing expression parser over a token list."""

    def __init__(self, toks):
        self.t = toks
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ('end', None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, kind, val=None):
        tok = self.peek()
        if tok[0] == kind and (val is None or tok[1] == val):
            self.i += 1
            return True
        return False

    def expect(self, kind, val=None):
        if not self.accept(kind, val):
            raise Unsupported(f"expected {val or kind}, got {self.peek()} in {self.t}")

    def at_end(self):
        return self.i >= len(self.t)

    # expr := equiv
    def expr(self):
        return self.equiv()

    def equiv(self):
        l = self.or_()
        while self.peek() in (('op', '.eqv.'), ('op', '.neqv.')):
            op = self.next()[1]
            l = ('bin', op, l, self.or_())
        return l

    def or_(self):
        l = self.and_()
        while self.accept('op', '.or.'):
            l = ('bin', '.or.', l, self.and_())
        return l

    def and_(self):
        l = self.not_()
        while self.accept('op', '.and.'):
            l = ('bin', '.and.', l, self.not_())
        return l

    def not_(self):
        if self.accept('op', '.not.'):
            return ('un', '.not.', self.not_())
        return self.rel()

    def rel(self):
        l = self.concat()
        if self.peek()[0] == 'op' and self.peek()[1] in ('==', '/=', '<', '<=', '>', '>='):
            op = self.next()[1]
            return ('bin', op, l, self.concat())
        return l

    def concat(self):
        l = self.add()
        while self.accept('op', '//'):
            l = ('bin', '//', l, self.add())
        return l

    def add(self):
        if self.peek() in (('op', '-'), ('op', '+')):
            op = self.next()[1]
            l = ('un', op, self.mul())
        else:
            l = self.mul()
        while self.peek() in (('op', '-'), ('op', '+')):
            op = self.next()[1]
            l = ('bin', op, l, self.mul())
        return l

    def mul(self):
        l = self.pow_()
        while self.peek() in (('op', '*'), ('op', '/')):
            op = self.next()[1]
            l = ('bin', op, l, self.pow_())
        return l

    def pow_(self):
        b = self.primary()
        if self.accept('op', '**'):
            # right associative; a unary minus is allowed in the exponent
            if self.peek() in (('op', '-'), ('op', '+')):
                op = self.next()[1]
                e = ('un', op, self.pow_())
            else:
                e = self.pow_()
            return ('bin', '**', b, e)
        return b

    def arglist(self):
        """after '(' : returns list of args up to the matching ')'."""
        args = []
        if self.accept('op', ')'):
            return args
        while True:
            args.append(self.arg())
            if self.accept('op', ','):
                continue
            self.expect('op', ')')
            return args

    def arg(self):
        # keyword argument  name = expr
        if self.peek()[0] == 'name' and self.peek(1) == ('op', '='):
            n = self.next()[1]
            self.next()
            return ('kw', n, self.expr())
        # section  [lo] : [hi]
        if self.peek() == ('op', ':'):
            self.next()
            hi = None if self.peek() in (('op', ','), ('op', ')')) else self.expr()
            return ('slice', None, hi)
        if self.peek() == ('op', '*'):  # READ(1,*) style format / assumed size
            self.next()
            return ('star',)
        e = self.expr()
        if self.accept('op', ':'):
            hi = None if self.peek() in (('op', ','), ('op', ')')) else self.expr()
            return ('slice', e, hi)
        return e

    def primary(self):
        k, v = self.next()
        if k == 'num':
            return ('num', v)
        if k == 'str':
            return ('str', v)
        if k == 'log':
            return ('log', v)
        if k == 'op' and v == '(':
            e = self.expr()
            self.expect('op', ')')
            return ('paren', e)
        if k == 'name':
            node = ('name', v)
            if self.accept('op', '('):
                node = ('call', v, self.arglist())
            while self.accept('op', '%'):
                kk, f = self.next()
                if kk != 'name':
                    raise Unsupported("component name expected")
                args = None
                if self.accept('op', '('):
                    args = self.arglist()
                node = ('comp', node, f, args)
            return node
        raise Unsupported(f"unexpected token {(k, v)} in {self.t}")


def parse_expr(s: str):
    p = _P(tokenize(s))
    e = p.expr()
    if not p.at_end():
        raise Unsupported(f"trailing tokens in expression {s!r}")
    return e


def split_top(s: str, sep=','):
    """split at top-level separators (outside parentheses and strings)."""
    out, cur, depth, q = [], [], 0, None
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == '(':
            depth += 1
        elif ch == ')':
            depth -= 1
        if ch == sep and depth == 0:
            out.append(''.join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    out.append(''.join(cur).strip())
    return out


def _match_paren(s: str, start: int) -> int:
    """index of the ')' matching the '(' at s[start]."""
    depth, q = 0, None
    for i in range(start, len(s)):
        ch = s[i]
        if q:
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == '(':
            depth += 1
        elif ch == ')':
            depth -= 1
            if depth == 0:
                return i
    raise Unsupported(f"unbalanced parentheses in {s!r}")


# --------------------------------------------------------------------------------------------------------------------
# symbols
# --------------------------------------------------------------------------------------------------------------------

_PY_RESERVED = set(keyword.kwlist) | {
    'int', 'float', 'range', 'len', 'min', 'max', 'abs', 'np', 'print', 'exit', 'str', 'bool', 'type', 'size', 'sum',
    'count', 'list', 'dict', 'set', 'object', 'id', 'input', 'open', 'file', 'format', 'iter', 'next', 'all', 'any',
    'self', 'Ref', 'map', 'filter', 'vars', 'dir', 'hash', 'pow', 'round', 'slice', 'tuple', 'zip'}


def mangle(name: str) -> str:
    n = name.lower()
    return n + '_' if n in _PY_RESERVED else n


# macro types of the PETSc Fortran headers (petsc/finclude/*.h define them as integer / real kinds)
MACRO_TYPES = {
    'petscint': 'i', 'petscerrorcode': 'i', 'petscoffset': 'i', 'petscbool': 'l', 'petscscalar': 'd',
    'petscreal': 'd', 'vec': 'h', 'mat': 'h', 'ksp': 'h', 'pc': 'h', 'vecscatter': 'h', 'matinfo': 'h',
    'kspconvergedreason': 'i', 'is': 'h',
}


@dataclass
class Sym:
    name: str
    typ: str                 # i r d l c h(andle) t:<derived>
    dims: list | None = None  # list of dim source strings (':' deferred / assumed)
    dummy: bool = False
    param: bool = False
    init: str | None = None
    alloc: bool = False
    pointer: bool = False
    value: bool = False
    optional: bool = False


@dataclass
class Interface:
    """one BIND(C) function of an INTERFACE block: what a call has to marshal."""
    name: str
    args: list
    syms: dict
    result: str


@dataclass
class Unit:
    kind: str                # program | subroutine
    name: str
    args: list = field(default_factory=list)
    syms: dict = field(default_factory=dict)
    body: list = field(default_factory=list)   # (lineno, text)
    module: str | None = None
    first_line: int = 0


@dataclass
class TypeDef:
    name: str
    fields: dict = field(default_factory=dict)   # name -> typ
    procs: list = field(default_factory=list)
    inits: dict = field(default_factory=dict)    # name -> initialiser source or None


_DECL_RE = re.compile(
    r'^(integer|real|double\s+precision|logical|character|type\s*\(|class\s*\()', re.I)


def _is_decl(stmt: str) -> bool:
    s = stmt.lstrip()
    if _DECL_RE.match(s):
        # "real = 3" style assignment to a variable named like a type is not in the subset
        return True
    first = re.match(r'^([A-Za-z_]\w*)\s+[A-Za-z]', s)
    if first and first.group(1).lower() in MACRO_TYPES and '=' not in s.split('!')[0].split('(')[0]:
        return True
    return False


def parse_decl(stmt: str, syms: dict, dummies: set):
    s = stmt.strip()
    low = s.lower()
    typ = None
    rest = None
    m = re.match(r'^double\s+precision', low)
    if m:
        typ, rest = 'd', s[m.end():]
    elif low.startswith('integer'):
        typ, rest = 'i', s[7:]
        if rest.lstrip().startswith('('):               # kind selector: INTEGER(C_INT), INTEGER(C_LONG_LONG)
            r2 = rest.lstrip()
            rest = r2[_match_paren(r2, 0) + 1:]
    elif low.startswith('real'):
        typ, rest = 'r', s[4:]
        mm = re.match(r'^\s*(\(\s*(kind\s*=\s*)?8\s*\)|\*\s*8)', rest, re.I)
        if mm:
            typ, rest = 'd', rest[mm.end():]
        elif rest.lstrip().startswith('('):            # REAL(C_DOUBLE) / REAL(C_FLOAT)
            r2 = rest.lstrip()
            e = _match_paren(r2, 0)
            typ = 'd' if 'double' in r2[:e].lower() else 'r'
            rest = r2[e + 1:]
    elif low.startswith('logical'):
        typ, rest = 'l', s[7:]
    elif low.startswith('character'):
        typ, rest = 'c', s[9:]
        r2 = rest.lstrip()
        if r2.startswith('('):
            e = _match_paren(r2, 0)
            rest = r2[e + 1:]
        elif r2.startswith('*'):
            mm = re.match(r'^\*\s*(\d+|\(\s*\*\s*\))', r2)
            rest = r2[mm.end():]
    elif re.match(r'^(type|class)\s*\(', low):
        st = s.index('(')
        e = _match_paren(s, st)
        typ, rest = 't:' + s[st + 1:e].strip().lower(), s[e + 1:]
        if typ == 't:c_ptr':
            typ = 'h'                                   # TYPE(C_PTR): an opaque handle
    else:
        first = re.match(r'^([A-Za-z_]\w*)\s+', s)
        typ = MACRO_TYPES[first.group(1).lower()]
        rest = s[first.end():]
        if '::' not in rest:
            rest = ':: ' + rest
    attrs_s, _, ents = rest.partition('::')
    if not _:
        # old style "INTEGER a, b" without ::
        attrs_s, ents = '', rest
    dims_attr, param, alloc, pointer, value, optional = None, False, False, False, False, False
    for a in split_top(attrs_s):
        al = a.strip().lower()
        if not al:
            continue
        if al.startswith('dimension'):
            st = a.index('(')
            dims_attr = split_top(a[st + 1:_match_paren(a, st)])
        elif al == 'parameter':
            param = True
        elif al == 'allocatable':
            alloc = True
        elif al == 'pointer':
            pointer = True
        elif al == 'value':
            value = True
        elif al == 'optional':
            optional = True
        elif al.startswith('intent') or al in ('save', 'target', 'public', 'private'):
            pass
        else:
            raise Unsupported(f"declaration attribute {a!r} in {stmt!r}")
    for ent in split_top(ents):
        if not ent:
            continue
        init = None
        if '=>' in ent:
            ent, _, _p = ent.partition('=>')
            pointer = True
        elif '=' in ent:
            # top-level '=' only
            depth = 0
            for i, ch in enumerate(ent):
                if ch == '(':
                    depth += 1
                elif ch == ')':
                    depth -= 1
                elif ch == '=' and depth == 0:
                    init = ent[i + 1:].strip()
                    ent = ent[:i]
                    break
        ent = ent.strip()
        dims = dims_attr
        m = re.match(r'^([A-Za-z_]\w*)\s*\(', ent)
        if m:
            st = ent.index('(')
            dims = split_top(ent[st + 1:_match_paren(ent, st)])
            name = m.group(1)
        else:
            name = ent
        if not re.match(r'^[A-Za-z_]\w*$', name):
            raise Unsupported(f"entity {ent!r} in {stmt!r}")
        name = name.lower()
        syms[name] = Sym(name, typ, dims, name in dummies, param, init, alloc, pointer, value, optional)


# --------------------------------------------------------------------------------------------------------------------
# units
# --------------------------------------------------------------------------------------------------------------------

def parse_file(text: str, free_form: bool = False):
    """-> (units, typedefs, module_params, interfaces) of one source file."""
    stmts = logical_statements_free(text) if free_form else logical_statements(text)
    units, typedefs, mod_syms, interfaces = [], {}, {}, {}
    i, module = 0, None
    cur = None          # current Unit
    in_type = None
    in_decl = False
    while i < len(stmts):
        no, s = stmts[i]
        i += 1
        low = s.lower().strip()
        if cur is None and in_type is None:
            m = re.match(r'^module\s+(\w+)$', low)
            if m:
                module = m.group(1)
                continue
            if re.match(r'^end\s*module', low):
                module = None
                continue
            if re.match(r'^use\b', low) or low.startswith('implicit') or low in ('contains', 'private', 'public'):
                continue
            if low == 'interface':
                # BIND(C) function interfaces: what each call has to marshal
                fn = None
                while True:
                    no, s = stmts[i]
                    i += 1
                    low = s.lower().strip()
                    if re.match(r'^end\s*interface', low):
                        break
                    if fn is None:
                        m = re.match(r'^(.*?)\bfunction\s+(\w+)\s*\((.*?)\)\s*(bind\s*\(.*\))?\s*$', s.strip(), re.I | re.S)
                        if not m:
                            raise Unsupported(f"line {no}: only FUNCTION interfaces are supported: {s!r}")
                        rt = m.group(1).strip().lower()
                        res = 'i' if rt.startswith('integer') else ('d' if 'double' in rt else ('r' if rt.startswith('real') else None))
                        fn = Interface(m.group(2).lower(), [a.strip().lower() for a in split_top(m.group(3)) if a.strip()], {}, res)
                    elif re.match(r'^end\s*function', low):
                        for a in fn.args:
                            if a not in fn.syms:
                                raise Unsupported(f"line {no}: interface {fn.name}: dummy {a} is not declared")
                        interfaces[fn.name] = fn
                        fn = None
                    elif low.startswith('import') or re.match(r'^use\b', low) or low.startswith('implicit'):
                        pass
                    else:
                        parse_decl(s, fn.syms, set(fn.args))
                continue
            m = re.match(r'^type\s*(?:,\s*\w+\s*)*(?:::)?\s*(\w+)$', low)
            if m and module:
                in_type = TypeDef(m.group(1))
                continue
            m = re.match(r'^program\s+(\w+)', low)
            if m:
                cur = Unit('program', m.group(1), module=None, first_line=no)
                in_decl = True
                continue
            m = re.match(r'^(?:recursive\s+)?subroutine\s+(\w+)\s*(\((.*)\))?\s*$', s.strip(), re.I | re.S)
            if m:
                args = [a.strip().lower() for a in split_top(m.group(3) or '') if a.strip()]
                cur = Unit('subroutine', m.group(1).lower(), args, module=module, first_line=no)
                in_decl = True
                continue
            if re.match(r'^(\w+\s+)*function\s+\w+', low):
                raise Unsupported(f"line {no}: FUNCTION units are not in the supported subset")
            if module and _is_decl(s):
                parse_decl(s, mod_syms, set())
                continue
            raise Unsupported(f"line {no}: statement outside a program unit: {s!r}")
        if in_type is not None:
            if re.match(r'^end\s*type', low):
                typedefs[in_type.name] = in_type
                in_type = None
            elif low == 'contains':
                pass
            elif low.startswith('procedure'):
                for p in split_top(s.partition('::')[2]):
                    in_type.procs.append(p.strip().lower())
            else:
                tmp = {}
                parse_decl(s, tmp, set())
                for k, v in tmp.items():
                    in_type.fields[k] = v.typ
                    in_type.inits[k] = v.init
            continue
        # inside a unit
        if re.match(r'^end\s*(program|subroutine)?(\s+\w+)?$', low) and not re.match(r'^end\s*(if|do|select|type)', low):
            units.append(cur)
            cur = None
            continue
        if in_decl:
            if re.match(r'^use\b', low) or low.startswith('implicit'):
                continue
            if _is_decl(s):
                parse_decl(s, cur.syms, set(cur.args))
                continue
            in_decl = False
        cur.body.append((no, s))
    if cur is not None:
        raise Unsupported("unterminated program unit " + cur.name)
    return units, typedefs, mod_syms, interfaces


# --------------------------------------------------------------------------------------------------------------------
# code generation
# --------------------------------------------------------------------------------------------------------------------

INTRINSIC_FUNCS = {'min', 'max', 'abs', 'sqrt', 'acos', 'asin', 'atan', 'cos', 'sin', 'tan', 'exp', 'log', 'dble',
                   'real', 'int', 'nint', 'mod', 'size', 'matmul', 'transpose', 'trim', 'adjustl', 'len_trim',
                   'iargc', 'command_argument_count', 'allocated', 'null', 'float', 'sum', 'dot_product', 'sign',
                   'floor', 'ceiling', 'huge', 'tiny', 'epsilon', 'maxval', 'minval', 'present'}

_PROMO = {'i': 0, 'r': 1, 'd': 2}


class Gen:
    def __init__(self, unit: Unit, typedefs: dict, mod_syms: dict, all_units: dict, rewrites=None, consts=None,
                 interfaces=None):
        self.u = unit
        self.consts = consts if consts is not None else {}
        self.interfaces = interfaces or {}
        self.pre, self.post = [], []      # statements hoisted around the current one (by-reference C arguments)
        self.zero_fill = False
        self.typedefs = typedefs
        self.mod_syms = mod_syms
        self.all_units = all_units
        self.rewrites = rewrites or {}
        self.lines = []
        self.ind = 1
        self.tmp = 0
        self.loop_names = []

    # ---- helpers
    def emit(self, s):
        if self.pre:
            pre, self.pre = self.pre, []
            for p in pre:
                self.lines.append('    ' * self.ind + p)
        self.lines.append('    ' * self.ind + s)

    def c_call(self, n, args):
        """call of a BIND(C) interface function inside an expression: VALUE dummies by value, arrays as arrays, other
        scalars by reference (component -> attribute reference; local -> a Ref hoisted before the statement and written
        back after it)."""
        itf = self.interfaces[n]
        if len(args) != len(itf.args):
            raise Unsupported(f"{n}: {len(args)} actual arguments for {len(itf.args)} dummies")
        out = []
        for a, dn in zip(args, itf.args):
            d = itf.syms[dn]
            if d.value or d.dims is not None:
                out.append(self.ex(a))
            elif a[0] == 'comp' and a[3] is None:
                out.append(f'_AttrRef({self.ex(a[1])}, {mangle(a[2])!r})')
            elif a[0] == 'name' and self.sym(a[1]) is not None and self.sym(a[1]).dims is None and not self.sym(a[1]).param:
                t = self.newtmp('_c')
                self.pre.append(f'{t} = Ref({mangle(a[1])})')
                self.post.append((a, t))
                out.append(t)
            else:
                out.append(f'Ref({self.ex(a)})')
        return f'{mangle(n)}({", ".join(out)})'

    def sym(self, n):
        return self.u.syms.get(n) or self.mod_syms.get(n)

    def newtmp(self, p='_t'):
        self.tmp += 1
        return f'{p}{self.tmp}'

    # ---- static type of an expression: i r d l c or None (unknown)
    def typeof(self, e):
        k = e[0]
        if k == 'num':
            v = e[1].lower()
            if 'd' in v:
                return 'd'
            return 'r' if ('.' in v or 'e' in v) else 'i'
        if k == 'str':
            return 'c'
        if k == 'log':
            return 'l'
        if k == 'paren':
            return self.typeof(e[1])
        if k == 'name':
            s = self.sym(e[1])
            return s.typ if s and s.typ in 'irdlc' else None
        if k == 'call':
            s = self.sym(e[1])
            if s and s.dims is not None:
                return s.typ if s.typ in 'irdlc' else None
            n = e[1]
            if n in self.interfaces:
                return self.interfaces[n].result
            if n in ('min', 'max', 'mod', 'abs', 'sign', 'sum', 'maxval', 'minval'):
                ts = [self.typeof(a) for a in e[2]]
                if any(t not in _PROMO for t in ts):
                    return None
                return max(ts, key=lambda t: _PROMO[t])
            if n in ('sqrt', 'acos', 'asin', 'atan', 'cos', 'sin', 'tan', 'exp', 'log'):
                return self.typeof(e[2][0])
            if n in ('size', 'iargc', 'command_argument_count', 'int', 'nint', 'len_trim', 'floor', 'ceiling'):
                return 'i'
            if n == 'dble':
                return 'd'
            if n in ('real', 'float'):
                return 'r'
            if n in ('matmul', 'transpose', 'dot_product'):
                return self.typeof(e[2][0])
            if n in ('trim', 'adjustl'):
                return 'c'
            if n == 'allocated':
                return 'l'
            return None
        if k == 'comp':
            bt = self.typeof_derived(e[1])
            if bt and bt in self.typedefs:
                t = self.typedefs[bt].fields.get(e[2])
                return t if t in ('i', 'r', 'd', 'l', 'c') else None
            return None
        if k == 'un':
            if e[1] == '.not.':
                return 'l'
            return self.typeof(e[2])
        if k == 'bin':
            op = e[1]
            if op in ('==', '/=', '<', '<=', '>', '>=', '.and.', '.or.', '.eqv.', '.neqv.'):
                return 'l'
            if op == '//':
                return 'c'
            a, b = self.typeof(e[2]), self.typeof(e[3])
            if op == '**':
                return a if b == 'i' else (max(a, b, key=lambda t: _PROMO[t]) if a in _PROMO and b in _PROMO else None)
            if a in _PROMO and b in _PROMO:
                return max(a, b, key=lambda t: _PROMO[t])
            return None
        return None

    def typeof_derived(self, e):
        if e[0] == 'name':
            s = self.sym(e[1])
            if s and s.typ.startswith('t:'):
                return s.typ[2:]
        return None

    def is_array_name(self, n):
        s = self.sym(n)
        return bool(s and s.dims is not None)

    # ---- expressions
    def idx(self, a):
        """subscript expression -> python 0-based index text."""
        if a[0] == 'slice':
            lo = '' if a[1] is None else f'{self.int_expr(a[1])}-1'
            hi = '' if a[2] is None else self.int_expr(a[2])
            return f'{lo}:{hi}'
        if a[0] == 'num' and self.typeof(a) == 'i':
            return str(int(a[1]) - 1)
        return f'{self.int_expr(a)}-1'

    def int_expr(self, e):
        t = self.typeof(e)
        x = self.ex(e)
        return x if t == 'i' else f'int({x})'

    def ex(self, e):
        k = e[0]
        if k == 'num':
            v = e[1].lower()
            t = self.typeof(e)
            if t == 'i':
                return v
            key = (t, v.replace('d', 'e'))
            if key not in self.consts:
                self.consts[key] = f'_K{len(self.consts)}'
            return self.consts[key]
        if k == 'str':
            return repr(e[1])
        if k == 'log':
            return 'True' if e[1] else 'False'
        if k == 'paren':
            return f'({self.ex(e[1])})'
        if k == 'name':
            return mangle(e[1])
        if k == 'call':
            n, args = e[1], e[2]
            s = self.sym(n)
            if s and s.dims is not None:
                sub = ', '.join(self.idx(a) for a in args)
                has_slice = any(a[0] == 'slice' for a in args)
                if s.typ == 'i' and not has_slice:
                    return f'int({mangle(n)}[{sub}])'
                return f'{mangle(n)}[{sub}]'
            if s and s.typ == 'c' and len(args) == 1 and args[0][0] == 'slice':
                a = args[0]
                lo = '' if a[1] is None else f'{self.int_expr(a[1])}-1'
                hi = '' if a[2] is None else self.int_expr(a[2])
                return f'{mangle(n)}[{lo}:{hi}]'
            if n in self.interfaces:
                return self.c_call(n, args)
            if n in INTRINSIC_FUNCS:
                return f'_in_{n}({", ".join(self.ex(a) for a in args)})'
            # external function (MPI_Wtime ...): provided by the mocks, scalars by value
            return f'{mangle(n)}({", ".join(self.ex(a) for a in args)})'
        if k == 'comp':
            base = self.ex(e[1])
            if e[3] is not None:
                raise Unsupported("function-valued component reference")
            return f'{base}.{mangle(e[2])}'
        if k == 'un':
            if e[1] == '.not.':
                return f'(not {self.ex(e[2])})'
            return f'({e[1]}{self.ex(e[2])})'
        if k == 'bin':
            op, a, b = e[1], e[2], e[3]
            x, y = self.ex(a), self.ex(b)
            if op == '/':
                ta, tb = self.typeof(a), self.typeof(b)
                if ta == 'i' and tb == 'i':
                    return f'_idiv({x}, {y})'
                if ta in ('r', 'd') or tb in ('r', 'd'):
                    return f'({x} / {y})'
                return f'_div({x}, {y})'
            if op == '**':
                return f'_pow({x}, {y})'
            if op == '//':
                return f'(str({x}) + str({y}))'
            pyop = {'/=': '!=', '.and.': 'and', '.or.': 'or', '.eqv.': '==', '.neqv.': '!='}.get(op, op)
            return f'({x} {pyop} {y})'
        raise Unsupported(f"expression node {e}")

    def conv(self, typ, e):
        """text of expression e converted for assignment to a scalar of type typ."""
        x = self.ex(e)
        t = self.typeof(e)
        if typ == t or typ not in ('i', 'r', 'd'):
            return x
        return {'i': f'int({x})', 'r': f'_f4({x})', 'd': f'_f8({x})'}[typ]

    # ---- lvalues
    def assign_to(self, target, value_text, value_type=None):
        """emit  target = value_text  (value_text already evaluated python text)."""
        k = target[0]
        if k == 'name':
            s = self.sym(target[1])
            n = mangle(target[1])
            if s and s.dims is not None:
                self.emit(f'_assign({n}, {value_text})')
            else:
                typ = s.typ if s and not s.pointer else None
                if typ in ('i', 'r', 'd') and typ != value_type:
                    value_text = {'i': f'int({value_text})', 'r': f'_f4({value_text})', 'd': f'_f8({value_text})'}[typ]
                self.emit(f'{n} = {value_text}')
            return
        if k == 'call':
            s = self.sym(target[1])
            if not (s and s.dims is not None):
                raise Unsupported(f"assignment to non-array reference {target[1]}")
            sub = ', '.join(self.idx(a) for a in target[2])
            self.emit(f'{mangle(target[1])}[{sub}] = {value_text}')
            return
        if k == 'comp':
            if target[3] is not None:
                raise Unsupported("assignment to an array component")
            typ = None
            bt = self.typeof_derived(target[1])
            if bt in self.typedefs:
                typ = self.typedefs[bt].fields.get(target[2])
            if typ in ('i', 'r', 'd') and typ != value_type:
                value_text = {'i': f'int({value_text})', 'r': f'_f4({value_text})', 'd': f'_f8({value_text})'}[typ]
            self.emit(f'{self.ex(target[1])}.{mangle(target[2])} = {value_text}')
            return
        raise Unsupported(f"assignment target {target}")

    # ---- calls
    def gen_call(self, callee_text, args, lineno):
        pre, actuals, post = [], [], []
        for a in args:
            if a[0] == 'kw':
                raise Unsupported("keyword actual argument")
            if a[0] == 'name':
                s = self.sym(a[1])
                if s is None:
                    actuals.append(f'Ref({mangle(a[1])})')     # constant supplied by the mocks
                elif s.dims is not None or s.typ.startswith('t:'):
                    actuals.append(mangle(a[1]))
                elif s.param:
                    actuals.append(f'Ref({mangle(a[1])})')
                else:
                    t = self.newtmp('_c')
                    pre.append(f'{t} = Ref({mangle(a[1])})')
                    actuals.append(t)
                    post.append((a, t, s.typ))
            elif a[0] == 'call' and self.is_array_name(a[1]):
                if any(x[0] == 'slice' for x in a[2]):
                    actuals.append(self.ex(a))                 # section: a view
                else:
                    t = self.newtmp('_c')
                    pre.append(f'{t} = Ref({self.ex(a)})')
                    actuals.append(t)
                    post.append((a, t, self.sym(a[1]).typ))
            elif a[0] == 'comp' and a[3] is None:
                t = self.newtmp('_c')
                pre.append(f'{t} = Ref({self.ex(a)})')
                actuals.append(t)
                post.append((a, t, None))
            else:
                actuals.append(f'Ref({self.ex(a)})')
        for p in pre:
            self.emit(p)
        self.emit(f'{callee_text}({", ".join(actuals)})')
        for a, t, typ in post:
            self.assign_to(a, f'{t}.v', None)

    # ---- statements
    def gen_body(self, body):
        """body: list of (lineno, text). Structured translation with an explicit block stack."""
        stack = []   # entries: ('if',) ('do', name) ('select', tmp, first)
        for no, s in body:
            try:
                depth = len(stack)
                self.gen_stmt(no, s, stack)
                if self.post:
                    if len(stack) != depth:
                        raise Unsupported("by-reference C argument in a block header")
                    post, self.post = self.post, []
                    for a, t in post:
                        self.assign_to(a, f'{t}.v', None)
            except Unsupported as ex:
                raise Unsupported(f"{self.u.name} line {no}: {ex}   [{s}]") from None
        if stack:
            raise Unsupported(f"{self.u.name}: unterminated block {stack[-1]}")

    def gen_stmt(self, no, s, stack):
        low = s.lower().strip()
        s = s.strip()
        self.emit(f'# line {no}')
        # construct name prefix   Name: DO ...
        m = re.match(r'^(\w+)\s*:\s*(do\b.*)$', s, re.I)
        cname = None
        if m:
            cname, s = m.group(1).lower(), m.group(2)
            low = s.lower()
        # ---- block ends
        if re.match(r'^end\s*if$', low):
            if not stack or stack[-1][0] != 'if':
                raise Unsupported("END IF without IF")
            stack.pop()
            self.ind -= 1
            return
        if re.match(r'^end\s*do(\s+\w+)?$', low):
            if not stack or stack[-1][0] != 'do':
                raise Unsupported("END DO without DO")
            ent = stack.pop()
            self.ind -= 1
            if ent[2]:
                self.emit('else:')
                self.emit('    ' + ent[2])
            return
        if re.match(r'^end\s*select$', low):
            ent = stack.pop()
            if ent[0] != 'select':
                raise Unsupported("END SELECT without SELECT")
            if not ent[2]:
                self.ind -= 1
            return
        # ---- IF family
        if re.match(r'^else\s*if\s*\(', low):
            st = s.index('(')
            e = _match_paren(s, st)
            if s[e + 1:].strip().lower() != 'then':
                raise Unsupported("ELSE IF without THEN")
            self.ind -= 1
            self.emit(f'elif {self.ex(parse_expr(s[st + 1:e]))}:')
            self.ind += 1
            self.emit('pass')
            return
        if low == 'else':
            self.ind -= 1
            self.emit('else:')
            self.ind += 1
            self.emit('pass')
            return
        if re.match(r'^if\s*\(', low):
            st = s.index('(')
            e = _match_paren(s, st)
            cond = self.ex(parse_expr(s[st + 1:e]))
            tail = s[e + 1:].strip()
            self.emit(f'if {cond}:')
            self.ind += 1
            if tail.lower() == 'then':
                self.emit('pass')
                stack.append(('if',))
            else:
                self.gen_stmt(no, tail, stack)
                self.ind -= 1
            return
        # ---- DO
        if re.match(r'^do$', low):
            self.emit('while True:')
            self.ind += 1
            self.emit('pass')
            stack.append(('do', cname, None))
            return
        m = re.match(r'^do\s+while\s*\(', low)
        if m:
            st = s.index('(')
            e = _match_paren(s, st)
            self.emit(f'while {self.ex(parse_expr(s[st + 1:e]))}:')
            self.ind += 1
            self.emit('pass')
            stack.append(('do', cname, None))
            return
        m = re.match(r'^do\s+(\w+)\s*=\s*(.*)$', s, re.I)
        if m:
            var = m.group(1).lower()
            parts = split_top(m.group(2))
            lo = self.int_expr(parse_expr(parts[0]))
            hi = self.int_expr(parse_expr(parts[1]))
            st = self.int_expr(parse_expr(parts[2])) if len(parts) > 2 else '1'
            a, n = self.newtmp('_lo'), self.newtmp('_n')
            v = mangle(var)
            self.emit(f'{a} = {lo}')
            if st == '1':
                self.emit(f'{n} = max(0, {hi} - {a} + 1)')
                self.emit(f'for {v} in range({a}, {a} + {n}):')
                final = f'{v} = {a} + {n}'
            else:
                stt = self.newtmp('_st')
                self.emit(f'{stt} = {st}')
                self.emit(f'{n} = max(0, _idiv({hi} - {a} + {stt}, {stt}))')
                self.emit(f'for {v} in range({a}, {a} + {n}*{stt}, {stt}):')
                final = f'{v} = {a} + {n}*{stt}'
            self.ind += 1
            self.emit('pass')
            stack.append(('do', cname, final))
            return
        # ---- SELECT CASE
        m = re.match(r'^select\s*case\s*\(', low)
        if m:
            st = s.index('(')
            e = _match_paren(s, st)
            t = self.newtmp('_sel')
            self.emit(f'{t} = {self.ex(parse_expr(s[st + 1:e]))}')
            stack.append(['select', t, True])
            return
        m = re.match(r'^case\s*(default|\()', low)
        if m:
            ent = stack[-1]
            if ent[0] != 'select':
                raise Unsupported("CASE outside SELECT")
            if not ent[2]:
                self.ind -= 1
            if m.group(1) == 'default':
                self.emit('else:' if not ent[2] else 'if True:')
            else:
                st = s.index('(')
                e = _match_paren(s, st)
                conds = []
                for c in split_top(s[st + 1:e]):
                    if ':' in c:
                        lo, _, hi = c.partition(':')
                        cc = []
                        if lo.strip():
                            cc.append(f'{ent[1]} >= {self.ex(parse_expr(lo))}')
                        if hi.strip():
                            cc.append(f'{ent[1]} <= {self.ex(parse_expr(hi))}')
                        conds.append('(' + ' and '.join(cc) + ')')
                    else:
                        conds.append(f'{ent[1]} == {self.ex(parse_expr(c))}')
                self.emit(('if ' if ent[2] else 'elif ') + ' or '.join(conds) + ':')
            ent[2] = False
            self.ind += 1
            self.emit('pass')
            return
        # ---- simple statements
        if low == 'continue':
            return
        if low == 'return':
            self.emit('return')
            return
        if low == 'exit' or low == 'cycle' or re.match(r'^(exit|cycle)\s+\w+$', low):
            w = low.split()
            if len(w) == 2:
                # named: only the innermost loop is supported
                inner = [x for x in stack if x[0] == 'do'][-1]
                if inner[1] != w[1]:
                    raise Unsupported("EXIT / CYCLE of an outer named loop")
            self.emit('break' if w[0] == 'exit' else 'continue')
            return
        m = re.match(r'^stop\b\s*(.*)$', s, re.I)
        if m:
            msg = m.group(1).strip()
            self.emit(f'raise FortranStop({self.ex(parse_expr(msg)) if msg else repr("")}, {no})')
            return
        m = re.match(r'^call\s+(.*)$', s, re.I | re.S)
        if m:
            self.gen_call_stmt(m.group(1).strip(), no)
            return
        if re.match(r'^allocate\s*\(', low):
            st = s.index('(')
            inner = s[st + 1:_match_paren(s, st)]
            for item in split_top(inner):
                mm = re.match(r'^(\w+)\s*\((.*)\)$', item.strip(), re.S)
                if not mm:
                    raise Unsupported(f"ALLOCATE item {item!r}")
                n = mm.group(1).lower()
                sy = self.sym(n)
                dims = ', '.join(self.int_expr(parse_expr(d)) for d in split_top(mm.group(2)))
                self.emit(f'{mangle(n)} = _alloc({sy.typ!r}, ({dims},))')
            return
        if re.match(r'^deallocate\s*\(', low):
            st = s.index('(')
            for item in split_top(s[st + 1:_match_paren(s, st)]):
                n = item.strip().lower()
                self.emit(f'_rt.on_dealloc({n!r}, {mangle(n)})')
                self.emit(f'{mangle(n)} = None')
            return
        if re.match(r'^(open|close|inquire|read|write)\s*\(', low):
            self.gen_io(s, no)
            return
        if re.match(r'^print\b', low):
            items = split_top(s[5:].strip())[1:]
            self.emit(f'_rt.write(None, [{", ".join(self.ex(parse_expr(i)) for i in items)}])')
            return
        # ---- assignment
        depth, q, eq = 0, None, -1
        for i, ch in enumerate(s):
            if q:
                if ch == q:
                    q = None
                continue
            if ch in "'\"":
                q = ch
            elif ch == '(':
                depth += 1
            elif ch == ')':
                depth -= 1
            elif ch == '=' and depth == 0:
                if s[i:i + 2] in ('==', '=>') or s[i - 1] in '/<>=':
                    continue
                eq = i
                break
        if eq > 0:
            target = parse_expr(s[:eq])
            value = parse_expr(s[eq + 1:])
            self.assign_to(target, self.ex(value), self.typeof(value))
            return
        raise Unsupported("statement not in the supported subset")

    def gen_call_stmt(self, text, no):
        p = _P(tokenize(text))
        node = p.primary()
        if not p.at_end():
            raise Unsupported("CALL syntax")
        if node[0] == 'name':
            name, args = node[1], []
        elif node[0] == 'call':
            name, args = node[1], node[2]
        elif node[0] == 'comp':
            # type-bound procedure call  obj%proc(args)
            self.gen_call(f'{self.ex(node[1])}.{mangle(node[2])}', node[3] or [], no)
            return
        else:
            raise Unsupported("CALL target")
        if name in self.rewrites:
            self.rewrites[name](self, args, no)
            return
        if name == 'exit':
            self.emit(f'raise FortranExit({self.ex(args[0]) if args else 0})')
            return
        if name == 'getarg':
            self.assign_to(args[1], f'_rt.getarg({self.ex(args[0])})', 'c')
            return
        if name == 'get_command_argument':
            self.assign_to(args[1], f'_rt.getarg({self.ex(args[0])})', 'c')
            if len(args) > 2:
                self.assign_to(args[2], f'len(_rt.getarg({self.ex(args[0])}))', 'i')
            if len(args) > 3:
                self.assign_to(args[3], '0', 'i')
            return
        self.gen_call(mangle(name), args, no)

    def gen_io(self, s, no):
        st = s.index('(')
        e = _match_paren(s, st)
        kwd = s[:st].strip().lower()
        ctl = split_top(s[st + 1:e])
        tail = s[e + 1:].strip()
        pos, kw = [], {}
        for c in ctl:
            m = re.match(r'^(\w+)\s*=\s*(.*)$', c, re.S)
            if m and not c.strip().startswith("'") and not c.strip().startswith('"') and '==' not in c[:m.end(1) + 3]:
                kw[m.group(1).lower()] = m.group(2).strip()
            else:
                pos.append(c.strip())
        if kwd == 'open':
            unit = pos[0] if pos else kw['unit']
            self.emit(f'_rt.open({self.ex(parse_expr(unit))}, {self.ex(parse_expr(kw["file"]))}, '
                      f'{self.ex(parse_expr(kw.get("action", chr(39) + "READWRITE" + chr(39))))})')
            return
        if kwd == 'close':
            unit = pos[0] if pos else kw['unit']
            self.emit(f'_rt.close({self.ex(parse_expr(unit))})')
            return
        if kwd == 'inquire':
            val = f'_rt.exists({self.ex(parse_expr(kw["file"]))})'
            self.assign_to(parse_expr(kw['exist']), val, 'l')
            return
        unit = pos[0] if pos else kw['unit']
        fmt = pos[1] if len(pos) > 1 else kw.get('fmt', '*')
        items = [parse_expr(i) for i in split_top(tail)] if tail else []
        if fmt.strip() != '*':
            # a format string is supported for WRITE to a character variable only (file names)
            ue = parse_expr(unit) if unit.strip() != '*' else None
            if kwd == 'write' and ue is not None and self.typeof(ue) == 'c':
                vals = ', '.join(self.ex(i) for i in items)
                self.assign_to(ue, f'_rt.fmt_formatted({self.ex(parse_expr(fmt))}, [{vals}])', 'c')
                return
            raise Unsupported("formatted I/O (only list-directed is supported)")
        internal = None
        if unit.strip() != '*':
            ue = parse_expr(unit)
            if self.typeof(ue) == 'c':
                internal = ue
        if kwd == 'write':
            vals = ', '.join(self.ex(i) for i in items)
            if internal is not None:
                self.assign_to(internal, f'_rt.fmt([{vals}])', 'c')
            elif unit.strip() == '*':
                self.emit(f'_rt.write(None, [{vals}])')
            else:
                self.emit(f'_rt.write({self.ex(parse_expr(unit))}, [{vals}])')
            return
        # READ
        types = []
        for it in items:
            t = self.typeof(it)
            if t is None:
                raise Unsupported("READ item of unknown type")
            types.append(t)
        src = f'_rt.read_internal({self.ex(internal)}, {types!r})' if internal is not None else \
            f'_rt.read({self.ex(parse_expr(unit))}, {types!r})'
        v, ios = self.newtmp('_rv'), self.newtmp('_ios')
        self.emit(f'{v}, {ios} = {src}')
        if 'iostat' in kw:
            self.assign_to(parse_expr(kw['iostat']), ios, 'i')
        else:
            self.emit(f'if {ios} != 0: raise FortranStop("READ failed (end of file)", {no})')
        if items:
            self.emit(f'if {ios} == 0:')
            self.ind += 1
            for j, it in enumerate(items):
                self.assign_to(it, f'{v}[{j}]', types[j])
            self.ind -= 1

    # ---- units
    def gen_unit(self):
        u = self.u
        fname = mangle(u.name) if u.kind == 'subroutine' else 'program_' + u.name.lower()
        params = [mangle(a) + '__a' + ('=None' if (u.syms.get(a) and u.syms[a].optional) else '') for a in u.args]
        self.lines.append(f'def {fname}({", ".join(params)}):')
        save_key = f'_save_{fname}'
        saved = []
        # dummies
        scalars_out = []
        for a in u.args:
            s = u.syms.get(a)
            if s is None:
                raise Unsupported(f"{u.name}: dummy {a} is not declared")
            if s.dims is not None or s.typ.startswith('t:'):
                self.emit(f'{mangle(a)} = {mangle(a)}__a')
            elif s.optional:
                self.emit(f'{mangle(a)} = {mangle(a)}__a.v if {mangle(a)}__a is not None else None')
            else:
                self.emit(f'{mangle(a)} = {mangle(a)}__a.v')
                scalars_out.append(a)
        # parameters first (they may size arrays), in declaration order
        for n, s in u.syms.items():
            if s.param:
                self.emit(f'{mangle(n)} = {self.conv(s.typ, parse_expr(s.init))}')
        for n, s in u.syms.items():
            if s.dummy or s.param:
                continue
            if s.typ.startswith('t:'):
                self.emit(f'{mangle(n)} = _new_{s.typ[2:]}()')
            elif s.alloc or s.pointer:
                self.emit(f'{mangle(n)} = None')
            elif s.dims is not None:
                dims = ', '.join(self.int_expr(parse_expr(d)) for d in s.dims)
                if s.init is not None:
                    raise Unsupported("array initialiser")
                self.emit(f'{mangle(n)} = _alloc({s.typ!r}, ({dims},){", zero=True" if self.zero_fill else ""})')
            elif s.init is not None:
                saved.append(n)
                self.emit(f'{mangle(n)} = {save_key}.get({n!r}, _UNSET)')
                self.emit(f'if {mangle(n)} is _UNSET: {mangle(n)} = {self.conv(s.typ, parse_expr(s.init))}')
            else:
                self.emit(f'{mangle(n)} = None')
        self.emit('try:')
        self.ind += 1
        self.emit('pass')
        self.gen_body(u.body)
        if u.kind == 'program':
            self.emit('_rt.on_end(locals())')
        self.ind -= 1
        self.emit('finally:')
        self.ind += 1
        self.emit('pass')
        for a in scalars_out:
            self.emit(f'{mangle(a)}__a.v = {mangle(a)}')
        for n in saved:
            self.emit(f'{save_key}[{n!r}] = {mangle(n)}')
        self.ind -= 1
        head = [f'{save_key} = {{}}']
        return '\n'.join(head + self.lines) + '\n'


def translate(sources: dict, rewrites=None, static_zero_programs: bool = False) -> str:
    """sources: {file name: text} -> python module text (functions of every unit, classes of every TYPE, one ctypes
    binding per BIND(C) interface function).  Files named *.f90 are free-form, everything else fixed-form."""
    all_units, typedefs, mod_syms, interfaces = {}, {}, {}, {}
    per_file = []
    for fn, text in sources.items():
        units, tds, ms, itf = parse_file(text, free_form=fn.lower().endswith('.f90'))
        per_file.append((fn, units, ms))
        typedefs.update(tds)
        interfaces.update(itf)
        mod_syms.update({k: v for k, v in ms.items() if v.param})    # module PARAMETERs are visible to every USEr
        for u in units:
            all_units[u.name] = u
    out = []
    consts = {}
    for name, itf in interfaces.items():
        spec = [(itf.syms[a].typ, itf.syms[a].value, itf.syms[a].dims is not None) for a in itf.args]
        out.append(f'{mangle(name)} = _cbind({name!r}, {spec!r}, {itf.result!r})')
    for fn, units, ms in per_file:
        out.append(f'# ---- {fn}')
        # module-level parameters
        g = Gen(Unit('program', '_mod'), typedefs, ms, all_units, None, consts)
        for n, s in ms.items():
            if s.param:
                out.append(f'{mangle(n)} = {g.conv(s.typ, parse_expr(s.init))}')
        scope = dict(mod_syms)
        scope.update(ms)
        for u in units:
            gen = Gen(u, typedefs, scope, all_units, rewrites, consts, interfaces)
            # gfortran keeps the variables of a main program in static, zero-filled storage; by default this run-time fills
            # them with NaN / a sentinel instead (visible), `static_zero_programs` reproduces the compiled behaviour
            gen.zero_fill = static_zero_programs and u.kind == 'program'
            out.append(gen.gen_unit())
    for td in typedefs.values():
        g = Gen(Unit('program', '_type'), typedefs, mod_syms, all_units, None, consts)
        out.append(f'class _T_{td.name}:')
        out.append('    def __init__(self):')
        for f_, t in td.fields.items():
            init = td.inits.get(f_)
            out.append(f'        self.{mangle(f_)} = {g.conv(t, parse_expr(init)) if init else None}')
        out.append('        pass')
        for p in td.procs:
            dummies = all_units[p].args[1:] if p in all_units else []
            out.append(f'    def {mangle(p)}(self, *a):')
            out.append(f'        return {mangle(p)}(self, *a)')
        out.append(f'def _new_{td.name}():')
        out.append(f'    return _T_{td.name}()')
        out.append('')
    head = ['# GENERATED from the reference sources by oracle/refrun/fortran_to_py.py -- do not commit',
            'from oracle.refrun.runtime import *', '']
    for (t, v), name in consts.items():
        head.append(f'{name} = {"_f8" if t == "d" else "_f4"}({v!r})')
    return '\n'.join(head + out)
