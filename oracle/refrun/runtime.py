"""Run-time support of the Fortran -> Python translation (oracle/refrun/fortran_to_py.py).  TEST INFRASTRUCTURE ONLY.

Fortran semantics kept here: typed scalars (int / float32 / float64), truncating integer division, integer powers by
repeated multiplication, MATMUL in gfortran's inline order, list-directed READ, NaN / sentinel fill of fresh arrays,
by-reference scalar arguments (`Ref`).  One `Runtime` per simulated MPI rank (a thread), reached through `_rt`.
"""
from __future__ import annotations

import os
import threading

import numpy as np

__all__ = ['np', '_assign', '_cbind', '_AttrRef', 'set_clib', 'c_null_ptr', 'c_null_char', 'Ref', 'FortranStop', 'FortranExit', '_UNSET', '_f4', '_f8', '_idiv', '_div', '_pow', '_alloc', '_rt',
           'Runtime', 'INT_SENTINEL'] + [
    '_in_' + n for n in ('min', 'max', 'abs', 'sqrt', 'acos', 'asin', 'atan', 'cos', 'sin', 'tan', 'exp', 'log', 'dble',
                         'real', 'int', 'nint', 'mod', 'size', 'matmul', 'transpose', 'trim', 'adjustl', 'len_trim',
                         'iargc', 'command_argument_count', 'allocated', 'null', 'float', 'sum', 'dot_product',
                         'sign', 'floor', 'ceiling', 'maxval', 'minval', 'present')]

_f4 = np.float32
_f8 = np.float64
_UNSET = object()
INT_SENTINEL = -2139062144   # 0x80808080: an uninitialised INTEGER array element


class FortranStop(Exception):
    def __init__(self, msg='', line=0):
        super().__init__(f"STOP {msg!r} (source line {line})")
        self.msg, self.line = msg, line


class FortranExit(Exception):
    def __init__(self, code=0):
        super().__init__(f"EXIT({code})")
        self.code = code


class Ref:
    """A scalar actual argument, passed by reference."""
    __slots__ = ('v',)

    def __init__(self, v=None):
        self.v = v

    def __repr__(self):
        return f'Ref({self.v!r})'


def _idiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _div(a, b):
    if isinstance(a, int) and isinstance(b, int):
        return _idiv(a, b)
    return a / b


def _powi(x, n):
    """x**n for integer n >= 0 the way GCC expands __builtin_powi: binary powering (x**2 = x*x, x**3 = x*x*x,
    x**4 = (x*x)*(x*x))."""
    if n == 0:
        return type(x)(1)
    r = None
    base = x
    while n:
        if n & 1:
            r = base if r is None else r * base
        n >>= 1
        if n:
            base = base * base
    return r


def _pow(a, b):
    if isinstance(b, int):
        if isinstance(a, int):
            return a ** b if b >= 0 else (1 if a == 1 else 0)
        return _powi(a, b) if b >= 0 else type(a)(1) / _powi(a, -b)
    return np.power(a, b)


def _alloc(typ, shape, zero=False):
    if zero and typ in ('i', 'd', 'r'):
        return np.zeros(shape, dtype={'i': np.int64, 'd': np.float64, 'r': np.float32}[typ], order='F')
    if typ == 'i':
        return np.full(shape, INT_SENTINEL, dtype=np.int64, order='F')
    if typ == 'd':
        return np.full(shape, np.nan, dtype=np.float64, order='F')
    if typ == 'r':
        return np.full(shape, np.nan, dtype=np.float32, order='F')
    if typ == 'l':
        return np.zeros(shape, dtype=bool, order='F')
    if typ == 'c':
        return np.full(shape, b'?', dtype='S1', order='F')
    raise TypeError(f"array of type {typ!r}")


def _assign(dst, v):
    """whole-array assignment.  Conforming shapes: plain copy.  A rank-1 source of a DIFFERENT length (non-conforming
    Fortran, e.g. `forAssyVec(6) = elemNodeConn(ee,:)` with three columns in triaelasticityparallelimpl1.F) copies the
    common leading part, fills the rest with the uninitialised sentinel and leaves a note on the run."""
    if isinstance(v, np.ndarray) and v.ndim == 1 and dst.ndim == 1 and v.shape != dst.shape:
        n = min(v.size, dst.size)
        dst[:n] = v[:n]
        dst[n:] = INT_SENTINEL if np.issubdtype(dst.dtype, np.integer) else np.nan
        _rt.notes.append(f"non-conforming array assignment: {v.size} elements into {dst.size}")
    else:
        dst[...] = v


# ---- ISO_C_BINDING: BIND(C) interface functions are bound to a shared library with ctypes -----------------------------

c_null_ptr = None
c_null_char = '\0'
_CLIB = [None]


def set_clib(lib):
    """the shared library the BIND(C) interface functions resolve in (a ctypes.CDLL)."""
    _CLIB[0] = lib


class _AttrRef:
    """a derived-type component passed by reference."""
    __slots__ = ('o', 'a')

    def __init__(self, o, a):
        self.o, self.a = o, a

    @property
    def v(self):
        return getattr(self.o, self.a)

    @v.setter
    def v(self, val):
        setattr(self.o, self.a, val)


def _cbind(name, spec, result):
    """python callable for `<result> FUNCTION name(...) BIND(C)`; spec = [(type, VALUE?, array?)] per dummy.
    Marshalling is what a Fortran processor does for these declarations: VALUE scalars by value (c_int / c_double /
    void*), other scalars by address (copied back), arrays by address of contiguous column-major storage of the declared C
    kind (INTEGER arrays of this run-time are 64-bit: converted to C_INT on the way in and copied back on the way out),
    CHARACTER arrays / strings as char*."""
    import ctypes as C

    def call(*args):
        lib = _CLIB[0]
        if lib is None:
            raise RuntimeError(f"{name}: no library bound (runtime.set_clib)")
        fn = getattr(lib, name)
        fn.restype = {'i': C.c_int, 'd': C.c_double, None: C.c_int}[result]
        cargs, after = [], []
        if len(args) != len(spec):
            raise TypeError(f"{name}: {len(args)} arguments for {len(spec)} dummies")
        for a, (typ, value, is_array) in zip(args, spec):
            if value:
                v = a.v if isinstance(a, (Ref, _AttrRef)) else a
                cargs.append({'i': C.c_int, 'd': C.c_double, 'h': C.c_void_p}[typ](v if typ == 'h' else (int(v) if typ == 'i' else float(v))))
            elif is_array:
                if typ == 'c':
                    raw = a.encode() if isinstance(a, str) else (np.asarray(a).tobytes() if a is not None else b'')
                    buf = C.create_string_buffer(raw, max(len(raw), 1) + 1)
                    cargs.append(buf)
                elif a is None:
                    cargs.append(None)
                else:
                    dt = np.int32 if typ == 'i' else np.float64
                    arr = np.asarray(a)
                    tmp = np.asfortranarray(arr, dtype=dt)
                    cargs.append(tmp.ctypes.data_as(C.c_void_p))
                    if tmp is not arr and not np.shares_memory(tmp, arr):
                        after.append((arr, tmp))
                    after.append((None, tmp))           # keep alive
            else:
                ct = {'i': C.c_int, 'd': C.c_double, 'h': C.c_void_p}[typ]
                cur = a.v
                box = ct() if cur is None else ct(cur if typ == 'h' else (int(cur) if typ == 'i' else float(cur)))
                cargs.append(C.byref(box))
                after.append((a, box))
        rc = fn(*cargs)
        for dst, src in after:
            if dst is None:
                continue
            if isinstance(dst, np.ndarray):
                dst[...] = src
            else:
                dst.v = src.value
        return int(rc) if result in ('i', None) else np.float64(rc)

    call.__name__ = name
    return call


# ---- intrinsics ---------------------------------------------------------------------------------------------------

def _promote(args):
    if all(isinstance(a, int) for a in args):
        return args
    t = np.result_type(*[a for a in args if not isinstance(a, int)])
    return [t.type(a) for a in args]


def _in_min(*a):
    a = _promote(a)
    r = a[0]
    for x in a[1:]:
        if x < r:
            r = x
    return r


def _in_max(*a):
    a = _promote(a)
    r = a[0]
    for x in a[1:]:
        if x > r:
            r = x
    return r


def _in_abs(x):
    return abs(x)


def _in_sign(a, b):
    return abs(a) if b >= 0 else -abs(a)


def _in_sqrt(x):
    return np.sqrt(x)


def _in_acos(x):
    return np.arccos(x)


def _in_asin(x):
    return np.arcsin(x)


def _in_atan(x):
    return np.arctan(x)


def _in_cos(x):
    return np.cos(x)


def _in_sin(x):
    return np.sin(x)


def _in_tan(x):
    return np.tan(x)


def _in_exp(x):
    return np.exp(x)


def _in_log(x):
    return np.log(x)


def _in_dble(x):
    return np.float64(x)


def _in_real(x):
    return np.float32(x)


_in_float = _in_real


def _in_int(x):
    return int(x)


def _in_nint(x):
    return int(np.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


def _in_floor(x):
    return int(np.floor(x))


def _in_ceiling(x):
    return int(np.ceil(x))


def _in_mod(a, b):
    if isinstance(a, int) and isinstance(b, int):
        return a - _idiv(a, b) * b
    return np.fmod(a, b)


def _in_size(a, dim=None):
    return int(a.size) if dim is None else int(a.shape[dim - 1])


def _in_allocated(a):
    return a is not None


def _in_null():
    return None


def _in_present(a):
    return a is not None


def _in_trim(s):
    return str(s).rstrip(' ')


def _in_adjustl(s):
    return str(s).lstrip(' ')


def _in_len_trim(s):
    return len(str(s).rstrip(' '))


def _in_iargc():
    return _rt.iargc()


_in_command_argument_count = _in_iargc


def _in_transpose(a):
    return np.array(a.T, order='F')


def _in_matmul(a, b):
    """gfortran's inline MATMUL: c = 0; DO j; DO l; DO i: c(i,j) = c(i,j) + a(i,l)*b(l,j)  (ascending l from zero)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.ndim == 2 and b.ndim == 1:
        c = np.zeros(a.shape[0], dtype=np.result_type(a, b))
        for l in range(a.shape[1]):
            c = c + a[:, l] * b[l]
        return c
    if a.ndim == 1 and b.ndim == 2:
        c = np.zeros(b.shape[1], dtype=np.result_type(a, b))
        for l in range(a.shape[0]):
            c = c + a[l] * b[l, :]
        return c
    c = np.zeros((a.shape[0], b.shape[1]), dtype=np.result_type(a, b), order='F')
    for l in range(a.shape[1]):
        c = c + a[:, l:l + 1] * b[l:l + 1, :]
    return c


def _in_sum(a):
    r = a.dtype.type(0)
    for x in np.asarray(a).ravel(order='F'):
        r = r + x
    return int(r) if np.issubdtype(a.dtype, np.integer) else r


def _in_dot_product(a, b):
    r = np.result_type(a, b).type(0)
    for x, y in zip(a, b):
        r = r + x * y
    return r


def _in_maxval(a):
    v = a.max()
    return int(v) if np.issubdtype(a.dtype, np.integer) else v


def _in_minval(a):
    v = a.min()
    return int(v) if np.issubdtype(a.dtype, np.integer) else v


# ---- per-rank run-time state: command line, units, captures ------------------------------------------------------

def _parse_value(tok, typ):
    if typ == 'i':
        return int(tok)
    if typ in ('d', 'r'):
        v = float(tok.lower().replace('d', 'e'))      # correctly rounded, like libgfortran's reader
        return np.float64(v) if typ == 'd' else np.float32(v)
    if typ == 'l':
        return tok.strip('.').lower().startswith('t')
    return tok


def _record_tokens(line):
    return line.replace(',', ' ').split()


class Runtime:
    def __init__(self, argv=(), cwd='.', rank=0, world=None, quiet=True):
        self.argv = list(argv)          # argv[0] = program name
        self.cwd = cwd
        self.rank = rank
        self.world = world
        self.quiet = quiet
        self.units = {}                 # unit -> dict(mode, lines, pos, records)
        self.written = {}               # file name -> list of records (each a list of values)
        self.stdout = []
        self.final_arrays = {}          # name -> last value of a deallocated array
        self.final_locals = {}
        self.seq = 0                    # creation counter of collective objects
        self.notes = []                 # undefined-behaviour notes (non-conforming assignments ...)

    # command line
    def iargc(self):
        return len(self.argv) - 1

    def getarg(self, i):
        return self.argv[i] if 0 <= i < len(self.argv) else ''

    # files
    def _path(self, name):
        name = str(name).strip()
        return name if os.path.isabs(name) else os.path.join(self.cwd, name)

    def exists(self, name):
        return os.path.exists(self._path(name))

    def open(self, unit, name, action='READWRITE'):
        action = str(action).upper()
        if action == 'READ':
            with open(self._path(name)) as f:
                self.units[unit] = {'mode': 'r', 'lines': f.read().splitlines(), 'pos': 0}
        else:
            rec = self.written.setdefault(str(name).strip(), [])
            del rec[:]
            self.units[unit] = {'mode': 'w', 'records': rec}

    def close(self, unit):
        self.units.pop(unit, None)

    def read(self, unit, types):
        """one list-directed READ: consumes records until len(types) values were found (at least one record)."""
        u = self.units[unit]
        vals = []
        first = True
        while first or len(vals) < len(types):
            if u['pos'] >= len(u['lines']):
                return None, -1
            toks = _record_tokens(u['lines'][u['pos']])
            u['pos'] += 1
            first = False
            vals.extend(toks)
            if not types:
                break
        try:
            return [_parse_value(v, t) for v, t in zip(vals, types)], 0
        except ValueError:
            return None, 5010

    def read_internal(self, s, types):
        toks = _record_tokens(str(s))
        if len(toks) < len(types):
            return None, -1
        return [_parse_value(v, t) for v, t in zip(toks, types)], 0

    def fmt(self, items):
        return ' '.join(str(i) for i in items)

    def fmt_formatted(self, fmt, items):
        """edit descriptors A, Iw, Iw.m (enough for the drivers' output file names)."""
        import re
        out, it = [], iter(items)
        for d in fmt.strip().strip('()').split(','):
            d = d.strip().upper()
            m = re.match(r'^I(\d+)(?:\.(\d+))?$', d)
            if d.startswith('A'):
                out.append(str(next(it, '')))
            elif m:
                v = next(it, None)
                if v is None:
                    continue
                body = str(abs(int(v))).zfill(int(m.group(2) or 0))
                out.append((('-' if int(v) < 0 else '') + body).rjust(int(m.group(1))))
            else:
                raise NotImplementedError(f"edit descriptor {d!r}")
        return ''.join(out)

    def write(self, unit, items):
        if unit is None or unit not in self.units:
            self.stdout.append(list(items))
            if not self.quiet:
                print(f'[rank {self.rank}]', *items)
            return
        self.units[unit]['records'].append(list(items))

    # captures
    def on_dealloc(self, name, value):
        if value is not None:
            self.final_arrays[name] = np.array(value, copy=True)

    def on_end(self, local_vars):
        for k, v in local_vars.items():
            if k.startswith('_'):
                continue
            if isinstance(v, np.ndarray):
                self.final_arrays.setdefault(k, np.array(v, copy=True))
            elif isinstance(v, (int, float, bool, str, np.floating)) and not isinstance(v, type):
                self.final_locals[k] = v


class _RtProxy:
    """`_rt` in generated code: the Runtime of the calling thread (= simulated rank)."""
    _tls = threading.local()

    def bind(self, rt):
        self._tls.rt = rt

    def current(self):
        rt = getattr(self._tls, 'rt', None)
        if rt is None:
            rt = self._tls.rt = Runtime()
        return rt

    def __getattr__(self, name):
        return getattr(self.current(), name)


_rt = _RtProxy()
