"""ctypes binding of libpfemb200.so: the host-side mirror of the reference's solver interface.

``SolverB200`` carries the same procedures as ``TYPE PetscSolver`` (solverpetsc.F:72-105: initialise, setZero,
free, printInfo, assembleMatrix, assembleVector, assembleMatrixAndVector, factorise, solve,
factoriseAndSolve) plus the batched calls that replace the drivers' element loops.  Every method is a thin
call through the C ABI of include/pfem_b200.h; there is no Python arithmetic on the data path and no
fallback: if the shared library is missing or no B200 is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "libpfemb200.so")

POISSON_TRIA, POISSON_TETRA, ELASTICITY_TRIA, ELASTICITY_TETRA = 0, 1, 2, 3
KIND_DIMS = {0: (3, 1, 2), 1: (4, 1, 3), 2: (3, 2, 2), 3: (4, 3, 3)}   # npElem, ndof, ndim
PC_NONE, PC_JACOBI, PC_BJACOBI_ILU0 = 0, 1, 2     # PC_BJACOBI_ILU0: the reference's default (solverpetsc.F:206)

OK, ERR_CUDA, ERR_ARG, ERR_STATE, ERR_NEG_JACOBIAN, ERR_NCCL, ERR_SIZE, ERR_NUMBERING, ERR_PATTERN = range(9)
SOLVER_EMPTY, PATTERN_OK, INIT_OK, ASSEMBLY_OK, FACTORISE_OK = 1, 2, 3, 4, 5


class PfemError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libpfemb200 status {status}: {message}")
        self.status = status


_lib = None


def load_library() -> C.CDLL:
    """Load the CUDA shared library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise PfemError(ERR_CUDA, f"{LIBPATH} is missing: run `python -m pfemfort_b200.build` (nvcc, sm_100a). "
                                  "There is no CPU fallback.")
    lib = C.CDLL(LIBPATH, mode=C.RTLD_GLOBAL)
    lib.pfem_last_error.restype = C.c_char_p
    lib.pfem_solver_add_value.argtypes = [C.c_void_p, C.c_int, C.c_double]
    lib.pfem_solver_set_options.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.pfem_solver_set_options_from_file.argtypes = [C.c_void_p, C.c_char_p]
    _lib = lib
    return lib


def _chk(status: int):
    if status != OK:
        raise PfemError(status, load_library().pfem_last_error().decode(errors="replace"))


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def device_count() -> int:
    n = C.c_int(0)
    load_library().pfem_device_count(C.byref(n))
    return n.value


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _chk(load_library().pfem_comm_unique_id(buf))
    return buf.raw


def element_ke_batch(kind, x, y, z, elemData, timeData, valC=None):
    """Batched StiffnessResidual* on the GPU.  x,y,z: [npElem, n]; returns K [nsize, nsize, n] with
    K[i, j, e] = Klocal(i+1, j+1), F [nsize, n], jac_neg [n]."""
    lib = load_library()
    npe, ndof, ndim = KIND_DIMS[kind]
    ns = npe * ndof
    x, y = _f64(x), _f64(y)
    n = x.shape[1]
    z = _f64(z) if ndim == 3 else None
    K = np.zeros((ns * ns, n))
    F = np.zeros((ns, n))
    neg = np.zeros(n, np.int32)
    ed = np.zeros(8)
    ed[: len(elemData)] = elemData
    td = np.zeros(8)
    td[: len(timeData)] = timeData
    vc = _f64(valC) if valC is not None else None
    _chk(lib.pfem_element_ke_batch(kind, n, _ptr(x, C.c_double), _ptr(y, C.c_double), _ptr(z, C.c_double),
                                   _ptr(ed, C.c_double), _ptr(td, C.c_double), _ptr(vc, C.c_double),
                                   _ptr(K, C.c_double), _ptr(F, C.c_double), _ptr(neg, C.c_int)))
    # K is stored (i + ns*j) major, element fastest -> [j, i, e] -> transpose to [i, j, e]
    return K.reshape(ns, ns, n).transpose(1, 0, 2).copy(), F, neg


def element_ke(kind, x, y, z, elemData, timeData, valC=None):
    """Single-element entry points (pfem_poisson_tria_ke, ...): returns (K [nsize,nsize], F)."""
    lib = load_library()
    npe, ndof, ndim = KIND_DIMS[kind]
    ns = npe * ndof
    x, y = _f64(x), _f64(y)
    z = _f64(z) if ndim == 3 else None
    K = np.zeros(ns * ns)
    F = np.zeros(ns)
    ed = np.zeros(8)
    ed[: len(elemData)] = elemData
    td = np.zeros(8)
    td[: len(timeData)] = timeData
    vc = _f64(valC) if valC is not None else np.zeros(ns)
    vd = np.zeros(ns)
    d = lambda a: _ptr(a, C.c_double)
    if kind == POISSON_TRIA:
        st = lib.pfem_poisson_tria_ke(d(x), d(y), d(ed), d(td), d(vc), d(vd), d(K), d(F))
    elif kind == POISSON_TETRA:
        st = lib.pfem_poisson_tetra_ke(d(x), d(y), d(z), d(ed), d(td), d(vc), d(vd), d(K), d(F))
    elif kind == ELASTICITY_TRIA:
        st = lib.pfem_elasticity_tria_ke(d(x), d(y), d(ed), d(td), d(vc), d(vd), d(K), d(F))
    else:
        st = lib.pfem_elasticity_tetra_ke(d(x), d(y), d(z), d(ed), d(td), d(vc), d(vd), d(K), d(F))
    _chk(st)
    return K.reshape(ns, ns).T.copy(), F


ASM_AUTO, ASM_ROWS, ASM_FAST = 0, 1, 2


class SolverB200:
    """Drop-in for ``TYPE(PetscSolver)``: one instance per rank / GPU."""

    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1, nccl_id: bytes | None = None):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.rank, self.nranks = rank, nranks
        idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
        _chk(self._lib.pfem_solver_create(C.byref(self._h), device, rank, nranks, idbuf))

    # ---- TYPE PetscSolver procedures (solverpetsc.F:94-103) ----
    def initialise(self, size_local: int, size_global: int, diag_nnz=None, offdiag_nnz=None):
        dn = _i32(diag_nnz) if diag_nnz is not None else None
        on = _i32(offdiag_nnz) if offdiag_nnz is not None else None
        _chk(self._lib.pfem_solver_initialise(self._h, size_local, size_global, _ptr(dn, C.c_int), _ptr(on, C.c_int)))

    def setZero(self):
        _chk(self._lib.pfem_solver_set_zero(self._h))

    def free(self):
        if self._h:
            self._lib.pfem_solver_free(self._h)
            self._h = C.c_void_p()

    def printInfo(self):
        _chk(self._lib.pfem_solver_print_info(self._h))

    def assembleMatrix(self, rindices, cindices, KLOCAL):
        r, c = _i32(rindices), _i32(cindices)
        K = np.asfortranarray(KLOCAL, dtype=np.float64)     # KLOCAL(ii,jj) column-major on the wire
        _chk(self._lib.pfem_solver_assemble_matrix(self._h, r.size, _ptr(r, C.c_int), _ptr(c, C.c_int),
                                                   K.ctypes.data_as(C.POINTER(C.c_double))))

    def assembleVector(self, rindices, FLOCAL):
        r, F = _i32(rindices), _f64(FLOCAL)
        _chk(self._lib.pfem_solver_assemble_vector(self._h, r.size, _ptr(r, C.c_int), _ptr(F, C.c_double)))

    def assembleMatrixAndVector(self, rindices, cindices, KLOCAL, FLOCAL):
        r, c, F = _i32(rindices), _i32(cindices), _f64(FLOCAL)
        K = np.asfortranarray(KLOCAL, dtype=np.float64)
        _chk(self._lib.pfem_solver_assemble_matrix_and_vector(self._h, r.size, _ptr(r, C.c_int), _ptr(c, C.c_int),
                                                              K.ctypes.data_as(C.POINTER(C.c_double)),
                                                              _ptr(F, C.c_double)))

    def factorise(self):
        _chk(self._lib.pfem_solver_factorise(self._h))

    def solve(self):
        _chk(self._lib.pfem_solver_solve(self._h))

    def factoriseAndSolve(self):
        _chk(self._lib.pfem_solver_factorise_and_solve(self._h))

    # ---- options, mesh, pattern, batched value pass ----
    def set_options(self, rtol=-1.0, abstol=-1.0, dtol=-1.0, max_it=-1, pc_type=-1):
        _chk(self._lib.pfem_solver_set_options(self._h, rtol, abstol, dtol, max_it, pc_type))

    def set_options_from_file(self, path="petsc_options.dat"):
        """PetscInitialize(..., "petsc_options.dat") + Set*FromOptions (tetrapoissonparallelimpl1.F:168, solverpetsc.F:190-210)."""
        _chk(self._lib.pfem_solver_set_options_from_file(self._h, os.fsencode(path)))

    def set_mesh(self, kind, conn, coords, node_map_get_old=None):
        conn, coords = _i32(conn), _f64(coords)
        old = _i32(node_map_get_old) if node_map_get_old is not None else None
        _chk(self._lib.pfem_solver_set_mesh(self._h, kind, conn.shape[1], _ptr(conn, C.c_int), coords.shape[1],
                                            _ptr(coords, C.c_double), _ptr(old, C.c_int)))

    def set_pattern(self, elemDof):
        ed = _i32(elemDof)
        _chk(self._lib.pfem_solver_set_pattern(self._h, ed.shape[1], ed.shape[0], _ptr(ed, C.c_int)))

    def set_pattern_nodal(self, NodeDofArrayNew):
        nda = _i32(NodeDofArrayNew)              # [ndof, nNode] = column-major nNode x ndof
        _chk(self._lib.pfem_solver_set_pattern_nodal(self._h, nda.shape[0], _ptr(nda, C.c_int)))

    def set_applied(self, solnApplied):
        sa = _f64(solnApplied)
        _chk(self._lib.pfem_solver_set_applied(self._h, _ptr(sa, C.c_double), sa.size))

    def assemble(self, elemData, timeData, check: bool = True) -> int:
        ed = np.zeros(8)
        ed[: len(elemData)] = elemData
        td = np.zeros(8)
        td[: len(timeData)] = timeData
        neg = C.c_int(0)
        st = self._lib.pfem_solver_assemble(self._h, _ptr(ed, C.c_double), _ptr(td, C.c_double), C.byref(neg))
        if st == ERR_NEG_JACOBIAN and not check:
            return neg.value
        _chk(st)
        return neg.value

    # MatSetValues / VecSetValues / VecSetValue as the drivers call them
    def add_matrix(self, rows, cols, Klocal):
        r, c = _i32(rows), _i32(cols)
        K = np.asfortranarray(Klocal, dtype=np.float64)
        _chk(self._lib.pfem_solver_add_matrix(self._h, r.size, _ptr(r, C.c_int), _ptr(c, C.c_int),
                                              K.ctypes.data_as(C.POINTER(C.c_double))))

    def add_vector(self, rows, F):
        r, F = _i32(rows), _f64(F)
        _chk(self._lib.pfem_solver_add_vector(self._h, r.size, _ptr(r, C.c_int), _ptr(F, C.c_double)))

    def add_value(self, row: int, val: float):
        _chk(self._lib.pfem_solver_add_value(self._h, int(row), float(val)))

    # ---- results ----
    def state(self):
        st, lo, hi, sg = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _chk(self._lib.pfem_solver_get_state(self._h, C.byref(st), C.byref(lo), C.byref(hi), C.byref(sg)))
        return dict(state=st.value, row_start=lo.value, row_end=hi.value, size_global=sg.value)

    def get_csr(self, values: bool = True):
        st = self.state()
        nloc = st["row_end"] - st["row_start"]
        nnz = C.c_longlong()
        _chk(self._lib.pfem_solver_get_nnz(self._h, C.byref(nnz)))
        rowptr = np.zeros(nloc + 1, np.int32)
        col = np.zeros(nnz.value, np.int32)
        val = np.zeros(nnz.value) if values else None
        _chk(self._lib.pfem_solver_get_csr(self._h, _ptr(rowptr, C.c_int), _ptr(col, C.c_int), _ptr(val, C.c_double)))
        return rowptr, col, val

    def get_ilu_factor(self):
        """(fval on the local CSR slots, inverted pivots) of the last PC_BJACOBI_ILU0 solve."""
        st = self.state()
        nnz = C.c_longlong()
        _chk(self._lib.pfem_solver_get_nnz(self._h, C.byref(nnz)))
        fval = np.zeros(nnz.value)
        invd = np.zeros(st["row_end"] - st["row_start"])
        _chk(self._lib.pfem_solver_get_ilu_factor(self._h, _ptr(fval, C.c_double), _ptr(invd, C.c_double)))
        return fval, invd

    def get_rhs(self):
        st = self.state()
        out = np.zeros(st["row_end"] - st["row_start"])
        _chk(self._lib.pfem_solver_get_rhs(self._h, _ptr(out, C.c_double)))
        return out

    def set_rhs(self, rhs):
        r = _f64(rhs)
        _chk(self._lib.pfem_solver_set_rhs(self._h, _ptr(r, C.c_double)))

    def get_solution(self, out=None):
        st = self.state()
        if out is None:
            out = np.zeros(st["size_global"])
        _chk(self._lib.pfem_solver_get_solution(self._h, _ptr(out, C.c_double)))
        return out

    def get_solution_local(self):
        st = self.state()
        out = np.zeros(st["row_end"] - st["row_start"])
        _chk(self._lib.pfem_solver_get_solution_local(self._h, _ptr(out, C.c_double)))
        return out

    def info(self):
        its, reason = C.c_int(), C.c_int()
        rnorm, ta, ts = C.c_double(), C.c_double(), C.c_double()
        _chk(self._lib.pfem_solver_get_info(self._h, C.byref(its), C.byref(reason), C.byref(rnorm), C.byref(ta), C.byref(ts)))
        return dict(its=its.value, reason=reason.value, rnorm=rnorm.value, t_assemble=ta.value, t_solve=ts.value)

    def comm_mode(self) -> int:
        m = C.c_int()
        _chk(self._lib.pfem_solver_comm_mode(self._h, C.byref(m)))
        return m.value

    def assembly_mode(self):
        """(mode, ntiles, visits_per_element) of the last value pass: 0/1 row gather, 2 tiled compute-once."""
        m, n, v = C.c_int(), C.c_int(), C.c_double()
        _chk(self._lib.pfem_solver_assembly_mode(self._h, C.byref(m), C.byref(n), C.byref(v)))
        return m.value, n.value, v.value

    def set_assembly_mode(self, mode: int):
        """ASM_AUTO / ASM_ROWS (default: reference-order no-FMA arithmetic, bit-identical to the sequential evaluation) or
        ASM_FAST (FMA cofactor-form arithmetic, 1e-12 contract)."""
        _chk(self._lib.pfem_solver_set_assembly_mode(self._h, int(mode)))

    def assembly_info(self):
        """Kernel of the last value pass, its FP64 instructions per element visit (counted from the SASS of this
        build, profiles/sass_*.txt), the element visits of one pass and the arithmetic mode."""
        name = C.create_string_buffer(256)
        fp64, visits, arith = C.c_double(), C.c_longlong(), C.c_int()
        _chk(self._lib.pfem_solver_assembly_info(self._h, name, 256, C.byref(fp64), C.byref(visits), C.byref(arith)))
        return dict(kernel=name.value.decode(), fp64_per_visit=fp64.value, visits=visits.value,
                    arith={0: "no-FMA (reference evaluation order)", 1: "FMA (1e-12 contract)"}.get(arith.value, str(arith.value)))

    def launch_count(self, reset: bool = False) -> int:
        n = C.c_longlong()
        _chk(self._lib.pfem_solver_launch_count(self._h, C.byref(n), 1 if reset else 0))
        return n.value

    def time_spmv(self, reps: int = 20) -> float:
        s = C.c_double()
        _chk(self._lib.pfem_solver_time_spmv(self._h, reps, C.byref(s)))
        return s.value

    def set_profiling(self, on: bool):
        _chk(self._lib.pfem_solver_set_profiling(self._h, 1 if on else 0))

    def get_profile(self):
        s, n = C.c_double(), C.c_longlong()
        _chk(self._lib.pfem_solver_get_profile(self._h, C.byref(s), C.byref(n)))
        return s.value, n.value

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
