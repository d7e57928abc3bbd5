"""What is EXTERNAL to the reference, supplied to the executed Fortran: MPI, PETSc 3.6 (Vec / Mat / KSP / PC), METIS and
the reference's VTK writer.  TEST INFRASTRUCTURE ONLY.

These are not in /root/reference (third-party, pinned by path in CMakeLists.txt:43).  Their published semantics are
restated here, minimally, so that the reference's own driver and solver-wrapper source can run:

* MPI: P ranks = P threads of one process; collectives exchange through a shared `World` (barrier + slots).
* PETSc Mat (MPIAIJ): MatSetValues takes the logically two-dimensional `v` ROW-major (v[i*n+j] <-> idxm[i], idxn[j]),
  ignores negative indices, applies INSERT / ADD to locally owned rows at once and stashes rows of other ranks until
  MatAssemblyEnd, where stashed entries arrive in source-rank order; inserted zeros stay in the pattern;
  MatZeroEntries keeps the pattern.  Vec likewise (negative indices ignored because the reference sets
  VEC_IGNORE_NEGATIVE_INDICES).  Every call is recorded (`world.trace`) so tests can assert the reference's options
  (KSPCG, PCBJACOBI ...) and call order.
* KSPSolve: the Krylov solver itself is PETSc's; here the assembled system is solved by a sparse direct solve (scipy),
  good to rounding, and the converged reason is reported as 2 (rtol).  The solve is NOT what these runs pin.
* METIS_PartMeshNodal: returns the element / node partition the harness supplies (any valid partition is a legal
  METIS answer; METIS itself is third-party).
* writeoutputvtk: records its arguments (the VTK writer is outside the hot path).
"""
from __future__ import annotations

import threading
import time

import numpy as np

from .runtime import Ref, _rt


def _val(x):
    return x.v if isinstance(x, Ref) else x


class World:
    def __init__(self, size, partition=None):
        self.size = size
        self.barrier = threading.Barrier(size)
        self.slots = [None] * size
        self.objects = {}
        self.lock = threading.Lock()
        self.trace = []                # (rank, call name, info)
        self.partition = partition     # (elem_proc_id, node_proc_id) 0-based part numbers
        self.vtk = None
        self.printed = []

    def log(self, name, info=None):
        with self.lock:
            self.trace.append((_rt.rank, name, info))

    def exchange(self, value):
        """all-gather of python objects between the rank threads."""
        r = _rt.rank
        self.slots[r] = value
        self.barrier.wait()
        out = list(self.slots)
        self.barrier.wait()
        return out

    def collective_object(self, factory):
        """the same object on every rank for the k-th collective creation call."""
        rt = _rt.current()
        key = rt.seq
        rt.seq += 1
        with self.lock:
            if key not in self.objects:
                self.objects[key] = factory()
        return self.objects[key]


def _world() -> World:
    return _rt.world


# ---- MPI ------------------------------------------------------------------------------------------------------------

PETSC_COMM_WORLD = 'PETSC_COMM_WORLD'
MPI_INT, MPI_DOUBLE, MPI_SUM, MPI_MAX, MPI_MIN = 'MPI_INT', 'MPI_DOUBLE', 'MPI_SUM', 'MPI_MAX', 'MPI_MIN'


def mpi_comm_size(comm, n, err):
    n.v = _world().size
    err.v = 0


def mpi_comm_rank(comm, r, err):
    r.v = _rt.rank
    err.v = 0


def mpi_barrier(comm, err):
    _world().barrier.wait()
    err.v = 0


def mpi_wtime():
    return np.float64(time.perf_counter())


def mpi_bcast(buf, count, dtype, root, comm, err):
    w = _world()
    n, root = _val(count), _val(root)
    src = w.exchange(buf if _rt.rank == root else None)[root]
    if _rt.rank != root:
        if isinstance(buf, Ref):
            buf.v = src.v
        else:
            buf.reshape(-1, order='F')[:n] = src.reshape(-1, order='F')[:n]
    w.barrier.wait()
    err.v = 0


def mpi_allgather(sbuf, scount, stype, rbuf, rcount, rtype, comm, err):
    w = _world()
    n = _val(scount)
    mine = [sbuf.v] if isinstance(sbuf, Ref) else list(sbuf.reshape(-1, order='F')[:n])
    parts = w.exchange(mine)
    flat = [x for p in parts for x in p]
    rbuf.reshape(-1, order='F')[:len(flat)] = flat
    err.v = 0


def mpi_allgatherv(sbuf, scount, stype, rbuf, rcounts, displs, rtype, comm, err):
    w = _world()
    n = _val(scount)
    mine = [sbuf.v] if isinstance(sbuf, Ref) else list(sbuf.reshape(-1, order='F')[:n])
    parts = w.exchange(mine)
    flat = rbuf.reshape(-1, order='F')
    for r, p in enumerate(parts):
        d, c = int(displs[r]), int(rcounts[r])
        assert c == len(p), "MPI_Allgatherv: recvcounts disagree with what rank %d sent" % r
        flat[d:d + c] = p
    err.v = 0


def mpi_allreduce(sbuf, rbuf, count, dtype, op, comm, err):
    w = _world()
    op = _val(op)
    mine = sbuf.v if isinstance(sbuf, Ref) else np.array(sbuf.reshape(-1, order='F')[:_val(count)])
    parts = w.exchange(mine)
    red = parts[0]
    for p in parts[1:]:
        red = {'MPI_SUM': lambda a, b: a + b, 'MPI_MAX': np.maximum, 'MPI_MIN': np.minimum}[op](red, p)
    if isinstance(rbuf, Ref):
        rbuf.v = int(red) if isinstance(mine, int) else red
    else:
        rbuf.reshape(-1, order='F')[:_val(count)] = red
    err.v = 0


# ---- PETSc ----------------------------------------------------------------------------------------------------------

INSERT_VALUES, ADD_VALUES = 'INSERT_VALUES', 'ADD_VALUES'
MAT_FINAL_ASSEMBLY, MAT_FLUSH_ASSEMBLY = 'MAT_FINAL_ASSEMBLY', 'MAT_FLUSH_ASSEMBLY'
PETSC_TRUE, PETSC_FALSE = True, False
PETSC_NULL_INTEGER = PETSC_NULL_OBJECT = PETSC_NULL_CHARACTER = None
PETSC_DECIDE = PETSC_DETERMINE = -1
VEC_IGNORE_NEGATIVE_INDICES = 'VEC_IGNORE_NEGATIVE_INDICES'
MAT_NEW_NONZERO_ALLOCATION_ERR = 'MAT_NEW_NONZERO_ALLOCATION_ERR'
MAT_NEW_NONZERO_LOCATIONS = 'MAT_NEW_NONZERO_LOCATIONS'
MAT_KEEP_NONZERO_PATTERN = 'MAT_KEEP_NONZERO_PATTERN'
KSPCG, KSPGMRES, KSPBCGS, KSPPREONLY = 'cg', 'gmres', 'bcgs', 'preonly'
PCBJACOBI, PCJACOBI, PCNONE, PCILU, PCLU = 'bjacobi', 'jacobi', 'none', 'ilu', 'lu'
SCATTER_FORWARD, SCATTER_REVERSE = 'SCATTER_FORWARD', 'SCATTER_REVERSE'


class MockVec:
    def __init__(self):
        self.n_local = None      # per rank
        self.N = None
        self.starts = None
        self.data = None
        self.stash = None        # per rank list of (idx, val, mode)
        self.options = {}
        self.lock = threading.Lock()

    def owner_range(self, r):
        return self.starts[r], self.starts[r + 1]


class MockMat:
    def __init__(self):
        self.N = None
        self.starts = None
        self.rows = None         # list of dict col -> value, global rows
        self.stash = None
        self.options = {}
        self.prealloc = {}
        self.assembled = 0
        self.lock = threading.Lock()

    def csr(self):
        rowptr = [0]
        col, val = [], []
        for r in self.rows:
            ks = sorted(r)
            col.extend(ks)
            val.extend(r[k] for k in ks)
            rowptr.append(len(col))
        return (np.array(rowptr, dtype=np.int64), np.array(col, dtype=np.int64),
                np.array(val, dtype=np.float64))


class MockKSP:
    def __init__(self):
        self.type = None
        self.pc = MockPC()
        self.mat = None
        self.its = 0
        self.reason = 0
        self.solves = 0


class MockPC:
    def __init__(self):
        self.type = None


def petscinitialize(fname, err):
    _world().log('PetscInitialize', _val(fname))
    err.v = 0


def petscfinalize(err):
    _world().log('PetscFinalize')
    err.v = 0


def petscprintf(comm, s, err):
    if _rt.rank == 0:
        _world().printed.append(str(_val(s)))
    err.v = 0


def _set_sizes(obj, n_local, N):
    w = _world()
    locs = w.exchange(int(n_local))
    starts = [0]
    for x in locs:
        starts.append(starts[-1] + x)
    assert starts[-1] == int(N), f"sum of local sizes {starts[-1]} /= global size {N}"
    return starts


def veccreate(comm, v, err):
    v.v = _world().collective_object(MockVec)
    _world().log('VecCreate')
    err.v = 0


def vecsetsizes(v, n, N, err):
    vec = _val(v)
    starts = _set_sizes(vec, _val(n), _val(N))
    with vec.lock:
        if vec.data is None:
            vec.N, vec.starts = int(_val(N)), starts
            vec.data = np.zeros(vec.N)
            vec.stash = [[] for _ in range(_world().size)]
    _world().barrier.wait()
    err.v = 0


def vecsetfromoptions(v, err):
    err.v = 0


def vecduplicate(v, out, err):
    src = _val(v)

    def make():
        d = MockVec()
        d.N, d.starts = src.N, list(src.starts)
        d.data = np.zeros(src.N)
        d.stash = [[] for _ in range(_world().size)]
        return d
    out.v = _world().collective_object(make)
    _world().log('VecDuplicate')
    err.v = 0


def vecsetoption(v, opt, flag, err):
    _val(v).options[_val(opt)] = _val(flag)
    _world().log('VecSetOption', (_val(opt), _val(flag)))
    err.v = 0


def _vec_add(vec, idx, val, mode):
    r = _rt.rank
    if idx < 0:
        if vec.options.get(VEC_IGNORE_NEGATIVE_INDICES):
            return
        raise IndexError("VecSetValues: negative index without VEC_IGNORE_NEGATIVE_INDICES")
    if idx >= vec.N:
        raise IndexError(f"VecSetValues: index {idx} out of range {vec.N}")
    lo, hi = vec.owner_range(r)
    if lo <= idx < hi:
        if mode == ADD_VALUES:
            vec.data[idx] = vec.data[idx] + val
        else:
            vec.data[idx] = val
    else:
        vec.stash[r].append((idx, val, mode))


def vecsetvalues(v, n, idx, vals, mode, err):
    vec, mode = _val(v), _val(mode)
    vv = vals.reshape(-1, order='F')
    ii = idx.reshape(-1, order='F')
    for k in range(_val(n)):
        _vec_add(vec, int(ii[k]), np.float64(vv[k]), mode)
    err.v = 0


def vecsetvalue(v, idx, val, mode, err=None):
    _vec_add(_val(v), int(_val(idx)), np.float64(_val(val)), _val(mode))
    if err is not None:
        err.v = 0


def vecassemblybegin(v, err):
    err.v = 0


def vecassemblyend(v, err):
    vec = _val(v)
    w = _world()
    w.barrier.wait()
    if _rt.rank == 0:
        for src in range(w.size):
            for idx, val, mode in vec.stash[src]:
                if mode == ADD_VALUES:
                    vec.data[idx] = vec.data[idx] + val
                else:
                    vec.data[idx] = val
            vec.stash[src] = []
    w.barrier.wait()
    err.v = 0


def veczeroentries(v, err):
    w = _world()
    w.barrier.wait()
    if _rt.rank == 0:
        _val(v).data[:] = 0.0
    w.barrier.wait()
    err.v = 0


def vecdestroy(v, err):
    err.v = 0


def matcreate(comm, m, err):
    m.v = _world().collective_object(MockMat)
    _world().log('MatCreate')
    err.v = 0


def matsetsizes(m, nl, ml, N, M, err):
    mat = _val(m)
    starts = _set_sizes(mat, _val(nl), _val(N))
    with mat.lock:
        if mat.rows is None:
            mat.N, mat.starts = int(_val(N)), starts
            mat.rows = [dict() for _ in range(mat.N)]
            mat.stash = [[] for _ in range(_world().size)]
    _world().barrier.wait()
    err.v = 0


def matsetfromoptions(m, err):
    err.v = 0


def matmpiaijsetpreallocation(m, dnz, dnnz, onz, onnz, err):
    _val(m).prealloc[_rt.rank] = (int(_val(dnz)), np.array(dnnz), int(_val(onz)), np.array(onnz))
    _world().log('MatMPIAIJSetPreallocation', (int(_val(dnz)), int(_val(onz))))
    err.v = 0


def matseqaijsetpreallocation(m, nz, nnz, err):
    _world().log('MatSeqAIJSetPreallocation', int(_val(nz)))
    err.v = 0


def matsetoption(m, opt, flag, err):
    _val(m).options[_val(opt)] = _val(flag)
    _world().log('MatSetOption', (_val(opt), _val(flag)))
    err.v = 0


def _mat_set(mat, row, col, val, mode):
    if row < 0 or col < 0:
        return                                  # PETSc: negative indices are ignored
    if row >= mat.N or col >= mat.N:
        raise IndexError(f"MatSetValues: ({row},{col}) outside {mat.N}")
    r = _rt.rank
    if mat.starts[r] <= row < mat.starts[r + 1]:
        d = mat.rows[row]
        if mode == ADD_VALUES and col in d:
            d[col] = d[col] + val
        else:
            d[col] = val                        # INSERT, or the first ADD at a new location
    else:
        mat.stash[r].append((row, col, val, mode))


def matsetvalues(m, nr, idxm, nc, idxn, v, mode, err):
    mat, mode = _val(m), _val(mode)
    nr, nc = _val(nr), _val(nc)
    vv = v.reshape(-1, order='F')               # the memory of the Fortran array, read ROW-major by PETSc
    im = idxm.reshape(-1, order='F')
    jn = idxn.reshape(-1, order='F')
    for i in range(nr):
        row = int(im[i])
        for j in range(nc):
            _mat_set(mat, row, int(jn[j]), np.float64(vv[i * nc + j]), mode)
    err.v = 0


def matsetvalue(m, row, col, val, mode, err=None):
    _mat_set(_val(m), int(_val(row)), int(_val(col)), np.float64(_val(val)), _val(mode))
    if err is not None:
        err.v = 0


def matassemblybegin(m, kind, err):
    err.v = 0


def matassemblyend(m, kind, err):
    mat = _val(m)
    w = _world()
    w.barrier.wait()
    if _rt.rank == 0:
        for src in range(w.size):
            for row, col, val, mode in mat.stash[src]:
                d = mat.rows[row]
                if mode == ADD_VALUES and col in d:
                    d[col] = d[col] + val
                else:
                    d[col] = val
            mat.stash[src] = []
        mat.assembled += 1
    w.barrier.wait()
    err.v = 0


def matzeroentries(m, err):
    w = _world()
    w.barrier.wait()
    if _rt.rank == 0:
        for d in _val(m).rows:
            for k in d:
                d[k] = np.float64(0.0)
    w.barrier.wait()
    err.v = 0


def matdestroy(m, err):
    err.v = 0


def kspcreate(comm, k, err):
    k.v = _world().collective_object(MockKSP)
    _world().log('KSPCreate')
    err.v = 0


def kspsetoperators(k, a, p, err):
    _val(k).mat = _val(a)
    err.v = 0


def kspsettype(k, t, err):
    _val(k).type = _val(t)
    _world().log('KSPSetType', _val(t))
    err.v = 0


def kspsetfromoptions(k, err):
    err.v = 0


def kspgetpc(k, pc, err):
    pc.v = _val(k).pc
    err.v = 0


def pcsettype(pc, t, err):
    _val(pc).type = _val(t)
    _world().log('PCSetType', _val(t))
    err.v = 0


def pcsetfromoptions(pc, err):
    err.v = 0


def kspsolve(k, b, x, err):
    ksp, rhs, sol = _val(k), _val(b), _val(x)
    w = _world()
    w.barrier.wait()
    if _rt.rank == 0:
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        rowptr, col, val = ksp.mat.csr()
        A = sp.csr_matrix((val, col, rowptr), shape=(ksp.mat.N, ksp.mat.N))
        w.system = (rowptr, col, val, rhs.data.copy())
        sol.data[:] = spla.spsolve(A.tocsc(), rhs.data) if ksp.mat.N else 0.0
        ksp.reason, ksp.its, ksp.solves = 2, 0, ksp.solves + 1
        w.log('KSPSolve', (ksp.type, ksp.pc.type))
    w.barrier.wait()
    err.v = 0


def kspgetconvergedreason(k, reason, err):
    reason.v = _val(k).reason
    err.v = 0


def kspgetiterationnumber(k, its, err):
    its.v = _val(k).its
    err.v = 0


def kspdestroy(k, err):
    err.v = 0


def vecscattercreatetoall(v, ctx, vseq, err):
    src = _val(v)
    ctx.v = 'scatter'
    seq = MockVec()
    seq.N = src.N
    seq.data = np.zeros(src.N)
    vseq.v = seq
    err.v = 0


def vecscatterbegin(ctx, v, vseq, mode, direction, err):
    _world().barrier.wait()
    _val(vseq).data[:] = _val(v).data
    err.v = 0


def vecscatterend(ctx, v, vseq, mode, direction, err):
    err.v = 0


def vecscatterdestroy(ctx, err):
    err.v = 0


def vecrestorearray(v, a, off, err):
    err.v = 0


def vecgetarray_rewrite(gen, args, no):
    """`call VecGetArray(vec, xx_v, xx_i, err)` -- the F77 idiom xx_v(xx_i + k): bind xx_v to the vector's storage, xx_i = 0."""
    from .fortran_to_py import mangle
    gen.emit(f'{mangle(args[1][1])} = {gen.ex(args[0])}.data')
    gen.emit(f'{mangle(args[2][1])} = 0')


# ---- METIS / VTK ------------------------------------------------------------------------------------------------------

def metis_setdefaultoptions(opts):
    pass


def _metis(ne, nn, eptr, eind, vwgt, vsize, nparts, tpwgts, options, objval, epart, npart):
    w = _world()
    if w.partition is None:
        raise RuntimeError("METIS was called but the harness supplied no partition")
    ep, npn = w.partition
    assert _val(nparts) == w.size
    epart[:] = ep
    npart[:] = npn
    objval.v = 0
    w.log('METIS', (int(_val(ne)), int(_val(nn)), int(_val(nparts))))
    w.metis_input = (np.array(eptr), np.array(eind))


metis_partmeshnodal = _metis


def metis_partmeshdual(ne, nn, eptr, eind, vwgt, vsize, ncommon, nparts, tpwgts, options, objval, epart, npart):
    _metis(ne, nn, eptr, eind, vwgt, vsize, nparts, tpwgts, options, objval, epart, npart)


def writeoutputvtk(*args):
    """(ndim, nElem, nNode, npElem, ndof, coords, conn, elem_procid, soln[, file name]); triaelasticityparallelimpl1.F:1070
    omits the file name."""
    _world().vtk = dict(ndim=_val(args[0]), nElem=_val(args[1]), nNode=_val(args[2]), npElem=_val(args[3]),
                        ndof=_val(args[4]), conn=np.array(args[6]), soln=np.array(args[8]),
                        file=str(_val(args[9])).strip() if len(args) > 9 else None)


def namespace():
    """every mock, keyed the way the translator mangles Fortran names."""
    from .fortran_to_py import mangle
    g = globals()
    out = {}
    for k, v in g.items():
        if k.startswith('_') or k in ('np', 'threading', 'time', 'Ref', 'World', 'annotations'):
            continue
        if k[0].isupper() and not k.isupper() and not k.startswith(('PETSC', 'MPI', 'MAT', 'VEC', 'KSP', 'PC',
                                                                     'INSERT', 'ADD', 'SCATTER')):
            continue                                  # classes
        out[mangle(k)] = v
    return out
