"""ctypes front-end of the C oracle (oracle/pfem_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
PARITY: pinned by golden vectors produced by executing the reference's own source (oracle/refrun ->
tests/golden/ref_*.npz, tests/test_reference_vectors.py); the Krylov iterates (PETSc's) are not.  See pfem_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "_build", "liborc.so")

POISSON_TRIA, POISSON_TETRA, ELASTICITY_TRIA, ELASTICITY_TETRA = 0, 1, 2, 3
KIND_DIMS = {0: (3, 1, 2), 1: (4, 1, 3), 2: (3, 2, 2), 3: (4, 3, 3)}   # npElem, ndof, ndim


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "pfem_oracle.c")
    if force or not os.path.exists(LIBPATH) or os.path.getmtime(LIBPATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B"], check=True, capture_output=True)
    return LIBPATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_pattern.restype = C.c_longlong
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _d(a):
    return _p(a, C.c_double)


def _i(a):
    return _p(a, C.c_int)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def element_ke(kind, x, y, z, elemData, timeData, valC=None):
    """One element; returns (K column-major [nsize,nsize] as K[i,j]=Klocal(i,j), F, jac_negative)."""
    npe, ndof, ndim = KIND_DIMS[kind]
    ns = npe * ndof
    K = np.zeros(ns * ns)
    F = np.zeros(ns)
    x, y = _f64(x), _f64(y)
    ed, td = _f64(elemData), _f64(timeData)
    vc = _f64(valC) if valC is not None else np.zeros(ns)
    vd = np.zeros(ns)
    L = lib()
    if kind == POISSON_TRIA:
        rc = L.orc_poisson_tria_ke(_d(x), _d(y), _d(ed), _d(td), _d(vc), _d(vd), _d(K), _d(F))
    elif kind == POISSON_TETRA:
        z = _f64(z)
        rc = L.orc_poisson_tetra_ke(_d(x), _d(y), _d(z), _d(ed), _d(td), _d(vc), _d(vd), _d(K), _d(F))
    elif kind == ELASTICITY_TRIA:
        rc = L.orc_elasticity_tria_ke(_d(x), _d(y), _d(ed), _d(td), _d(vc), _d(vd), _d(K), _d(F))
    else:
        z = _f64(z)
        rc = L.orc_elasticity_tetra_ke(_d(x), _d(y), _d(z), _d(ed), _d(td), _d(vc), _d(vd), _d(K), _d(F))
    return K.reshape(ns, ns).T.copy(), F, rc      # K[i, j] = Klocal(i+1, j+1)


def tria_ke_closed_form(x, y):
    K = np.zeros(9)
    x, y = _f64(x), _f64(y)
    lib().orc_poisson_tria_ke_closed_form(_d(x), _d(y), _d(K))
    return K.reshape(3, 3).T.copy()


def number_dofs(nNode, ndof, dbc_node, dbc_dof, dbc_val, nparts=1, node_proc_id=None):
    dbc_node, dbc_dof, dbc_val = _i32(dbc_node), _i32(dbc_dof), _f64(dbc_val)
    old = np.zeros(nNode, np.int32)
    new = np.zeros(nNode, np.int32)
    nda = np.zeros((ndof, nNode), np.int32)
    applied = np.zeros(nNode * ndof)
    ns, ne, rs, re, sl = (np.zeros(nparts, np.int32) for _ in range(5))
    npid = _i32(node_proc_id) if node_proc_id is not None else None
    sg = lib().orc_number_dofs(nNode, ndof, dbc_node.size, _i(dbc_node), _i(dbc_dof), _d(dbc_val), nparts, _i(npid),
                               _i(old), _i(new), _i(nda), _d(applied), _i(ns), _i(ne), _i(rs), _i(re), _i(sl))
    if sg < 0:
        raise RuntimeError("orc_number_dofs: the reference would STOP")
    return dict(size_global=sg, node_map_get_old=old, node_map_get_new=new, NodeDofArrayNew=nda, solnApplied=applied,
                node_start=ns, node_end=ne, row_start=rs, row_end=re, size_local=sl)


def elem_dof_array(conn_new, NodeDofArrayNew):
    conn_new = _i32(conn_new)
    npe, nElem = conn_new.shape
    ndof, nNode = NodeDofArrayNew.shape
    edof = np.zeros((npe * ndof, nElem), np.int32)
    lib().orc_elem_dof_array(nElem, npe, ndof, nNode, _i(conn_new), _i(_i32(NodeDofArrayNew)), _i(edof))
    return edof


def pattern(edof, N):
    edof = _i32(edof)
    nsize, nElem = edof.shape
    rowptr = np.zeros(N + 1, np.int32)
    nnz = lib().orc_pattern(nElem, nsize, _i(edof), N, _i(rowptr), None)
    col = np.zeros(max(nnz, 1), np.int32)
    lib().orc_pattern(nElem, nsize, _i(edof), N, _i(rowptr), _i(col))
    return rowptr, col[:nnz]


def assemble(kind, conn_new, coords, node_map_get_old, edof, solnApplied, elemData, timeData, rowptr, col,
             elem_mask=None, row_lo=0, row_hi=None, threads=1, val=None, rhs=None):
    conn_new, edof = _i32(conn_new), _i32(edof)
    coords = _f64(coords)
    nElem = conn_new.shape[1]
    nNode = coords.shape[1]
    N = rowptr.size - 1
    if row_hi is None:
        row_hi = N
    if val is None:
        val = np.zeros(col.size)
    if rhs is None:
        rhs = np.zeros(N)
    m = np.ascontiguousarray(elem_mask, dtype=np.uint8) if elem_mask is not None else None
    old = _i32(node_map_get_old) if node_map_get_old is not None else None
    ed, td, sa = _f64(elemData), _f64(timeData), _f64(solnApplied)
    nbad = lib().orc_assemble(kind, nElem, _i(conn_new), nNode, _d(coords), _i(old), _i(edof), _d(sa), _d(ed), _d(td),
                              _p(m, C.c_ubyte), row_lo, row_hi, _i(rowptr), _i(col), _d(val), _d(rhs), threads)
    return val, rhs, nbad


def add_force_bc(rhs, fbc_node, fbc_dof, fbc_val, ndof, node_map_get_new, NodeDofArrayNew, size_global, fix=False):
    fbc_node, fbc_dof, fbc_val = _i32(fbc_node), _i32(fbc_dof), _f64(fbc_val)
    nNode = NodeDofArrayNew.shape[1]
    lib().orc_add_force_bc(fbc_node.size, _i(fbc_node), _i(fbc_dof), _d(fbc_val), ndof, nNode, _i(_i32(node_map_get_new)),
                           _i(_i32(NodeDofArrayNew)), 1, size_global, size_global, 1 if fix else 0, _d(rhs))
    return rhs


def cg_jacobi(rowptr, col, val, b, rtol=1e-5, abstol=1e-50, dtol=1e4, max_it=10000, threads=1, fixed_its=0):
    N = rowptr.size - 1
    x = np.zeros(N)
    its, reason = C.c_int(0), C.c_int(0)
    rnorm = C.c_double(0)
    lib().orc_cg_jacobi(N, _i(_i32(rowptr)), _i(_i32(col)), _d(_f64(val)), _d(_f64(b)), _d(x), C.c_double(rtol),
                        C.c_double(abstol), C.c_double(dtol), max_it, threads, fixed_its, C.byref(its), C.byref(reason),
                        C.byref(rnorm))
    return x, its.value, reason.value, rnorm.value


def cg_bjacobi_ilu0(rowptr, col, val, b, block_start=None, rtol=1e-5, abstol=1e-50, dtol=1e4, max_it=10000):
    """KSPCG + PCBJACOBI/ILU(0): the reference's default (solverpetsc.F:187,206).  block_start: row ranges of the ranks
    (one block per rank), default one block."""
    N = len(rowptr) - 1
    bs = _i32(np.array([0, N] if block_start is None else block_start))
    x = np.zeros(N)
    its, reason, rn = C.c_int(0), C.c_int(0), C.c_double(0)
    lib().orc_cg_bjacobi_ilu0(N, _i(rowptr), _i(col), _d(val), _d(b), _d(x), len(bs) - 1, _i(bs), C.c_double(rtol),
                              C.c_double(abstol), C.c_double(dtol), max_it, C.byref(its), C.byref(reason),
                              C.byref(rn))
    return x, its.value, reason.value, rn.value


def ilu0_factor(rowptr, col, val, block_start=None):
    N = len(rowptr) - 1
    bs = _i32(np.array([0, N] if block_start is None else block_start))
    fval = np.zeros(len(col))
    invd = np.zeros(N)
    rc = lib().orc_ilu0_factor(N, _i(rowptr), _i(col), _d(val), len(bs) - 1, _i(bs), _d(fval), _d(invd))
    return fval, invd, rc


def ilu0_solve(rowptr, col, fval, invdiag, r, block_start=None):
    N = len(rowptr) - 1
    bs = _i32(np.array([0, N] if block_start is None else block_start))
    z = np.zeros(N)
    lib().orc_ilu0_solve(N, _i(rowptr), _i(col), _d(fval), _d(invdiag), len(bs) - 1, _i(bs), _d(np.ascontiguousarray(r, np.float64)), _d(z))
    return z


# ---- explicit dynamics ----

def residual_elasticity(kind, x, y, z, elemData, timeData, dispC, veloC=None):
    """ResidualElasticityLinearTria / ...Tetra for one element: (Flocal, jac_negative)."""
    npe, ndof, ndim = KIND_DIMS[kind]
    F = np.zeros(npe * ndof)
    x, y = _f64(x), _f64(y)
    ed, td, dc = _f64(elemData), _f64(timeData), _f64(dispC)
    vc = _f64(veloC) if veloC is not None else np.zeros(npe * ndof)
    if kind == ELASTICITY_TRIA:
        rc = lib().orc_residual_elasticity_tria(_d(x), _d(y), _d(ed), _d(td), _d(dc), _d(vc), _d(F))
    else:
        zz = _f64(z)
        rc = lib().orc_residual_elasticity_tet(_d(x), _d(y), _d(zz), _d(ed), _d(td), _d(dc), _d(vc), _d(F))
    return F, rc


def mass_matrix(kind, x, y, z, elemData):
    """MassMatrixLinearTria / ...Tetra for one element: (Mlocal, jac_negative)."""
    npe, ndof, ndim = KIND_DIMS[kind]
    Ml = np.zeros(npe * ndof)
    x, y, ed = _f64(x), _f64(y), _f64(elemData)
    if kind == ELASTICITY_TRIA:
        rc = lib().orc_mass_matrix_tria(_d(x), _d(y), _d(ed), _d(Ml))
    else:
        zz = _f64(z)
        rc = lib().orc_mass_matrix_tet(_d(x), _d(y), _d(zz), _d(ed), _d(Ml))
    return Ml, rc


def explicit_lumped_mass(kind, conn, coords, elemData):
    conn, coords = _i32(conn), _f64(coords)
    npe, ndof, ndim = KIND_DIMS[kind]
    M = np.zeros(coords.shape[1] * ndof)
    nbad = lib().orc_explicit_lumped_mass(kind, conn.shape[1], _i(conn), coords.shape[1], _d(coords), _d(_f64(elemData)), _d(M))
    return M, nbad


def explicit_advance(kind, conn, coords, free_slots, elemData, timeData, dt, nsteps, M, state=None):
    """nsteps central-difference steps (triaelasticityexplicit.F:972-1121).  state = dict(disp, dispPrev, dispPrev2, velo,
    acce) is advanced in place (created zeroed when None) and returned."""
    conn, coords, fs = _i32(conn), _f64(coords), _i32(free_slots)
    npe, ndof, ndim = KIND_DIMS[kind]
    nd = coords.shape[1] * ndof
    if state is None:
        state = {k: np.zeros(nd) for k in ("disp", "dispPrev", "dispPrev2", "velo", "acce")}
    nbad = lib().orc_explicit_advance(kind, conn.shape[1], _i(conn), coords.shape[1], _d(coords), fs.size, _i(fs), _d(_f64(elemData)),
                                      _d(_f64(timeData)), C.c_double(dt), nsteps, _d(M), _d(state["disp"]), _d(state["dispPrev"]),
                                      _d(state["dispPrev2"]), _d(state["velo"]), _d(state["acce"]))
    assert nbad == 0
    return state


def num_threads():
    return lib().orc_num_threads()
