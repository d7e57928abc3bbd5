"""Host side of the four *parallelimpl1 drivers, above the C ABI.

Mirrors what the Fortran PROGRAMs do around the hot path (tetrapoissonparallelimpl1.F and siblings):
METIS partition (:457-467), node renumbering and DOF numbering (:500-677), ElemDofArray (:698-713),
solver initialise (:759-779), pattern pass (:791-802), setZero (:817), the value pass (:828-884, one batched
call here), the ForceBC add (tetraelasticityparallelimpl1.F:971-982), factoriseAndSolve (:900) and the
gather of the solution into node order (:911-943).  The numbering / partition arithmetic runs in the C++ host
functions of libpfemb200.so (csrc/host_driver.cu); this file only sequences the calls, like the PROGRAM body.
One Python process = one MPI rank of the reference = one GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import solver as S
from .mesh import Mesh

# material / time-integration constants hard-coded in the drivers (single-precision literals widened):
# tetrapoissonparallelimpl1.F:822-824, tetraelasticityparallelimpl1.F:895-899, triaelasticityparallelimpl1.F:907
_f = lambda v: float(np.float32(v))
DEFAULT_ELEMDATA = {
    S.POISSON_TRIA: [1.0, 1.0, 1.0],
    S.POISSON_TETRA: [1.0, 1.0, 1.0],
    S.ELASTICITY_TRIA: [_f(240.565), _f(0.3), 1.0, 0.0, 0.0],               # thick = 1, b = 0 (documented intent)
    S.ELASTICITY_TETRA: [_f(240.565), _f(0.3), 1.0, _f(0.1), 0.0, 0.0],
}
DEFAULT_TIMEDATA = [0.0, 1.0, 0.0]      # timeData(2) = af = 1, timeData(3) = 0


@dataclass
class Numbering:
    kind: int
    nparts: int
    size_global: int
    node_map_get_old: np.ndarray     # int32 [nNode], 1-based
    node_map_get_new: np.ndarray
    NodeDofArrayNew: np.ndarray      # int32 [ndof, nNode], 1-based dof id, 0 = Dirichlet
    solnApplied: np.ndarray          # float64 [nNode*ndof], NEW numbering
    part_info: np.ndarray            # int32 [nparts, 5]: node_start, node_end, row_start, row_end (1-based), size_local
    conn_new: np.ndarray             # int32 [npElem, nElem] NEW 1-based node ids
    elemDof: np.ndarray              # int32 [nsize, nElem] 0-based dof ids, -1 = Dirichlet

    def row_range(self, rank: int):
        """0-based half-open owned row block of a rank (contiguous, in rank order)."""
        sl = self.part_info[:, 4]
        lo = int(sl[:rank].sum())
        return lo, lo + int(sl[rank])


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def partition(mesh: Mesh, kind: int, nparts: int):
    """METIS_PartMeshNodal for tetrahedra, METIS_PartMeshDual(ncommon=2) for triangles."""
    lib = S.load_library()
    conn = np.ascontiguousarray(mesh.conn, np.int32)
    epart = np.zeros(mesh.nElem, np.int32)
    npart = np.zeros(mesh.nNode, np.int32)
    objval = C.c_longlong()
    dual = 1 if mesh.npElem == 3 else 0
    S._chk(lib.pfem_host_partition_mesh(mesh.nElem, mesh.nNode, mesh.npElem, _ip(conn), nparts, dual, 2, _ip(epart),
                                        _ip(npart), C.byref(objval)))
    return epart, npart


def number(mesh: Mesh, kind: int, nparts: int = 1, node_proc_id=None, on_gpu: bool = False, device: int = 0) -> Numbering:
    """The drivers' numbering block (tetrapoissonparallelimpl1.F:357-734).  on_gpu=True runs it as sorts / scans / gathers on
    the GPU (csrc/gpu_setup.cu); the outputs are bit-identical (tests/test_gpu_setup.py)."""
    if on_gpu:
        return _number_gpu(mesh, kind, nparts, node_proc_id, device)
    lib = S.load_library()
    npe, ndof, ndim = S.KIND_DIMS[kind]
    nNode, nElem = mesh.nNode, mesh.nElem
    old = np.zeros(nNode, np.int32)
    new = np.zeros(nNode, np.int32)
    nda = np.zeros((ndof, nNode), np.int32)
    applied = np.zeros(nNode * ndof)
    info = np.zeros((max(nparts, 1), 5), np.int32)
    dn, dd, dv = (np.ascontiguousarray(mesh.dbc_node, np.int32), np.ascontiguousarray(mesh.dbc_dof, np.int32),
                  np.ascontiguousarray(mesh.dbc_val, np.float64))
    npid = np.ascontiguousarray(node_proc_id, np.int32) if (node_proc_id is not None and nparts > 1) else None
    sg = lib.pfem_host_number_dofs(nNode, ndof, dn.size, _ip(dn), _ip(dd), _dp(dv), nparts, _ip(npid), _ip(old), _ip(new),
                                   _ip(nda), _dp(applied), _ip(info))
    if sg < 0:
        raise S.PfemError(-sg, lib.pfem_last_error().decode())
    conn_new = np.ascontiguousarray(mesh.conn, np.int32).copy()
    lib.pfem_host_renumber_conn(C.c_longlong(conn_new.size), _ip(conn_new), _ip(new))
    edof = np.zeros((npe * ndof, nElem), np.int32)
    lib.pfem_host_elem_dof_array(nElem, npe, ndof, nNode, _ip(conn_new), _ip(nda), _ip(edof))
    return Numbering(kind, max(nparts, 1), sg, old, new, nda, applied, info, conn_new, edof)


def _number_gpu(mesh: Mesh, kind: int, nparts: int, node_proc_id, device: int) -> Numbering:
    lib = S.load_library()
    npe, ndof, ndim = S.KIND_DIMS[kind]
    nNode, nElem = mesh.nNode, mesh.nElem
    old = np.zeros(nNode, np.int32)
    new = np.zeros(nNode, np.int32)
    nda = np.zeros((ndof, nNode), np.int32)
    applied = np.zeros(nNode * ndof)
    info = np.zeros((max(nparts, 1), 5), np.int32)
    dn, dd, dv = (np.ascontiguousarray(mesh.dbc_node, np.int32), np.ascontiguousarray(mesh.dbc_dof, np.int32),
                  np.ascontiguousarray(mesh.dbc_val, np.float64))
    npid = np.ascontiguousarray(node_proc_id, np.int32) if (node_proc_id is not None and nparts > 1) else None
    sg = lib.pfem_gpu_number_dofs(device, nNode, ndof, dn.size, _ip(dn), _ip(dd), _dp(dv), nparts, _ip(npid), _ip(old), _ip(new),
                                  _ip(nda), _dp(applied), _ip(info))
    if sg < 0:
        raise S.PfemError(-sg, lib.pfem_last_error().decode())
    conn_new = np.ascontiguousarray(mesh.conn, np.int32).copy()
    rc = lib.pfem_gpu_renumber_conn(device, C.c_longlong(conn_new.size), _ip(conn_new), nNode, _ip(new))
    if rc < 0:
        raise S.PfemError(-rc, lib.pfem_last_error().decode())
    edof = np.zeros((npe * ndof, nElem), np.int32)
    rc = lib.pfem_gpu_elem_dof_array(device, nElem, npe, ndof, nNode, _ip(conn_new), _ip(nda), sg, _ip(edof), None, 0, 0, None)
    if rc < 0:
        raise S.PfemError(-rc, lib.pfem_last_error().decode())
    return Numbering(kind, max(nparts, 1), sg, old, new, nda, applied, info, conn_new, edof)


def gpu_local_elements_and_assy(num: Numbering, rank: int, device: int = 0):
    """(owned + overlap elements of the rank, assyForSoln) formed on the GPU."""
    lib = S.load_library()
    npe, ndof, ndim = S.KIND_DIMS[num.kind]
    lo, hi = num.row_range(rank)
    nsize, nElem = num.elemDof.shape
    nNode = num.NodeDofArrayNew.shape[1]
    edof = np.zeros((nsize, nElem), np.int32)
    assy = np.zeros(num.size_global, np.int32)
    lst = np.zeros(nElem, np.int32)
    n = lib.pfem_gpu_elem_dof_array(device, nElem, npe, ndof, nNode, _ip(np.ascontiguousarray(num.conn_new, np.int32)),
                                    _ip(np.ascontiguousarray(num.NodeDofArrayNew, np.int32)), num.size_global, _ip(edof), _ip(assy),
                                    lo, hi, _ip(lst))
    if n < 0:
        raise S.PfemError(-n, lib.pfem_last_error().decode())
    return lst[:n].copy(), assy, edof


def local_elements(num: Numbering, rank: int) -> np.ndarray:
    """Owned + overlap elements of a rank: every element with a dof in its row block, ascending id."""
    lib = S.load_library()
    lo, hi = num.row_range(rank)
    nsize, nElem = num.elemDof.shape
    n = lib.pfem_host_select_elements(nElem, nsize, _ip(num.elemDof), lo, hi, None)
    lst = np.zeros(n, np.int32)
    lib.pfem_host_select_elements(nElem, nsize, _ip(num.elemDof), lo, hi, _ip(lst))
    return lst


def force_bc_rows(mesh: Mesh, num: Numbering, ndof: int, fix: bool = False):
    """(row, value) pairs of the ForceBC add (tetraelasticityparallelimpl1.F:971-982).

    Default = the reference's row formula ``(newnode-1)*ndof + dof - 1`` (node based: it ignores the
    eliminated DOFs, and its 0-based row is range-tested against the 1-based row_start/row_end, so row 0
    is never added).  fix=True uses the NodeDofArrayNew index (documented-intent switch).
    """
    rows, vals = [], []
    for n_old, d, v in zip(mesh.fbc_node, mesh.fbc_dof, mesh.fbc_val):
        n1 = int(num.node_map_get_new[n_old - 1])
        if fix:
            dof = int(num.NodeDofArrayNew[d - 1, n1 - 1])
            if dof >= 1:
                rows.append(dof - 1)
                vals.append(float(v))
        else:
            row = (n1 - 1) * ndof + int(d) - 1
            if 1 <= row <= num.size_global and row < num.size_global:
                rows.append(row)
                vals.append(float(v))
    return rows, vals


def run_rank(solver: S.SolverB200, mesh: Mesh, num: Numbering, rank: int = 0, elemData=None, timeData=None,
             rtol: float = 1e-5, max_it: int = 10000, pc_type: int = S.PC_JACOBI, apply_force_bc: bool = True,
             fix_forcebc: bool = False, restrict_elements: bool = True, do_solve: bool = True):
    """The PROGRAM body from ``solverpetsc%initialise`` to ``factoriseAndSolve`` for one rank."""
    kind = num.kind
    npe, ndof, ndim = S.KIND_DIMS[kind]
    elemData = DEFAULT_ELEMDATA[kind] if elemData is None else elemData
    timeData = DEFAULT_TIMEDATA if timeData is None else timeData
    lo, hi = num.row_range(rank)
    size_local = hi - lo
    n1, n2 = (50, 25) if size_local >= 50 else (size_local, size_local)          # :759-773
    solver.initialise(size_local, num.size_global, np.full(max(size_local, 1), n1, np.int32),
                      np.full(max(size_local, 1), n2, np.int32))
    solver.set_options(rtol=rtol, max_it=max_it, pc_type=pc_type)
    if num.nparts > 1 and restrict_elements:
        lst = local_elements(num, rank)
        conn = np.ascontiguousarray(num.conn_new[:, lst])
        edof = np.ascontiguousarray(num.elemDof[:, lst])
    else:
        conn, edof = num.conn_new, num.elemDof
    old = num.node_map_get_old if num.nparts > 1 else None
    solver.set_mesh(kind, conn, mesh.coords, old)
    solver.set_pattern(edof)                      # pattern pass
    solver.setZero()
    solver.set_applied(num.solnApplied)
    solver.assemble(elemData, timeData)           # value pass
    if apply_force_bc and mesh.fbc_node.size:
        rows, vals = force_bc_rows(mesh, num, ndof, fix_forcebc)
        for r, v in zip(rows, vals):
            # the reference range-tests the row before VecSetValue (tetraelasticityparallelimpl1.F:977); with its stash every
            # admissible row is added exactly once overall: here by its owner (a row handed to a non-owner would be stashed and
            # shipped to the owner at the next solve, like PETSc does -- and then count twice)
            if lo <= r < hi:
                solver.add_value(r, v)
    if do_solve:
        solver.factoriseAndSolve()
    return solver.info()


def nodal_solution(num: Numbering, x_global: np.ndarray) -> np.ndarray:
    """solnVTK of the drivers (tetrapoissonparallelimpl1.F:911-943): [ndof, nNode] in OLD node numbering,
    applied values on Dirichlet dofs, the solution elsewhere."""
    ndof, nNode = num.NodeDofArrayNew.shape
    out = np.zeros((ndof, nNode))
    applied = num.solnApplied.reshape(nNode, ndof)
    old = num.node_map_get_old - 1
    for d in range(ndof):
        ids = num.NodeDofArrayNew[d]
        vals = np.where(ids > 0, x_global[np.maximum(ids, 1) - 1], applied[:, d])
        out[d, old] = vals
    return out
