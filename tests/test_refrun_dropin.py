"""The drop-in, executed on the CPU: the reference's own tetrapoissonparallelimpl1.F with the INTEGRATION.md diff applied
(oracle/refrun/dropin.py) runs to completion against a stand-in for `SolverB200` built on the oracle.  This checks the
PLUMBING of the diff -- what the edited program passes across the boundary and what it does with the answer; the same
program runs against the real libpfemb200.so in tests/test_gpu_zzzz_reference_vectors.py."""
import gzip
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from oracle.refrun import run_reference as R

pytestmark = [pytest.mark.skipif(not R.available(), reason="the reference tree exists only in the build container"),
              pytest.mark.filterwarnings("ignore::RuntimeWarning")]

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class OracleBackedSolver:
    """the python-level interface of pfemfort_b200.solver.SolverB200, answered by the CPU oracle (tests only)."""

    def __init__(self, device, rank, nranks):
        assert (device, rank, nranks) == (0, 0, 1)
        self.calls = ["create"]

    def initialise(self, size_local, size_global, diag_nnz, offdiag_nnz):
        assert size_local == size_global and diag_nnz.size == size_local and set(diag_nnz) == {50} and set(offdiag_nnz) == {25}
        self.N = size_global
        self.calls.append("initialise")

    def set_mesh(self, kind, conn, coords, old):
        self.kind, self.conn, self.coords, self.old = kind, conn, coords, old
        self.calls.append("set_mesh")

    def set_pattern(self, edof):
        self.edof = edof
        self.rp, self.col = O.pattern(edof, self.N)
        self.calls.append("set_pattern")

    def setZero(self):
        self.val, self.rhs = np.zeros(self.col.size), np.zeros(self.N)
        self.calls.append("setZero")

    def set_applied(self, applied):
        self.applied = applied
        self.calls.append("set_applied")

    def assemble(self, elemData, timeData):
        self.val, self.rhs, nbad = O.assemble(self.kind, self.conn, self.coords, self.old, self.edof, self.applied, elemData[:6],
                                              [0.0] + list(timeData[1:4]), self.rp, self.col)
        assert nbad == 0
        self.calls.append("assemble")

    def get_csr(self):
        return self.rp, self.col, self.val

    def get_rhs(self):
        return self.rhs

    def factoriseAndSolve(self):
        self.x, self.its, self.reason, _ = O.cg_bjacobi_ilu0(self.rp, self.col, self.val, self.rhs)     # the reference's defaults
        self.calls.append("factoriseAndSolve")

    def info(self):
        return dict(its=self.its, reason=self.reason)

    def get_solution(self):
        return self.x

    def free(self):
        self.calls.append("free")


def stage(tmp_path, prefix, kinds=("nodes", "elems", "DirichBC")):
    for kind in kinds:
        with gzip.open(os.path.join(GOLDEN, "input", f"{prefix}-{kind}.dat.gz")) as g, open(tmp_path / f"{prefix}-{kind}.dat", "wb") as o:
            o.write(g.read())
    return [f"{prefix}-{kind}.dat" for kind in kinds]


def test_edited_reference_program_runs_against_the_solver_interface(tmp_path):
    from oracle.refrun import dropin
    argv = stage(tmp_path, "tet10")
    bridge, rt = dropin.run("tetrapoissonparallelimpl1.F", argv, OracleBackedSolver, cwd=str(tmp_path))
    g = np.load(os.path.join(GOLDEN, "ref_driver_tet10_p1.npz"))
    # the order in which the edited PROGRAM drives the interface = the order of INTEGRATION.md
    assert bridge.h.calls == ["create", "initialise", "set_mesh", "set_pattern", "setZero", "set_applied", "assemble",
                              "factoriseAndSolve", "free"]
    # what crossed the boundary is what the unedited PROGRAM hands to PETSc
    c = bridge.captured
    assert np.array_equal(c["rowptr"], g["rowptr"]) and np.array_equal(c["col"], g["col"])
    assert np.array_equal(c["val"], g["val"]) and np.array_equal(c["rhs"], g["rhs"])
    assert np.array_equal(bridge.h.edof.T, g["ElemDofArray"]) and np.array_equal(bridge.h.applied, g["solnApplied"])
    assert c["info"]["reason"] > 0
    # and the reference's own temp.dat loop wrote the solution it got back: same index columns, values to the KSP tolerance
    rec = rt.written["temp.dat"]
    assert np.array_equal(np.array([[r[0], r[1]] for r in rec]), g["temp_dat_index"])
    x = np.array([r[2] for r in rec])
    assert np.array_equal(x, bridge.h.x)
    assert np.abs(x - g["temp_dat_value"]).max() <= 1e-4 * np.abs(g["temp_dat_value"]).max()


def test_the_edit_anchors_guard_the_line_numbers():
    from oracle.refrun import dropin
    text = "\n".join(l for l in dropin.patched_source("tetrapoissonparallelimpl1.F").split("\n") if not l.lstrip().startswith("!"))
    for gone in ("Module_SolverPetsc", "MatSetValues", "VecSetValues", "VecScatterCreateToAll", "VecGetArray", "xx_v(xx_i"):
        assert gone not in text, gone
    for kept in ("call solverpetsc%initialise(size_local, size_global,", "call solverpetsc%setZero()",
                 "call solverpetsc%factoriseAndSolve()", "call solverpetsc%free()", "write(1,*) ii, ind, fact"):
        assert kept in text, kept


# ---- the same PROGRAM through the real Fortran module (include/pfem_b200.f90), against a C test double of the ABI ----------

@pytest.fixture(scope="module")
def fake_lib(tmp_path_factory):
    import ctypes
    import subprocess
    d = tmp_path_factory.mktemp("fake_abi")
    so = str(d / "libfakepfem.so")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fake_abi", "fake_pfem.c")
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-Wall", "-o", so, src], check=True)
    lib = ctypes.CDLL(so)
    lib.fake_last.restype = ctypes.c_void_p
    lib.fake_scalar.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.fake_ints.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.fake_ints.restype = ctypes.POINTER(ctypes.c_int)
    lib.fake_doubles.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.fake_doubles.restype = ctypes.POINTER(ctypes.c_double)
    return lib


def test_edited_program_through_the_fortran_module_marshals_the_abi_exactly(tmp_path, fake_lib):
    """include/pfem_b200.f90 itself is executed (translated by oracle/refrun), its BIND(C) interfaces bound to a C test double
    with the prototypes of include/pfem_b200.h: every scalar and array the reference's edited PROGRAM passes arrives as the C
    side expects it (int32, SoA), in the order of INTEGRATION.md, and the PROGRAM's own temp.dat loop writes what
    pfem_solver_get_solution returned."""
    from oracle.refrun import dropin
    argv = stage(tmp_path, "tet10")
    seen = {}
    rt = dropin.run_through_module("tetrapoissonparallelimpl1.F", argv, fake_lib, cwd=str(tmp_path),
                                   before_free=lambda h: seen.setdefault("handle", h))
    g = np.load(os.path.join(GOLDEN, "ref_driver_tet10_p1.npz"))
    f = fake_lib.fake_last()
    assert seen["handle"] == f
    sc = [fake_lib.fake_scalar(f, k) for k in range(11)]
    nElem, nNode, N = 6000, 1331, 729
    assert sc[:10] == [0, 0, 1, N, N, 1, nElem, nNode, 4, nNode]       # device, rank, nranks, sizes, kind = PFEM_POISSON_TETRA ...
    ints = lambda k, n: np.ctypeslib.as_array(fake_lib.fake_ints(f, k), shape=(n,)).copy()          # noqa: E731
    dbls = lambda k, n: np.ctypeslib.as_array(fake_lib.fake_doubles(f, k), shape=(n,)).copy()       # noqa: E731
    assert list(ints(5, sc[10])) == [1, 2, 3, 4, 5, 6, 7, 8, 9, 10]    # create ... get_solution, free
    assert set(ints(0, N)) == {50} and set(ints(1, N)) == {25}
    from pfemfort_b200 import mesh as M
    m = M.read_mesh(os.path.join(GOLDEN, "input", "tet10"))
    assert np.array_equal(ints(2, 4 * nElem).reshape(4, nElem), m.conn)                    # SoA [npElem][nElem], one rank: new = old
    assert np.array_equal(dbls(0, 3 * nNode).reshape(3, nNode), m.coords)
    assert np.array_equal(ints(3, nNode), np.arange(1, nNode + 1))
    assert np.array_equal(ints(4, 4 * nElem).reshape(4, nElem), g["ElemDofArray"].T)
    assert np.array_equal(dbls(1, nNode), g["solnApplied"])
    assert list(dbls(2, 8)) == [1.0, 1.0, 1.0, 0, 0, 0, 0, 0] and list(dbls(3, 8)) == [0, 1.0, 0, 0, 0, 0, 0, 0]
    rec = rt.written["temp.dat"]
    assert np.array_equal(np.array([[r[0], r[1]] for r in rec]), g["temp_dat_index"])
    assert [r[2] for r in rec] == [1000.0 + i for i in range(N)]


def test_edited_elasticity_program_with_its_forcebc_loop(tmp_path, fake_lib):
    """tetraelasticityparallelimpl1.F with the same diff: ndof = 3, the driver's own material constants, and its ForceBC loop
    kept as is with VecSetValue replaced by pfem_solver_add_value -- rows and values as the unedited PROGRAM
    adds them (the reference's node-based row formula and 0- / 1-based range test included)."""
    from oracle.refrun import dropin
    from pfemfort_b200 import driver as D, mesh as M, solver as S
    argv = stage(tmp_path, "beam3Dtet6366", ("nodes", "elems", "DirichBC", "ForceBC"))
    # documented-intent input: local nodes 3 <-> 4 (the shipped file has negative Jacobians)
    lines = [l.split() for l in open(tmp_path / argv[1]) if l.strip()]
    with open(tmp_path / argv[1], "w") as f:
        for t in lines:
            f.write(f"{t[0]} {t[1]} {t[2]} {t[4]} {t[3]}\n")
    rt = dropin.run_through_module("tetraelasticityparallelimpl1.F", argv, fake_lib, cwd=str(tmp_path))
    g = np.load(os.path.join(GOLDEN, "ref_driver_beam3Dtet6366_p1.npz"))
    f = fake_lib.fake_last()
    sc = [fake_lib.fake_scalar(f, k) for k in range(12)]
    m = M.read_mesh(os.path.join(GOLDEN, "input", "beam3Dtet6366"), swap_34=True)
    N = g["rowptr"].size - 1
    assert sc[3:10] == [N, N, 3, m.nElem, m.nNode, 12, 3 * m.nNode]                 # kind = PFEM_ELASTICITY_TETRA, nsize = 12
    ints = lambda k, n: np.ctypeslib.as_array(fake_lib.fake_ints(f, k), shape=(n,)).copy()          # noqa: E731
    dbls = lambda k, n: np.ctypeslib.as_array(fake_lib.fake_doubles(f, k), shape=(n,)).copy()       # noqa: E731
    assert np.array_equal(ints(2, 4 * m.nElem).reshape(4, m.nElem), m.conn)
    assert np.array_equal(ints(4, 12 * m.nElem).reshape(12, m.nElem), g["ElemDofArray"].T)
    assert np.array_equal(dbls(1, 3 * m.nNode), g["solnApplied"])
    assert list(dbls(2, 8)[:6]) == D.DEFAULT_ELEMDATA[S.ELASTICITY_TETRA]           # 240.565, 0.3, 1.0, 0.1, 0, 0 as single literals
    # the ForceBC adds = what driver.force_bc_rows hands to add_value on one rank
    num = D.number(m, S.ELASTICITY_TETRA)
    rows, vals = D.force_bc_rows(m, num, 3)
    assert sc[11] == len(rows) > 0
    assert list(ints(6, sc[11])) == list(rows) and list(dbls(4, sc[11])) == list(vals)
