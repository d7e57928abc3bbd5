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


def stage(tmp_path, prefix):
    for kind in ("nodes", "elems", "DirichBC"):
        with gzip.open(os.path.join(GOLDEN, "input", f"{prefix}-{kind}.dat.gz")) as g, open(tmp_path / f"{prefix}-{kind}.dat", "wb") as o:
            o.write(g.read())
    return [f"{prefix}-{kind}.dat" for kind in ("nodes", "elems", "DirichBC")]


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
