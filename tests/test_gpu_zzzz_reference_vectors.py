"""The CUDA path (through the C ABI) against GOLDEN VECTORS PRODUCED BY RUNNING THE REFERENCE'S OWN SOURCE
(tests/golden/ref_*.npz; see tests/golden/make_reference_vectors.py and tests/test_reference_vectors.py, which holds the
oracle to the same files on the CPU).  Default arithmetic mode (reference order, no FMA): bit for bit.  Nothing here reads
/root/reference or the oracle: the committed files are the reference's outputs."""
import os

import numpy as np
import pytest

from pfemfort_b200 import driver as D, explicit as X, mesh as M, solver as S

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KINDS = [S.POISSON_TRIA, S.POISSON_TETRA, S.ELASTICITY_TRIA, S.ELASTICITY_TETRA]
CASES = {"tria20x20": (S.POISSON_TRIA, False), "tet10": (S.POISSON_TETRA, False),
         "cookmembranetria32": (S.ELASTICITY_TRIA, False), "beam3Dtet6366": (S.ELASTICITY_TETRA, True)}


@pytest.fixture(autouse=True)
def default_mode(monkeypatch):
    monkeypatch.delenv("PFEM_ASM", raising=False)


@pytest.fixture(scope="module")
def elements():
    return np.load(os.path.join(GOLDEN, "ref_elements.npz"))


@pytest.mark.parametrize("kind", KINDS)
def test_cuda_element_routines_equal_the_executed_reference(gpu, elements, kind):
    g = elements
    npe, ndof, ndim = S.KIND_DIMS[kind]
    xyz, ed, td, vc = g[f"ke{kind}_xyz"], g[f"ke{kind}_ed"], g[f"ke{kind}_td"], g[f"ke{kind}_valc"]
    n = xyz.shape[0]
    checked = 0
    for e in range(n):      # elemData / timeData differ per element in the golden set: one launch per element
        K, F, neg = S.element_ke_batch(kind, xyz[e, 0][:, None], xyz[e, 1][:, None],
                                       xyz[e, 2][:, None] if ndim == 3 else None, list(ed[e]), list(td[e]), vc[e][:, None])
        assert bool(neg[0]) == bool(g[f"ke{kind}_neg"][e])        # PFEM_ERR_NEG_JACOBIAN exactly where the reference STOPs
        if neg[0]:
            continue
        assert np.array_equal(K[:, :, 0], g[f"ke{kind}_K"][e]), (kind, e)
        assert np.array_equal(F[:, 0], g[f"ke{kind}_F"][e]), (kind, e)
        checked += 1
    assert checked > n // 3


@pytest.mark.parametrize("kind", [S.ELASTICITY_TRIA, S.ELASTICITY_TETRA])
def test_cuda_explicit_routines_equal_the_executed_reference(gpu, elements, kind):
    g = elements
    npe, ndof, ndim = S.KIND_DIMS[kind]
    xyz, ed, td, vc = g[f"ke{kind}_xyz"], g[f"ke{kind}_ed"], g[f"ke{kind}_td"], g[f"ke{kind}_valc"]
    for e in range(xyz.shape[0]):
        if g[f"ke{kind}_neg"][e]:
            continue
        z = xyz[e, 2] if ndim == 3 else None
        assert np.array_equal(X.residual_elasticity(kind, xyz[e, 0], xyz[e, 1], z, list(ed[e]), list(td[e]), vc[e]),
                              g[f"res{kind}_F"][e]), (kind, e)
        assert np.array_equal(X.mass_matrix(kind, xyz[e, 0], xyz[e, 1], z, list(ed[e])), g[f"mass{kind}_M"][e]), (kind, e)


@pytest.mark.parametrize("name", list(CASES))
def test_cuda_system_equals_what_the_executed_driver_hands_to_petsc(gpu, input_dir, name):
    """One rank: numbering on the GPU, pattern, value pass with lifting, ForceBC -- the matrix and right-hand side the
    reference's PROGRAM hands to KSPSolve, bit for bit; the solution against the reference run's (direct) solve."""
    kind, swap = CASES[name]
    g = np.load(os.path.join(GOLDEN, f"ref_driver_{name}_p1.npz"))
    m = M.read_mesh(os.path.join(input_dir, name), swap_34=swap)
    num = D.number(m, kind, on_gpu=True)
    assert np.array_equal(num.NodeDofArrayNew.T, g["NodeDofArrayNew"]) and np.array_equal(num.elemDof.T, g["ElemDofArray"])
    assert np.array_equal(num.solnApplied, g["solnApplied"])
    s = S.SolverB200(0)
    D.run_rank(s, m, num, rtol=1e-12, max_it=50000)
    rp, col, val = s.get_csr()
    rhs = s.get_rhs()
    assert np.array_equal(rp, g["rowptr"]) and np.array_equal(col, g["col"])
    assert s.assembly_mode()[0] in (0, 1)
    assert np.array_equal(val, g["val"]) and np.array_equal(rhs, g["rhs"])
    x = s.get_solution()
    ref = g["temp_dat_value"]                 # the mock KSPSolve is a direct solve; CG here stops at rtol 1e-12
    assert np.abs(x - ref).max() <= 1e-5 * np.abs(ref).max()
    s.free()


@pytest.mark.parametrize("name,p", [("tria20x20", 3), ("tet10", 2), ("tet10", 4), ("cookmembranetria32", 2)])
def test_cuda_numbering_equals_the_executed_p_rank_driver(gpu, input_dir, name, p):
    _numbering_against_the_executed_driver(input_dir, name, p)


def _numbering_against_the_executed_driver(input_dir, name, p):
    """P simulated ranks of the reference: the GPU numbering block (csrc/gpu_setup.cu) with the reference run's node
    partition, bit for bit -- renumbering, NodeDofArrayNew, ElemDofArray, applied values, per-rank node / row ranges,
    assyForSoln.  (The P-rank matrix blocks are held to the oracle in tests/test_gpu_multi.py and the oracle to these files
    in tests/test_reference_vectors.py.)"""
    kind, swap = CASES[name]
    g = np.load(os.path.join(GOLDEN, f"ref_driver_{name}_p{p}.npz"))
    m = M.read_mesh(os.path.join(input_dir, name), swap_34=swap)
    num = D.number(m, kind, p, g["node_proc_id"], on_gpu=True)
    assert np.array_equal(num.node_map_get_old, g["node_map_get_old"])
    assert np.array_equal(num.node_map_get_new, g["node_map_get_new"])
    assert np.array_equal(num.NodeDofArrayNew.T, g["NodeDofArrayNew"]) and np.array_equal(num.elemDof.T, g["ElemDofArray"])
    assert np.array_equal(num.solnApplied, g["solnApplied"])
    info = g["part_info"]
    assert np.array_equal(num.part_info[:, [0, 1, 4]], info[:, [0, 1, 4]])
    own = info[:, 4] > 0
    assert np.array_equal(num.part_info[own][:, [2, 3]], info[own][:, [2, 3]])
    lst, assy, edof = D.gpu_local_elements_and_assy(num, 0)
    assert np.array_equal(assy, g["assyForSoln"])


def test_cuda_gen_tetra_writes_the_reference_generators_files(gpu):
    """pfem_gpu_gen_tetra against the digests of the files the COMPILED reference generator writes
    (oracle/_ref/genTetranovtk; tests/test_reference_gentetra.py), and against a live run of that binary where it is present."""
    import test_reference_gentetra as T
    from oracle import ref_gentetra as G
    for name in sorted(T.GRIDS):
        T.check_against_golden(name, M.gen_tetra_gpu(*T.GRIDS[name]["grid"]))
    if G.available():
        grid = (-1, 1, 24, -1, 1, 24, -1, 1, 24)
        r, dev = G.run(*grid), M.gen_tetra_gpu(*grid)
        assert np.array_equal(dev.coords, r["coords"]) and np.array_equal(dev.conn, r["conn"])
        assert np.array_equal(dev.dbc_node, r["dbc_node"])


def test_cuda_explicit_time_loop_equals_the_executed_program(gpu, input_dir):
    """triaelasticityexplicit.F executed end to end (40 steps, cook membrane; tests/golden/ref_explicit_*.npz) against the
    fused one-launch-per-step GPU loop: lumped mass and state bit for bit."""
    g = np.load(os.path.join(GOLDEN, "ref_explicit_cookmembranetria32.npz"))
    kind = S.ELASTICITY_TRIA
    m = M.read_mesh(os.path.join(input_dir, "cookmembranetria32"))
    num = D.number(m, kind)
    ex = X.ExplicitB200(0)
    ex.set_mesh(kind, num.conn_new, m.coords)
    ex.set_free_dofs(X.free_slots(num))
    ex.lumped_mass(X.DRIVER_ELEMDATA_TRIA)
    n = int(g["steps"])
    ex.advance(n // 2, X.DRIVER_DT, X.DRIVER_ELEMDATA_TRIA, X.DRIVER_TIMEDATA)
    mid = ex.get_state()
    assert [mid["disp"][670], mid["disp"][671], mid["velo"][670], mid["velo"][671]] == list(g["solnoutput"][n // 2 - 1][1:])
    ex.advance(n - n // 2, X.DRIVER_DT, X.DRIVER_ELEMDATA_TRIA, X.DRIVER_TIMEDATA)
    st = ex.get_state()
    assert np.array_equal(st["mass"], g["globalM"])
    for key in ("disp", "dispPrev2", "velo", "acce"):
        assert np.array_equal(st[key], g[key]), key
    ex.free()


# ---- the last four (added late in round 2; run on the B200 in the round's last GPU call, profiles/r02_pytest_gpu_reference_vectors.log)

def test_cuda_numbering_equals_the_executed_p_rank_driver_beam(gpu, input_dir):
    _numbering_against_the_executed_driver(input_dir, "beam3Dtet6366", 2)


def test_cuda_petscsolver_procedures_equal_the_executed_wrapper(gpu):
    """TYPE PetscSolver's procedures as solverpetsc.F executes them (tests/golden/ref_solver_procedures.npz; the call sequence
    is test_reference_vectors.solver_procedure_calls) against the library's mirrors: same matrix, same right-hand side,
    PFEM_ERR_STATE where the reference STOPs."""
    import test_reference_vectors as T
    g = np.load(os.path.join(GOLDEN, "ref_solver_procedures.npz"))
    seq = T.solver_procedure_calls()
    s = S.SolverB200(0)
    s.initialise(729, 729)
    for bad in (s.factorise, s.solve):                       # out of order: the reference STOPs (:418, :444)
        with pytest.raises(S.PfemError) as ei:
            bad()
        assert ei.value.status == S.ERR_STATE
    # rows 0..5 carry the harness's three index sets; a 20 x 20 triangle grid (all dofs free) follows on rows 6.., so that the
    # pattern pass runs at the size of the bundled fixtures
    grid = M.gen_tria_poisson(20)
    special = np.array(seq["pattern"], np.int32).T           # [nsize, 3 elements], dofs = nodes - 1
    conn = np.concatenate([special + 1, grid.conn + 6], axis=1).astype(np.int32)
    coords = np.concatenate([np.array([[0.0, 1.0, 2.0, 0.0, 1.0, 2.0], [0.0, 0.0, 0.0, 1.0, 1.0, 1.5]]), grid.coords + 10.0], axis=1)
    N = coords.shape[1]
    s.free()
    s = S.SolverB200(0)
    s.initialise(N, N)
    s.set_mesh(S.POISSON_TRIA, conn, coords)
    s.set_pattern(conn - 1)
    s.setZero()
    for c in seq["calls"]:
        if c[0] == "mv":
            s.assembleMatrixAndVector(c[1], c[2], c[3], c[4])
        elif c[0] == "m":
            s.assembleMatrix(c[1], c[2], c[3])
        elif c[0] == "v":
            s.assembleVector(c[1], c[4])
        else:
            s.add_value(c[1], c[2])
    rp, col, val = s.get_csr()
    k = g["rowptr"][-1]
    assert np.array_equal(rp[:7], g["rowptr"]) and np.array_equal(col[:k], g["col"])
    assert np.array_equal(val[:k], g["val"]) and np.array_equal(s.get_rhs()[:6], g["rhs"])
    assert not val[k:].any() and not s.get_rhs()[6:].any()
    s.free()


def test_the_references_own_program_with_the_integration_diff_runs_on_the_library(gpu, tmp_path):
    """tetrapoissonparallelimpl1.F, read from the reference tree, with the INTEGRATION.md diff applied to its text, executed
    statement by statement (oracle/refrun/dropin.py) against the real ctypes binding of libpfemb200.so: everything the diff
    does not touch is the reference's own code.  The translated program is an oracle/_ref/ artefact made by build() where the
    reference tree is (it travels with the snapshot, like oracle/_ref/genTetranovtk)."""
    import gzip
    from oracle.refrun import dropin
    if not dropin.available():
        pytest.skip("oracle/_ref/dropin_tetrapoissonparallelimpl1.py was not built (no reference tree at build time)")
    argv = []
    for kind in ("nodes", "elems", "DirichBC"):
        with gzip.open(os.path.join(GOLDEN, "input", f"tet10-{kind}.dat.gz")) as g_, open(tmp_path / f"tet10-{kind}.dat", "wb") as o:
            o.write(g_.read())
        argv.append(f"tet10-{kind}.dat")
    bridge, rt = dropin.run("tetrapoissonparallelimpl1.F", argv, S.SolverB200, cwd=str(tmp_path))
    g = np.load(os.path.join(GOLDEN, "ref_driver_tet10_p1.npz"))
    c = bridge.captured
    assert np.array_equal(c["rowptr"], g["rowptr"]) and np.array_equal(c["col"], g["col"])
    assert np.array_equal(c["val"], g["val"]) and np.array_equal(c["rhs"], g["rhs"])
    assert c["info"]["reason"] > 0 and c["info"]["its"] > 0
    rec = rt.written["temp.dat"]
    assert np.array_equal(np.array([[r[0], r[1]] for r in rec]), g["temp_dat_index"])
    x = np.array([r[2] for r in rec])
    assert np.abs(x - g["temp_dat_value"]).max() <= 1e-4 * np.abs(g["temp_dat_value"]).max()


def test_the_references_own_program_runs_on_the_library_through_the_fortran_module(gpu, tmp_path):
    """As above, but Module_SolverB200 is include/pfem_b200.f90 ITSELF, translated and executed, its BIND(C) interfaces bound
    to the real libpfemb200.so (the marshalling is checked against a C test double on the CPU in tests/test_refrun_dropin.py)."""
    import ctypes
    import gzip
    from oracle.refrun import dropin
    if not dropin.module_available():
        pytest.skip("oracle/_ref/dropin_f90_tetrapoissonparallelimpl1.py was not built (no reference tree at build time)")
    argv = []
    for kind in ("nodes", "elems", "DirichBC"):
        with gzip.open(os.path.join(GOLDEN, "input", f"tet10-{kind}.dat.gz")) as g_, open(tmp_path / f"tet10-{kind}.dat", "wb") as o:
            o.write(g_.read())
        argv.append(f"tet10-{kind}.dat")
    lib = S.load_library()
    c = {}

    def capture(handle):                       # just before the PROGRAM's own `call solverpetsc%free()`
        s = object.__new__(S.SolverB200)
        s._lib, s._h, s.rank, s.nranks = lib, ctypes.c_void_p(handle), 0, 1
        rp, col, val = s.get_csr()
        c.update(rowptr=np.array(rp), col=np.array(col), val=np.array(val), rhs=np.array(s.get_rhs()), info=dict(s.info()))
        s._h = ctypes.c_void_p()               # the handle stays the PROGRAM's to free

    rt = dropin.run_through_module("tetrapoissonparallelimpl1.F", argv, lib, cwd=str(tmp_path), before_free=capture)
    g = np.load(os.path.join(GOLDEN, "ref_driver_tet10_p1.npz"))
    assert np.array_equal(c["rowptr"], g["rowptr"]) and np.array_equal(c["col"], g["col"])
    assert np.array_equal(c["val"], g["val"]) and np.array_equal(c["rhs"], g["rhs"])
    assert c["info"]["reason"] > 0 and c["info"]["its"] > 0
    rec = rt.written["temp.dat"]
    assert np.array_equal(np.array([[r[0], r[1]] for r in rec]), g["temp_dat_index"])
    x = np.array([r[2] for r in rec])
    assert np.abs(x - g["temp_dat_value"]).max() <= 1e-4 * np.abs(g["temp_dat_value"]).max()
