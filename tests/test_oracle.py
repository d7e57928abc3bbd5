"""CPU tests of the oracle (the checker): known answers of the reference's own fixtures, the independent
closed-form Ke, and the identities of SURVEY.md section 8(c).  (The golden vectors produced by executing the reference's own
source are in tests/test_reference_vectors.py; the pins here are independent of them.)"""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from pfemfort_b200 import driver as D, mesh as M, solver as S

ED = D.DEFAULT_ELEMDATA
TD = D.DEFAULT_TIMEDATA


def _system(m, kind, fbc=True, fix=False):
    npe, ndof, ndim = S.KIND_DIMS[kind]
    o = O.number_dofs(m.nNode, ndof, m.dbc_node, m.dbc_dof, m.dbc_val)
    conn_new = o["node_map_get_new"][m.conn - 1]
    edof = O.elem_dof_array(conn_new, o["NodeDofArrayNew"])
    rp, col = O.pattern(edof, o["size_global"])
    val, rhs, nbad = O.assemble(kind, conn_new, m.coords, o["node_map_get_old"], edof, o["solnApplied"], ED[kind], TD, rp, col)
    if fbc and m.fbc_node.size:
        O.add_force_bc(rhs, m.fbc_node, m.fbc_dof, m.fbc_val, ndof, o["node_map_get_new"], o["NodeDofArrayNew"], o["size_global"], fix=fix)
    return o, edof, rp, col, val, rhs, nbad


def _diag_sum(rp, col, val):
    rows = np.repeat(np.arange(rp.size - 1), np.diff(rp))
    return val[col == rows].sum()


def test_single_precision_literals():
    # w = REAL(1.0/6.0) = 0.1666666716337204 (elementutilitiespoisson.F:142): unit tet volume Jac = 1
    x, y, z = [1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 0, 1.0]
    K, F, rc = O.element_ke(O.POISSON_TETRA, x, y, z, [1, 1, 1], TD)
    assert rc == 0
    assert F[0] == 0.25 * 0.1666666716337204 * -6.0
    assert K[0, 0] == 0.1666666716337204          # grad N1 = (1,0,0), dvol = w


def test_tria20x20_known_answers(input_dir):
    m = M.read_mesh(os.path.join(input_dir, "tria20x20"))
    o, edof, rp, col, val, rhs, nbad = _system(m, O.POISSON_TRIA)
    assert (o["size_global"], col.size, nbad) == (361, 2377, 0)
    assert np.linalg.norm(rhs) == pytest.approx(3.162277655753922, rel=1e-13)
    assert _diag_sum(rp, col, val) == pytest.approx(1444.0, rel=1e-13)
    for rtol, its_expected in ((1e-5, 26), (1e-10, 31)):
        x, its, reason, _ = O.cg_jacobi(rp, col, val, rhs, rtol=rtol)
        assert (its, reason) == (its_expected, 2)
    free = o["NodeDofArrayNew"][0] > 0
    exact = M.exact_poisson_tria(m.coords[0, free], m.coords[1, free])
    assert np.abs(x - exact).max() == pytest.approx(7.1146e-4, rel=1e-3)


def test_tet10_known_answers(input_dir):
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    o, edof, rp, col, val, rhs, nbad = _system(m, O.POISSON_TETRA)
    assert (o["size_global"], col.size, nbad) == (729, 9097, 0)
    assert np.linalg.norm(rhs) == pytest.approx(24.573880283096567, rel=1e-13)
    assert _diag_sum(rp, col, val) == pytest.approx(1506.6000449001795, rel=1e-13)   # needs w = 0.1666666716337204
    for rtol, its_expected in ((1e-5, 28), (1e-10, 51)):
        x, its, reason, _ = O.cg_jacobi(rp, col, val, rhs, rtol=rtol)
        assert (its, reason) == (its_expected, 2)
    free = o["NodeDofArrayNew"][0] > 0
    assert np.abs(x - (m.coords[:, free] ** 2).sum(0)).max() == pytest.approx(1.157e-7, rel=1e-2)


def test_beam3Dtet6366_documented_intent(input_dir):
    raw = M.read_mesh(os.path.join(input_dir, "beam3Dtet6366"))
    *_, nbad = _system(raw, O.ELASTICITY_TETRA)
    assert nbad == raw.nElem == 7776                      # as shipped: every Jacobian negative (the reference STOPs)
    m = M.read_mesh(os.path.join(input_dir, "beam3Dtet6366"), swap_34=True)
    o, edof, rp, col, val, rhs, nbad = _system(m, O.ELASTICITY_TETRA, fbc=False)
    assert (o["size_global"], col.size, nbad) == (5292, 200106, 0)
    assert _diag_sum(rp, col, val) == pytest.approx(722620.29174, rel=1e-10)
    assert np.linalg.norm(rhs) == pytest.approx(0.0151252780473, rel=1e-10)     # body force only
    o, edof, rp, col, val, rhs, _ = _system(m, O.ELASTICITY_TETRA, fbc=True)
    assert np.linalg.norm(rhs) == pytest.approx(1.73225048334, rel=1e-10)       # ForceBC via the reference's row formula
    x, its, reason, _ = O.cg_jacobi(rp, col, val, rhs, rtol=1e-10)
    assert reason == 2 and its == pytest.approx(661, abs=3)
    u = np.zeros((3, m.nNode))
    nda = o["NodeDofArrayNew"]
    for d in range(3):
        u[d, nda[d] > 0] = x[nda[d][nda[d] > 0] - 1]
    assert np.sqrt((u ** 2).sum(0)).max() == pytest.approx(0.7078, abs=2e-4)


def test_cookmembrane_documented_intent(input_dir):
    m = M.read_mesh(os.path.join(input_dir, "cookmembranetria32"))
    o, edof, rp, col, val, rhs, nbad = _system(m, O.ELASTICITY_TRIA)
    assert (o["size_global"], col.size, nbad) == (2112, 28536, 0)
    assert _diag_sum(rp, col, val) == pytest.approx(3372223.35481, rel=1e-10)
    assert np.linalg.norm(rhs) == pytest.approx(17.5390190005, rel=1e-10)
    x, its, reason, _ = O.cg_jacobi(rp, col, val, rhs, rtol=1e-10)
    assert reason == 2 and its == pytest.approx(676, abs=3)
    # zero right-hand side: PETSc converges at iteration 0 with x = 0 (never divides by p.w = 0)
    x0, its0, reason0, _ = O.cg_jacobi(rp, col, val, np.zeros_like(rhs))
    assert (its0, reason0) == (0, 3) and np.all(x0 == 0)


def test_closed_form_tria_ke_cross_check():
    rng = np.random.default_rng(5)
    for _ in range(50):
        x = np.array([0.0, 1.0, 0.0]) + 0.3 * rng.standard_normal(3)
        y = np.array([0.0, 0.0, 1.0]) + 0.3 * rng.standard_normal(3)
        K, F, rc = O.element_ke(O.POISSON_TRIA, x, y, None, [1, 1], TD)
        if rc:
            continue
        Kc = O.tria_ke_closed_form(x, y)                  # triapoissonserialimpl1.F:580-594
        assert np.allclose(K, Kc, rtol=1e-12, atol=1e-14)
        assert np.abs(K.sum(1)).max() < 1e-13 and np.all(F == 0)


def test_identities(input_dir):
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    vol = 0.0
    for e in range(m.nElem):
        n = m.conn[:, e] - 1
        K, F, rc = O.element_ke(O.POISSON_TETRA, m.coords[0, n], m.coords[1, n], m.coords[2, n], [1, 1, 1], TD)
        assert rc == 0 and np.abs(K.sum(1)).max() < 1e-12            # row sums of the Poisson Ke vanish
        vol += F.sum() / -6.0                                          # sum Fe = force * volume
    assert vol == pytest.approx(16.0 * 0.1666666716337204 * 6, rel=1e-12)   # domain volume 16 (w is the float 1/6)


def test_explicit_zeros_are_part_of_the_pattern():
    m = M.gen_tria_poisson(100)
    o, edof, rp, col, val, rhs, _ = _system(m, O.POISSON_TRIA)
    assert col.size == 67817                               # a value-pruning library would give 48609
    assert (val == 0).sum() > 15000
    for r in range(0, rp.size - 1, 97):
        c = col[rp[r]:rp[r + 1]]
        assert np.all(np.diff(c) > 0)                      # sorted unique columns


def test_generators_reproduce_the_bundled_fixtures(input_dir):
    t = M.read_mesh(os.path.join(input_dir, "tet10"))
    g = M.gen_tetra(-2, 2, 10, -1, 1, 10, -1, 1, 10)
    assert np.array_equal(g.coords, t.coords) and np.array_equal(g.conn, t.conn)
    assert np.array_equal(g.dbc_node, t.dbc_node) and np.array_equal(g.dbc_val, t.dbc_val)
    t = M.read_mesh(os.path.join(input_dir, "tria20x20"))
    g = M.gen_tria_poisson(20)
    assert np.array_equal(g.coords, t.coords) and np.array_equal(g.conn, t.conn)
    assert np.array_equal(g.dbc_node, t.dbc_node) and np.array_equal(g.dbc_val, t.dbc_val)
    b = M.read_mesh(os.path.join(input_dir, "beam3Dtet6366"), swap_34=True)
    g = M.gen_tetra(-0.5, 0.5, 6, 0.0, 6.0, 36, -0.5, 0.5, 6, dbc="clamp_y0", ndof=3)
    assert np.array_equal(g.conn, b.conn)                 # file ordering with 3<->4 swapped == genTetra ordering
    assert np.abs(g.coords - b.coords).max() < 1e-8
    assert np.array_equal(g.dbc_node, b.dbc_node) and np.array_equal(g.dbc_dof, b.dbc_dof)


@pytest.mark.skipif(not os.path.exists("/root/reference/input/tet100-DirichBC.dat.gz"), reason="reference tree not present")
def test_generators_against_the_large_reference_files():
    import gzip
    ref = np.loadtxt(gzip.open("/root/reference/input/tet100-DirichBC.dat.gz", "rt"))
    g = M.gen_tetra(-1, 1, 100, -1, 1, 100, -1, 1, 100)
    assert np.array_equal(g.dbc_node, ref[:, 0].astype(np.int32)) and np.array_equal(g.dbc_val, ref[:, 2])
    ref = np.loadtxt(gzip.open("/root/reference/input/tria1000x1000-DirichBC.dat.gz", "rt"))
    g = M.gen_tria_poisson(1000)
    assert np.array_equal(g.dbc_node, ref[:, 0].astype(np.int32)) and np.array_equal(g.dbc_val, ref[:, 2])
    nodes = np.loadtxt(gzip.open("/root/reference/input/tria1000x1000-nodes.dat.gz", "rt"))
    assert np.array_equal(g.coords, nodes[:, 1:].T)


def test_force_bc_rows_host_logic_matches_oracle(input_dir):
    """driver.force_bc_rows (host side of the ForceBC loop, tetraelasticityparallelimpl1.F:971-982) against the oracle,
    with the reference's row formula and with the corrected index."""
    for name, kind, swap in (("beam3Dtet6366", S.ELASTICITY_TETRA, True), ("cookmembranetria32", S.ELASTICITY_TRIA, False)):
        m = M.read_mesh(os.path.join(input_dir, name), swap_34=swap)
        npe, ndof, ndim = S.KIND_DIMS[kind]
        num = D.number(m, kind)
        for fix in (False, True):
            rhs = np.zeros(num.size_global)
            O.add_force_bc(rhs, m.fbc_node, m.fbc_dof, m.fbc_val, ndof, num.node_map_get_new, num.NodeDofArrayNew, num.size_global, fix=fix)
            mine = np.zeros(num.size_global)
            rows, vals = D.force_bc_rows(m, num, ndof, fix)
            for r, v in zip(rows, vals):
                mine[r] += v
            assert np.array_equal(rhs, mine)


def test_nodal_solution_scatter(input_dir):
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    _, npart = D.partition(m, S.POISSON_TETRA, 3)
    num = D.number(m, S.POISSON_TETRA, 3, npart)
    x = np.arange(1, num.size_global + 1, dtype=np.float64)
    u = D.nodal_solution(num, x)[0]
    dbc = dict(zip(m.dbc_node, m.dbc_val))
    for old in range(1, m.nNode + 1, 37):
        new = num.node_map_get_new[old - 1]
        dof = num.NodeDofArrayNew[0, new - 1]
        assert u[old - 1] == (x[dof - 1] if dof > 0 else dbc[old])


def test_oracle_cg_against_a_direct_solve(input_dir):
    """The oracle's Jacobi-CG (PETSc semantics) converges to the solution of a sparse direct solve."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    for name, kind, swap in (("tet10", S.POISSON_TETRA, False), ("beam3Dtet6366", S.ELASTICITY_TETRA, True)):
        m = M.read_mesh(os.path.join(input_dir, name), swap_34=swap)
        o, edof, rp, col, val, rhs, _ = _system(m, kind)
        A = sp.csr_matrix((val, col, rp), shape=(rp.size - 1, rp.size - 1))
        assert abs(A - A.T).max() <= 1e-12 * abs(A).max()          # symmetric to rounding (Klocal is read transposed)
        xd = spla.spsolve(A.tocsc(), rhs)
        x, its, reason, rnorm = O.cg_jacobi(rp, col, val, rhs, rtol=1e-12, max_it=20000)
        assert reason == 2
        assert np.abs(x - xd).max() <= 1e-8 * np.abs(xd).max()


# ---- CG + PCBJACOBI/ILU(0), the reference's default preconditioner (solverpetsc.F:187,206) -------------------------

def _assembled(name, kind, input_dir, swap=False):
    m = M.read_mesh(os.path.join(input_dir, name), swap_34=swap)
    num = D.number(m, kind)
    rp, col = O.pattern(num.elemDof, num.size_global)
    val, rhs, nbad = O.assemble(kind, num.conn_new, m.coords, None, num.elemDof, num.solnApplied, D.DEFAULT_ELEMDATA[kind],
                                D.DEFAULT_TIMEDATA, rp, col)
    assert nbad == 0
    if m.fbc_node.size:
        O.add_force_bc(rhs, m.fbc_node, m.fbc_dof, m.fbc_val, S.KIND_DIMS[kind][1], num.node_map_get_new, num.NodeDofArrayNew,
                       num.size_global)
    return num, rp, col, val, rhs


def test_ilu0_factor_equals_dense_pattern_restricted_elimination(input_dir):
    """orc_ilu0_factor against an independent dense IKJ elimination restricted to the pattern (tet10, 729 rows), for one
    block and for two blocks (entries outside a block are ignored: block Jacobi)."""
    num, rp, col, val, rhs = _assembled("tet10", S.POISSON_TETRA, input_dir)
    N = num.size_global
    for starts in ([0, N], [0, N // 3, N]):
        A = np.zeros((N, N))
        P = np.zeros((N, N), bool)
        for i in range(N):
            b = max(k for k in range(len(starts) - 1) if starts[k] <= i)
            for q in range(rp[i], rp[i + 1]):
                if starts[b] <= col[q] < starts[b + 1]:
                    A[i, col[q]] = val[q]
                    P[i, col[q]] = True
        F = A.copy()
        for i in range(N):
            for k in np.nonzero(P[i, :i])[0]:
                F[i, k] = F[i, k] * (1.0 / F[k, k])
                js = np.nonzero(P[i, k + 1:] & P[k, k + 1:])[0] + k + 1
                F[i, js] -= F[i, k] * F[k, js]
        fval, invd, rc = O.ilu0_factor(rp, col, val, starts)
        assert rc == 0
        G = np.zeros((N, N))
        for i in range(N):
            G[i, col[rp[i]:rp[i + 1]]] = fval[rp[i]:rp[i + 1]]
        assert np.array_equal(G * P, F * P) and np.array_equal(invd, 1.0 / np.diag(F))
        # M z = r with M = L U: the two substitutions invert the factor exactly (to rounding)
        r = np.linspace(-1, 1, N)
        z = O.ilu0_solve(rp, col, fval, invd, r, starts)
        L = np.tril(F * P, -1) + np.eye(N)
        U = np.triu(F * P)
        assert np.abs(L @ (U @ z) - r).max() < 1e-12


@pytest.mark.parametrize("name,kind,swap,its5,its10", [("tet10", S.POISSON_TETRA, False, 9, 16), ("tria20x20", S.POISSON_TRIA, False, 16, 26),
                                                       ("beam3Dtet6366", S.ELASTICITY_TETRA, True, 131, 145),
                                                       ("cookmembranetria32", S.ELASTICITY_TRIA, False, 129, 170)])
def test_cg_bjacobi_ilu0_known_answers(input_dir, name, kind, swap, its5, its10):
    """Recorded iteration counts of the restated KSPCG + PCBJACOBI/ILU(0) (1 block), agreement with the Jacobi-CG
    solution, and fewer iterations than Jacobi; with more blocks the preconditioner weakens monotonically here."""
    num, rp, col, val, rhs = _assembled(name, kind, input_dir, swap)
    xj, itsj, _, _ = O.cg_jacobi(rp, col, val, rhs, rtol=1e-10)
    x5, i5, r5, _ = O.cg_bjacobi_ilu0(rp, col, val, rhs, rtol=1e-5)
    x10, i10, r10, _ = O.cg_bjacobi_ilu0(rp, col, val, rhs, rtol=1e-10)
    assert (i5, r5, i10, r10) == (its5, 2, its10, 2)
    assert i10 < itsj and np.abs(x10 - xj).max() <= 1e-8 * np.abs(xj).max()
    N = num.size_global
    _, i2, r2, _ = O.cg_bjacobi_ilu0(rp, col, val, rhs, block_start=[0, N // 3, N], rtol=1e-10)
    assert r2 == 2 and i10 <= i2 <= itsj
