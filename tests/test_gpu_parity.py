"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): pattern bit-exact; matrix/RHS values within 1e-12 relative (we assert the
stronger bit-exact equality against the no-FMA oracle, and 1e-12 as the stated contract); CG iteration
counts within +-2 %; solution within the KSP tolerance.
"""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from pfemfort_b200 import driver as D, mesh as M, solver as S
from properties import same_system, values_within, vector_within

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["rows", "fast"], autouse=True)
def asm_mode(request, monkeypatch):
    """Every test of this module runs with both arithmetic modes of the value pass: "rows" = the default (reference-order
    no-FMA operators: bit-identical to the sequential oracle), "fast" = PFEM_ASM=fast (FMA cofactor-form operators:
    1e-12 contract)."""
    if request.param == "fast":
        monkeypatch.setenv("PFEM_ASM", "fast")
    else:
        monkeypatch.delenv("PFEM_ASM", raising=False)
    return request.param

KINDS = [S.POISSON_TRIA, S.POISSON_TETRA, S.ELASTICITY_TRIA, S.ELASTICITY_TETRA]
REL_TOL_VALUES = 1e-12          # north_star tolerance for assembled values
ITS_TOL = 0.02                  # +-2 % iteration count


def _rand_elems(kind, n, seed):
    rng = np.random.default_rng(seed)
    npe, ndof, ndim = S.KIND_DIMS[kind]
    if ndim == 2:
        base = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    else:   # reference ordering: local node 3 is the origin of the parametric map
        base = np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]])
    xyz = base[:, :, None] + 0.2 * rng.standard_normal((ndim, npe, n)) + 3.0 * rng.standard_normal((ndim, 1, n))
    return xyz


@pytest.mark.parametrize("kind", KINDS)
def test_element_batch_bit_exact(gpu, kind):
    npe, ndof, ndim = S.KIND_DIMS[kind]
    n = 257
    xyz = _rand_elems(kind, n, 1234 + kind)
    elemData = {0: [1.3, 0.7], 1: [1.3, 0.7, 2.1], 2: [240.565, 0.3, 0.5, 0.1, -0.2],
                3: [float(np.float32(240.565)), float(np.float32(0.3)), 1.0, float(np.float32(0.1)), 0.05, -0.3]}[kind]
    timeData = [0.0, 0.9, 0.0]
    rng = np.random.default_rng(99)
    valC = rng.standard_normal((npe * ndof, n)) if ndof == 1 else None
    K, F, neg = S.element_ke_batch(kind, xyz[0], xyz[1], xyz[2] if ndim == 3 else None, elemData, timeData, valC)
    for e in range(n):
        Ko, Fo, rc = O.element_ke(kind, xyz[0, :, e], xyz[1, :, e], xyz[2, :, e] if ndim == 3 else None, elemData, timeData,
                                  valC[:, e] if valC is not None else None)
        assert rc == neg[e]
        if rc:
            continue
        assert np.array_equal(K[:, :, e], Ko), f"K differs for element {e}"
        assert np.array_equal(F[:, e], Fo), f"F differs for element {e}"
    assert (neg == 0).sum() > n // 2


@pytest.mark.parametrize("kind", KINDS)
def test_single_element_entry_points(gpu, kind):
    npe, ndof, ndim = S.KIND_DIMS[kind]
    xyz = _rand_elems(kind, 1, 7)[:, :, 0]
    elemData = [2.0, 0.25, 1.5, 0.1, 0.2, 0.3][: {0: 2, 1: 3, 2: 5, 3: 6}[kind]]
    K, F = S.element_ke(kind, xyz[0], xyz[1], xyz[2] if ndim == 3 else None, elemData, [0, 1, 0])
    Ko, Fo, rc = O.element_ke(kind, xyz[0], xyz[1], xyz[2] if ndim == 3 else None, elemData, [0, 1, 0])
    assert rc == 0 and np.array_equal(K, Ko) and np.array_equal(F, Fo)
    # negative Jacobian -> error code where the reference STOPs
    sw = xyz.copy()
    sw[:, [0, 1]] = sw[:, [1, 0]]
    with pytest.raises(S.PfemError) as ei:
        S.element_ke(kind, sw[0], sw[1], sw[2] if ndim == 3 else None, elemData, [0, 1, 0])
    assert ei.value.status == S.ERR_NEG_JACOBIAN


def _load(name, input_dir):
    if name == "beam3Dtet6366":
        return M.read_mesh(os.path.join(input_dir, name), swap_34=True), S.ELASTICITY_TETRA
    kind = {"tria20x20": S.POISSON_TRIA, "tet10": S.POISSON_TETRA, "cookmembranetria32": S.ELASTICITY_TRIA}[name]
    return M.read_mesh(os.path.join(input_dir, name)), kind


def _oracle_system(m, kind, num, fbc=True):
    npe, ndof, ndim = S.KIND_DIMS[kind]
    rp, col = O.pattern(num.elemDof, num.size_global)
    val, rhs, nbad = O.assemble(kind, num.conn_new, m.coords, num.node_map_get_old, num.elemDof, num.solnApplied,
                                D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA, rp, col)
    assert nbad == 0
    if fbc and m.fbc_node.size:
        O.add_force_bc(rhs, m.fbc_node, m.fbc_dof, m.fbc_val, ndof, num.node_map_get_new, num.NodeDofArrayNew, num.size_global)
    return rp, col, val, rhs


@pytest.mark.parametrize("name", ["tria20x20", "tet10", "cookmembranetria32", "beam3Dtet6366"])
def test_fixture_assembly_and_solve(gpu, input_dir, name, asm_mode):
    m, kind = _load(name, input_dir)
    num = D.number(m, kind)
    s = S.SolverB200(0)
    info = D.run_rank(s, m, num, rtol=1e-10)
    rp, col, val = s.get_csr()
    rhs = s.get_rhs()
    x = s.get_solution()
    orp, ocol, oval, orhs = _oracle_system(m, kind, num)
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol)                 # pattern: bit-exact
    scale = np.abs(oval).max()
    assert np.abs(val - oval).max() <= REL_TOL_VALUES * scale                    # contract
    assert same_system(s, rp, val, oval, rhs, orhs), "values differ from the oracle beyond the bar of the kernel that ran"
    if asm_mode == "rows":
        assert s.assembly_mode()[0] in (0, 1) and np.array_equal(val, oval) and np.array_equal(rhs, orhs)
    else:
        assert s.assembly_mode()[0] == 4
    ox, oits, oreason, ornorm = O.cg_jacobi(orp, ocol, oval, orhs, rtol=1e-10)
    assert info["reason"] == oreason == 2
    assert abs(info["its"] - oits) <= max(1, ITS_TOL * oits), (info["its"], oits)
    assert np.abs(x - ox).max() <= 1e-7 * np.abs(ox).max()
    s.free()


def test_known_answers(gpu, input_dir):
    """The analytic solutions baked into the reference's BC files."""
    m, kind = _load("tria20x20", input_dir)
    num = D.number(m, kind)
    s = S.SolverB200(0)
    D.run_rank(s, m, num, rtol=1e-10)
    u = D.nodal_solution(num, s.get_solution())[0]
    exact = M.exact_poisson_tria(m.coords[0], m.coords[1])
    assert abs(np.abs(u - exact).max() - 7.1146e-4) < 1e-6
    s.free()
    m, kind = _load("tet10", input_dir)
    num = D.number(m, kind)
    s = S.SolverB200(0)
    D.run_rank(s, m, num, rtol=1e-10)
    u = D.nodal_solution(num, s.get_solution())[0]
    assert np.abs(u - (m.coords ** 2).sum(0)).max() < 2e-7
    s.free()


def test_state_machine_and_slow_path(gpu, input_dir):
    m, kind = _load("tet10", input_dir)
    num = D.number(m, kind)
    s = S.SolverB200(0)
    s.initialise(num.size_global, num.size_global)
    with pytest.raises(S.PfemError) as ei:
        s.setZero()
    assert ei.value.status == S.ERR_STATE
    s.set_mesh(kind, num.conn_new, m.coords)
    s.set_pattern(num.elemDof)
    with pytest.raises(S.PfemError) as ei:
        s.factorise()                       # "Assemble matrix first before solving it!"
    assert ei.value.status == S.ERR_STATE
    with pytest.raises(S.PfemError) as ei:
        s.solve()                           # "Factorise matrix first before solving it!"
    assert ei.value.status == S.ERR_STATE
    # per-element MatSetValues / VecSetValues mirrors reproduce the batched value pass
    s.setZero()
    ed, td = D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA
    nE = 40
    for e in range(nE):
        nodes = num.conn_new[:, e] - 1
        K, F, rc = O.element_ke(kind, m.coords[0, nodes], m.coords[1, nodes], m.coords[2, nodes], ed, td)
        dofs = num.elemDof[:, e]
        s.add_matrix(dofs, dofs, K)
        for ii in range(4):
            if dofs[ii] == -1:
                for jj in range(4):
                    if dofs[jj] != -1:
                        F[jj] = F[jj] - K[jj, ii] * num.solnApplied[nodes[ii]]
        s.add_vector(dofs, F)
    rp, col, val = s.get_csr()
    rhs = s.get_rhs()
    mask = np.zeros(m.nElem, np.uint8)
    mask[:nE] = 1
    oval, orhs, _ = O.assemble(kind, num.conn_new, m.coords, None, num.elemDof, num.solnApplied, ed, td, rp, col, elem_mask=mask)
    assert np.array_equal(val, oval) and np.array_equal(rhs, orhs)
    s.free()


def test_accumulate_without_set_zero_and_determinism(gpu, input_dir):
    m, kind = _load("tet10", input_dir)
    num = D.number(m, kind)
    s = S.SolverB200(0)
    D.run_rank(s, m, num, do_solve=False)
    _, _, v1 = s.get_csr()
    r1 = s.get_rhs()
    s.assemble(D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA)      # ADD on top, like a second MatSetValues sweep
    _, _, v2 = s.get_csr()
    rp, col = O.pattern(num.elemDof, num.size_global)
    oval, orhs, _ = O.assemble(kind, num.conn_new, m.coords, None, num.elemDof, num.solnApplied,
                               D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA, rp, col)
    oval, orhs, _ = O.assemble(kind, num.conn_new, m.coords, None, num.elemDof, num.solnApplied,
                               D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA, rp, col, val=oval, rhs=orhs)
    assert same_system(s, rp, v2, oval, s.get_rhs(), orhs)
    s.setZero()
    s.assemble(D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA)
    _, _, v3 = s.get_csr()
    assert np.array_equal(v3, v1) and np.array_equal(s.get_rhs(), r1)   # run-to-run bit-identical
    s.free()


def test_negative_jacobian_is_reported(gpu, input_dir):
    m = M.read_mesh(os.path.join(input_dir, "beam3Dtet6366"))          # as shipped: all 7776 Jacobians negative
    num = D.number(m, S.ELASTICITY_TETRA)
    s = S.SolverB200(0)
    with pytest.raises(S.PfemError) as ei:
        D.run_rank(s, m, num)
    assert ei.value.status == S.ERR_NEG_JACOBIAN
    s.free()


def test_zero_rhs_converges_at_iteration_zero(gpu, input_dir):
    m, kind = _load("cookmembranetria32", input_dir)
    num = D.number(m, kind)
    s = S.SolverB200(0)
    info = D.run_rank(s, m, num, apply_force_bc=False)     # b = 0: PETSc converges at iteration 0, x = 0
    assert info["its"] == 0 and info["reason"] == 3
    assert np.all(s.get_solution() == 0.0)
    s.free()


@pytest.mark.parametrize("n", [100])
def test_generated_tria_and_tet(gpu, n):
    """Mid-size generated meshes (tria100x100, tet 20^3): oracle finishes in seconds."""
    m = M.gen_tria_poisson(n)
    num = D.number(m, S.POISSON_TRIA)
    s = S.SolverB200(0)
    info = D.run_rank(s, m, num, rtol=1e-10)
    rp, col, val = s.get_csr()
    orp, ocol, oval, orhs = _oracle_system(m, S.POISSON_TRIA, num)
    assert col.size == 67817                                   # explicit zeros kept (SURVEY.md 8d pattern trap)
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol) and same_system(s, rp, val, oval, s.get_rhs(), orhs)
    ox, oits, _, _ = O.cg_jacobi(orp, ocol, oval, orhs, rtol=1e-10)
    assert abs(info["its"] - oits) <= max(1, ITS_TOL * oits)
    s.free()
    m = M.gen_tetra(-1, 1, 20, -1, 1, 20, -1, 1, 20)
    num = D.number(m, S.POISSON_TETRA)
    s = S.SolverB200(0)
    info = D.run_rank(s, m, num, rtol=1e-10)
    rp, col, val = s.get_csr()
    orp, ocol, oval, orhs = _oracle_system(m, S.POISSON_TETRA, num)
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol) and same_system(s, rp, val, oval, s.get_rhs(), orhs)
    ox, oits, _, _ = O.cg_jacobi(orp, ocol, oval, orhs, rtol=1e-10)
    assert abs(info["its"] - oits) <= max(1, ITS_TOL * oits)
    u = D.nodal_solution(num, s.get_solution())[0]
    assert np.abs(u - (m.coords ** 2).sum(0)).max() < 1e-6
    s.free()


def test_generic_assembly_kernel_matches_streamed(gpu, input_dir, monkeypatch, asm_mode):
    """Rows wider than 254 entries fall back to the binary-search kernel; force it and compare bit for bit (under
    PFEM_ASM=fast the fallback keeps the reference-order operators: 1e-12 contract between the two)."""
    for name in ("tet10", "beam3Dtet6366"):
        m, kind = _load(name, input_dir)
        num = D.number(m, kind)
        out = []
        for force in ("0", "1"):
            monkeypatch.setenv("PFEM_FORCE_GENERIC_ASM", force)
            s = S.SolverB200(0)
            D.run_rank(s, m, num, do_solve=False)
            out.append((s.get_csr()[2], s.get_rhs()))
            s.free()
        if asm_mode == "rows":
            assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
        else:
            scale = np.abs(out[1][0]).max()
            assert np.abs(out[0][0] - out[1][0]).max() <= REL_TOL_VALUES * scale
            assert np.abs(out[0][1] - out[1][1]).max() <= REL_TOL_VALUES * max(np.abs(out[1][1]).max(), 1e-300)


# (PFEM_CG, PFEM_CG_SR, PFEM_PCG_FUSED, PFEM_PCG_SYNC, PFEM_PCG_CFG); entry 1 is the launch-per-phase reference point
CG_VARIANTS = (("persistent", "0", "0", "", ""), ("kernels", "0", "0", "", ""), ("persistent", "1", "0", "", ""),
               ("persistent", "0", "1", "", ""), ("persistent", "0", "0", "last", "1024x1"), ("persistent", "0", "0", "lean", "1024x1"),
               ("persistent", "0", "0", "lean", "256x5"), ("persistent", "1", "0", "lean", "512x2"), ("persistent", "0", "1", "last", "256x5"))


def _set_cg_variant(monkeypatch, v):
    mode, sr, fused, sync, cfg = v
    monkeypatch.setenv("PFEM_CG", mode)
    monkeypatch.setenv("PFEM_CG_SR", sr)     # 1: PETSc's single-reduction recurrences
    monkeypatch.setenv("PFEM_PCG_FUSED", fused)   # 1: direction folded into the SpMV
    for key, val in (("PFEM_PCG_SYNC", sync), ("PFEM_PCG_CFG", cfg)):   # barrier flavour / CTA shape of the persistent kernel
        if val:
            monkeypatch.setenv(key, val)
        else:
            monkeypatch.delenv(key, raising=False)


def test_persistent_and_multi_kernel_cg_agree(gpu, input_dir, monkeypatch):
    """The persistent cooperative CG kernel (default; both barrier flavours, every CTA shape) and the launch-per-phase
    path follow the same PETSc semantics: same reason, same iteration count, same solution to rounding."""
    for name in ("tet10", "cookmembranetria32"):
        m, kind = _load(name, input_dir)
        num = D.number(m, kind)
        res = []
        for v in CG_VARIANTS:
            _set_cg_variant(monkeypatch, v)
            s = S.SolverB200(0)
            info = D.run_rank(s, m, num, rtol=1e-10)
            res.append((info["its"], info["reason"], s.get_solution()))
            info2 = D.run_rank(s, m, num, rtol=1e-10)          # a second solve on the same handle (tags carry the solve number)
            assert (info2["its"], info2["reason"]) == (info["its"], info["reason"])
            assert np.array_equal(s.get_solution(), res[-1][2]), "run-to-run determinism"
            s.free()
        assert all(r[1] == 2 for r in res)
        for k, v in enumerate(CG_VARIANTS):
            tol = max(1, ITS_TOL * res[1][0]) if v[1] == "1" else 1
            assert abs(res[k][0] - res[1][0]) <= tol, (v, res[k][0], res[1][0])
            assert np.abs(res[k][2] - res[1][2]).max() <= 1e-8 * np.abs(res[1][2]).max(), v
    # max_it reached -> DIVERGED_ITS (-3) with its = max_it, on every path
    m, kind = _load("tet10", input_dir)
    num = D.number(m, kind)
    for v in CG_VARIANTS:
        _set_cg_variant(monkeypatch, v)
        s = S.SolverB200(0)
        info = D.run_rank(s, m, num, rtol=1e-10, max_it=7)
        assert (info["its"], info["reason"]) == (7, -3), v
        s.free()


def test_nodal_pattern_pass_equals_element_dof_pattern_pass(gpu, input_dir):
    """pfem_solver_set_pattern_nodal forms ElemDofArray on the GPU (tetrapoissonparallelimpl1.F:698-713)."""
    for name in ("tet10", "beam3Dtet6366", "cookmembranetria32"):
        m, kind = _load(name, input_dir)
        num = D.number(m, kind)
        out = []
        for nodal in (False, True):
            s = S.SolverB200(0)
            s.initialise(num.size_global, num.size_global)
            s.set_mesh(kind, num.conn_new, m.coords)
            if nodal:
                s.set_pattern_nodal(num.NodeDofArrayNew)
            else:
                s.set_pattern(num.elemDof)
            s.setZero()
            s.set_applied(num.solnApplied)
            s.assemble(D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA)
            out.append(s.get_csr() + (s.get_rhs(),))
            s.free()
        for a, b in zip(out[0], out[1]):
            assert np.array_equal(a, b)


def _check_against_oracle(m, kind, num=None, rtol=1e-10):
    num = num or D.number(m, kind)
    s = S.SolverB200(0)
    info = D.run_rank(s, m, num, rtol=rtol)
    rp, col, val = s.get_csr()
    rhs = s.get_rhs()
    x = s.get_solution()
    orp, ocol, oval, orhs = _oracle_system(m, kind, num)
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
    assert same_system(s, rp, val, oval, rhs, orhs)
    s.free()
    ox, oits, oreason, _ = O.cg_jacobi(orp, ocol, oval, orhs, rtol=rtol)
    assert info["reason"] == oreason
    assert abs(info["its"] - oits) <= max(1, ITS_TOL * oits)
    assert np.abs(x - ox).max() <= 1e-6 * max(np.abs(ox).max(), 1e-300)
    return num, x


def test_nonzero_and_partial_dirichlet_elasticity(gpu, input_dir):
    """Lifting with ndof > 1: non-zero applied displacements, and nodes where only SOME dofs are constrained
    (tetraelasticityparallelimpl1.F:938-958 indexes solnApplied by (node-1)*ndof+dof)."""
    m, kind = _load("beam3Dtet6366", input_dir)
    m.fbc_node = m.fbc_node[:0]; m.fbc_dof = m.fbc_dof[:0]; m.fbc_val = m.fbc_val[:0]
    top = np.flatnonzero(np.abs(m.coords[1] - m.coords[1].max()) < 1e-9) + 1          # y = 6 face
    m.dbc_node = np.concatenate([m.dbc_node, top, top[::2]]).astype(np.int32)
    m.dbc_dof = np.concatenate([m.dbc_dof, np.full(top.size, 2), np.full(top[::2].size, 1)]).astype(np.int32)   # uy everywhere, ux on every other node
    m.dbc_val = np.concatenate([m.dbc_val, np.full(top.size, 0.05), 0.01 * np.arange(top[::2].size)])
    _check_against_oracle(m, kind)
    m2, kind2 = _load("cookmembranetria32", input_dir)
    right = np.flatnonzero(np.abs(m2.coords[0] - m2.coords[0].max()) < 1e-9) + 1
    m2.dbc_node = np.concatenate([m2.dbc_node, right]).astype(np.int32)
    m2.dbc_dof = np.concatenate([m2.dbc_dof, np.full(right.size, 2)]).astype(np.int32)
    m2.dbc_val = np.concatenate([m2.dbc_val, np.linspace(0.1, 0.3, right.size)])
    _check_against_oracle(m2, kind2)


def test_randomly_renumbered_mesh(gpu, input_dir):
    """Unstructured numbering: nodes and elements of tet10 / cookmembrane randomly permuted (wide, irregular rows, no
    locality).  The solution must be the permuted solution of the original mesh."""
    rng = np.random.default_rng(2024)
    for name in ("tet10", "cookmembranetria32"):
        m, kind = _load(name, input_dir)
        npe, ndof, ndim = S.KIND_DIMS[kind]
        perm = rng.permutation(m.nNode)                 # new position of old node n is inv[n]
        inv = np.empty(m.nNode, np.int64)
        inv[perm] = np.arange(m.nNode)
        eperm = rng.permutation(m.nElem)
        m2 = M.Mesh(np.ascontiguousarray(m.coords[:, perm]), np.ascontiguousarray((inv[m.conn - 1] + 1)[:, eperm]).astype(np.int32),
                    (inv[m.dbc_node - 1] + 1).astype(np.int32), m.dbc_dof.copy(), m.dbc_val.copy(),
                    (inv[m.fbc_node - 1] + 1).astype(np.int32) if m.fbc_node.size else m.fbc_node, m.fbc_dof, m.fbc_val, name + "-perm")
        num2, x2 = _check_against_oracle(m2, kind, rtol=1e-12)
        if m.fbc_node.size:
            continue        # the reference's node-based ForceBC row formula is numbering dependent by construction
        num1 = D.number(m, kind)
        s = S.SolverB200(0)
        D.run_rank(s, m, num1, rtol=1e-12)
        u1 = D.nodal_solution(num1, s.get_solution())
        s.free()
        u2 = D.nodal_solution(num2, x2)
        assert np.abs(u2[:, inv] - u1).max() <= 1e-8 * np.abs(u1).max()


def test_element_with_no_free_dof_and_single_element_mesh(gpu):
    # one triangle, two Dirichlet nodes: a 1 x 1 system
    coords = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    m = M.Mesh(coords, np.array([[1], [2], [3]], np.int32), np.array([1, 2], np.int32), np.array([1, 1], np.int32), np.array([2.0, -1.0]))
    num, x = _check_against_oracle(m, S.POISSON_TRIA)
    assert num.size_global == 1
    # two triangles, the second one entirely on the Dirichlet boundary (no free dof): it must be ignored, not crash
    coords = np.array([[0.0, 1.0, 0.0, 1.0], [0.0, 0.0, 1.0, 1.0]])
    m = M.Mesh(coords, np.array([[1, 2], [2, 4], [3, 3]], np.int32), np.array([2, 3, 4], np.int32), np.ones(3, np.int32), np.array([1.0, 2.0, 3.0]))
    num, x = _check_against_oracle(m, S.POISSON_TRIA)
    assert num.size_global == 1


def test_petscsolver_assemble_mirrors(gpu, input_dir):
    """PetscSolver%assembleMatrix / assembleVector / assembleMatrixAndVector (solverpetsc.F:328-401): entry
    (R(ii), C(jj)) += KLOCAL(ii,jj), i.e. NOT transposed, unlike the drivers' direct MatSetValues."""
    m, kind = _load("tet10", input_dir)
    num = D.number(m, kind)
    s = S.SolverB200(0)
    s.initialise(num.size_global, num.size_global)
    s.set_mesh(kind, num.conn_new, m.coords)
    s.set_pattern(num.elemDof)
    s.setZero()
    rng = np.random.default_rng(3)
    rp, col, _ = s.get_csr()
    ref_val = np.zeros(col.size)
    ref_rhs = np.zeros(num.size_global)

    def ref_add(r, c, v):
        if r < 0 or c < 0:
            return
        k = rp[r] + np.searchsorted(col[rp[r]:rp[r + 1]], c)
        ref_val[k] += v

    for e in (0, 17, 4321, 5999):
        dofs = num.elemDof[:, e]
        K = rng.standard_normal((4, 4))
        F = rng.standard_normal(4)
        if e % 2:
            s.assembleMatrix(dofs, dofs, K)
            s.assembleVector(dofs, F)
        else:
            s.assembleMatrixAndVector(dofs, dofs, K, F)
        for i in range(4):
            if dofs[i] >= 0:
                ref_rhs[dofs[i]] += F[i]
            for j in range(4):
                ref_add(dofs[i], dofs[j], K[i, j])
    _, _, val = s.get_csr()
    assert np.array_equal(val, ref_val) and np.array_equal(s.get_rhs(), ref_rhs)
    # ForceBC-style single adds, negative / foreign rows ignored
    s.add_value(5, 1.5)
    s.add_value(-1, 9.0)
    ref_rhs[5] += 1.5
    assert np.array_equal(s.get_rhs(), ref_rhs)
    assert s.launch_count() > 0 and s.time_spmv(3) > 0.0
    s.free()


# ---- the reference's default preconditioner: PCBJACOBI / ILU(0) (solverpetsc.F:206) ----------------------------------

@pytest.mark.parametrize("name", ["tet10", "tria20x20", "beam3Dtet6366", "cookmembranetria32"])
def test_bjacobi_ilu0_cg_matches_oracle(gpu, input_dir, name):
    """CG + block-Jacobi/ILU(0) on one rank (= ILU(0) of the whole matrix): the factor and the inverted pivots are
    bit-identical to the sequential oracle (the sync-free kernels keep the sequential operation order), iteration
    counts and reason equal the oracle's, solutions agree to the solver tolerance."""
    m, kind = _load(name, input_dir)
    num = D.number(m, kind)
    s = S.SolverB200(0)
    for rtol in (1e-5, 1e-10):
        info = D.run_rank(s, m, num, rtol=rtol, pc_type=S.PC_BJACOBI_ILU0)
        rp, col, val = s.get_csr()
        rhs = s.get_rhs()
        ofv, oinv, rc = O.ilu0_factor(rp, col, val)
        assert rc == 0
        fv, inv = s.get_ilu_factor()
        assert np.array_equal(fv, ofv) and np.array_equal(inv, oinv), "ILU(0) factor differs from the oracle"
        ox, oits, oreason, _ = O.cg_bjacobi_ilu0(rp, col, val, rhs, rtol=rtol)
        assert info["reason"] == oreason == 2
        assert abs(info["its"] - oits) <= max(1, ITS_TOL * oits), (info["its"], oits)
        x = s.get_solution()
        assert np.abs(x - ox).max() <= 10 * rtol * np.abs(ox).max()
        info2 = D.run_rank(s, m, num, rtol=rtol, pc_type=S.PC_BJACOBI_ILU0)
        assert (info2["its"], info2["reason"]) == (info["its"], info["reason"]) and np.array_equal(s.get_solution(), x), "determinism"
    s.free()


def test_default_pc_is_the_references_and_options_file(gpu, input_dir, tmp_path):
    """pfem_solver_initialise leaves the reference's coded defaults (CG + PCBJACOBI/ILU(0), rtol 1e-5); an options file
    in PETSc's syntax overrides them; unsupported types are refused."""
    m, kind = _load("tet10", input_dir)
    num = D.number(m, kind)
    s = S.SolverB200(0)
    info = D.run_rank(s, m, num, rtol=-1.0, pc_type=-1)            # nothing overridden
    rp, col, val = s.get_csr()
    ox, oits, oreason, _ = O.cg_bjacobi_ilu0(rp, col, val, s.get_rhs(), rtol=1e-5)
    assert (info["its"], info["reason"]) == (oits, oreason)
    opt = tmp_path / "petsc_options.dat"
    opt.write_text("-ksp_type cg\n-pc_type jacobi   # north-star run\n-ksp_rtol 1e-10\n-log_view\n")
    s.set_options_from_file(str(opt))
    s.factoriseAndSolve()                                          # same assembled system, new options
    ox, oits, oreason, _ = O.cg_jacobi(rp, col, val, s.get_rhs(), rtol=1e-10)
    assert abs(s.info()["its"] - oits) <= 1 and s.info()["reason"] == oreason
    opt.write_text("-pc_type gamg\n")
    with pytest.raises(S.PfemError):
        s.set_options_from_file(str(opt))
    s.set_options_from_file(str(tmp_path / "absent.dat"))          # no file: defaults stay
    s.free()


def test_slow_path_never_drops_silently(gpu, input_dir):
    """A MatSetValues-style add on a location the pattern pass did not create is reported (PFEM_ERR_PATTERN), an index beyond
    the global system is an argument error (PETSc: 'Row too large'); negative indices are skipped like PETSc does."""
    m, kind = _load("tet10", input_dir)
    num = D.number(m, kind)
    s = S.SolverB200(0)
    D.run_rank(s, m, num, do_solve=False)
    rp, col, v0 = s.get_csr()
    far = int(np.setdiff1d(np.arange(num.size_global), col[rp[0]:rp[1]])[0])       # a dof not coupled with dof 0
    with pytest.raises(S.PfemError) as ei:
        s.add_matrix([0, far], [0, far], np.ones((2, 2)))
    assert ei.value.status == S.ERR_PATTERN
    v1 = s.get_csr()[2]
    k00 = rp[0] + int(np.searchsorted(col[rp[0]:rp[1]], 0))
    assert v1[k00] == v0[k00] + 1.0                                                # the entries inside the pattern were added
    with pytest.raises(S.PfemError) as ei:
        s.add_vector([num.size_global], [1.0])
    assert ei.value.status == S.ERR_ARG
    r0 = s.get_rhs()
    s.add_matrix([-1, 0], [-1, 0], np.full((2, 2), 2.0))                           # negative row / column: dropped, no error
    s.add_vector([-1], [5.0])
    assert np.array_equal(s.get_rhs(), r0) and s.get_csr()[2][k00] == v0[k00] + 3.0
    s.free()
