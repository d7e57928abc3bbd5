"""GPU parity of the colour-scheduled tile value pass (opt-in, PFEM_ASM=ctile, Poisson kinds; csrc/assembly_ctile.cu/.cuh):
each element computed once per tile, FMA arithmetic, a fixed summation order that is not the sequential one.  Bars:
pattern untouched (bit-exact, built by the pattern pass), values / RHS within 1e-12 relative (scale = the row's largest
entry) of the sequential no-FMA oracle -- the north-star contract --, run-to-run bit-identity, identical results for every
tile size / CTA shape is NOT required (the order depends on the tiling) but each must meet the 1e-12 bar."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from pfemfort_b200 import driver as D, mesh as M, solver as S
from properties import values_within, vector_within

pytestmark = pytest.mark.gpu


@pytest.fixture()
def env():
    keys = ("PFEM_ASM", "PFEM_TILE_ROWS", "PFEM_TILE_THREADS", "PFEM_TILE_RULE", "PFEM_TILE_GRID")
    old = {k: os.environ.get(k) for k in keys}
    for k in keys:
        os.environ.pop(k, None)
    os.environ["PFEM_ASM"] = "ctile"
    yield os.environ
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _assemble(m, kind, num, elemData=None, timeData=None, twice=False, rank=0):
    s = S.SolverB200(0)
    D.run_rank(s, m, num, elemData=elemData, timeData=timeData, do_solve=False)
    if twice:
        s.assemble(D.DEFAULT_ELEMDATA[kind] if elemData is None else elemData, D.DEFAULT_TIMEDATA if timeData is None else timeData)
    rp, col, val = s.get_csr()
    rhs = s.get_rhs()
    mode = s.assembly_mode()
    info = s.assembly_info()
    s.free()
    return rp, col, val, rhs, mode, info


def _oracle(m, kind, num, elemData=None, timeData=None):
    rp, col = O.pattern(num.elemDof, num.size_global)
    val, rhs, nbad = O.assemble(kind, num.conn_new, m.coords, num.node_map_get_old, num.elemDof, num.solnApplied,
                                D.DEFAULT_ELEMDATA[kind] if elemData is None else elemData,
                                D.DEFAULT_TIMEDATA if timeData is None else timeData, rp, col)
    assert nbad == 0
    return rp, col, val, rhs


CASES = {
    "tria20x20": lambda d: (M.read_mesh(os.path.join(d, "tria20x20")), S.POISSON_TRIA),
    "tet10": lambda d: (M.read_mesh(os.path.join(d, "tet10")), S.POISSON_TETRA),
    "gen_tet_17x13x11": lambda d: (M.gen_tetra(-1, 1, 17, -1, 1, 13, -1, 1, 11), S.POISSON_TETRA),
    "gen_tet_40": lambda d: (M.gen_tetra(-1, 1, 40, -1, 1, 40, -1, 1, 40), S.POISSON_TETRA),
    "gen_tria_61": lambda d: (M.gen_tria_poisson(61), S.POISSON_TRIA),
    "gen_tria_300": lambda d: (M.gen_tria_poisson(300), S.POISSON_TRIA),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("threads,rows,rule,grid", [(384, 1024, "full", 0), (512, 1024, "position", 0), (1024, 1024, "position", 3),
                                                     (256, 96, "full", 2), (768, 320, "position", 0)])
def test_ctile_value_pass_within_contract(gpu, input_dir, env, name, threads, rows, rule, grid):
    m, kind = CASES[name](input_dir)
    num = D.number(m, kind)
    env["PFEM_TILE_THREADS"] = str(threads)
    env["PFEM_TILE_ROWS"] = str(rows)
    env["PFEM_TILE_RULE"] = rule
    if grid:
        env["PFEM_TILE_GRID"] = str(grid)      # several tiles per persistent CTA even on a small mesh
    rp, col, val, rhs, mode, info = _assemble(m, kind, num)
    assert mode[0] == 3, info
    orp, ocol, oval, orhs = _oracle(m, kind, num)
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
    assert values_within(rp, val, oval, 1e-12), float(np.abs(val - oval).max() / np.abs(oval).max())
    assert vector_within(rhs, orhs, 1e-12)
    # run-to-run: the schedule is fixed => bit-identical
    _, _, val2, rhs2, _, _ = _assemble(m, kind, num)
    assert np.array_equal(val, val2) and np.array_equal(rhs, rhs2)


def test_ctile_non_unit_coefficients_interior_dirichlet_and_accumulate(gpu, input_dir, env):
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    kind = S.POISSON_TETRA
    # Dirichlet nodes in the interior too (lifting from every side), non-unit conductivities and af
    rng = np.random.default_rng(5)
    extra = rng.choice(m.nNode, 60, replace=False) + 1
    extra = np.setdiff1d(extra, m.dbc_node)
    m.dbc_node = np.concatenate([m.dbc_node, extra]).astype(np.int32)
    m.dbc_dof = np.concatenate([m.dbc_dof, np.ones(extra.size, np.int32)]).astype(np.int32)
    m.dbc_val = np.concatenate([m.dbc_val, rng.standard_normal(extra.size)])
    num = D.number(m, kind)
    ed, td = [1.7, 0.6, 2.3], [0.0, 0.8, 0.0]
    rp, col, val, rhs, mode, _ = _assemble(m, kind, num, ed, td)
    orp, ocol, oval, orhs = _oracle(m, kind, num, ed, td)
    assert mode[0] == 3 and np.array_equal(col, ocol)
    assert values_within(rp, val, oval) and vector_within(rhs, orhs)
    # ADD on top without setZero (a second MatSetValues sweep)
    _, _, v2, r2, _, _ = _assemble(m, kind, num, ed, td, twice=True)
    o2, or2, _ = O.assemble(kind, num.conn_new, m.coords, num.node_map_get_old, num.elemDof, num.solnApplied, ed, td, orp, ocol,
                            val=oval.copy(), rhs=orhs.copy())
    assert values_within(rp, v2, o2) and vector_within(r2, or2)


def test_ctile_randomly_renumbered_and_solve(gpu, input_dir, env):
    """No locality in the numbering (the Morton order comes from the coordinates, not from the ids); full solve."""
    rng = np.random.default_rng(11)
    m = M.gen_tetra(-1, 1, 14, -1, 1, 14, -1, 1, 14)
    perm = rng.permutation(m.nNode)
    inv = np.empty(m.nNode, np.int64)
    inv[perm] = np.arange(m.nNode)
    eperm = rng.permutation(m.nElem)
    m2 = M.Mesh(np.ascontiguousarray(m.coords[:, perm]), np.ascontiguousarray((inv[m.conn - 1] + 1)[:, eperm]).astype(np.int32),
                (inv[m.dbc_node - 1] + 1).astype(np.int32), m.dbc_dof.copy(), m.dbc_val.copy(), name="tet14-perm")
    kind = S.POISSON_TETRA
    num = D.number(m2, kind)
    s = S.SolverB200(0)
    info = D.run_rank(s, m2, num, rtol=1e-10)
    rp, col, val = s.get_csr()
    orp, ocol, oval, orhs = _oracle(m2, kind, num)
    assert s.assembly_mode()[0] == 3 and np.array_equal(col, ocol)
    assert values_within(rp, val, oval) and vector_within(s.get_rhs(), orhs)
    ox, oits, oreason, _ = O.cg_jacobi(orp, ocol, oval, orhs, rtol=1e-10)
    assert info["reason"] == oreason == 2 and abs(info["its"] - oits) <= max(1, 0.02 * oits)
    u = D.nodal_solution(num, s.get_solution())[0]
    assert np.abs(u - (m2.coords ** 2).sum(0)).max() < 1e-6
    s.free()


def test_ctile_negative_jacobian_and_mode_api(gpu, input_dir, env):
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    m.conn[[0, 1]] = m.conn[[1, 0]]            # every tetrahedron inverted
    num = D.number(m, S.POISSON_TETRA)
    s = S.SolverB200(0)
    with pytest.raises(S.PfemError) as ei:
        D.run_rank(s, m, num)
    assert ei.value.status == S.ERR_NEG_JACOBIAN and s.assembly_mode()[0] == 3
    s.free()
    # pfem_solver_set_assembly_mode: the handle-level switch to the bit-identical row kernels
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    num = D.number(m, S.POISSON_TETRA)
    del env["PFEM_ASM"]
    s = S.SolverB200(0)
    s.set_assembly_mode(S.ASM_ROWS)
    D.run_rank(s, m, num, do_solve=False)
    orp, ocol, oval, orhs = _oracle(m, S.POISSON_TETRA, num)
    assert s.assembly_mode()[0] == 1 and np.array_equal(s.get_csr()[2], oval) and np.array_equal(s.get_rhs(), orhs)
    s.free()
    s = S.SolverB200(0)                       # handle-level switch to the FMA row gather
    s.set_assembly_mode(S.ASM_FAST)
    D.run_rank(s, m, num, do_solve=False)
    rp = s.get_csr()[0]
    assert s.assembly_mode()[0] == 4 and values_within(rp, s.get_csr()[2], oval) and vector_within(s.get_rhs(), orhs)
    s.free()
