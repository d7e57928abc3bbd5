"""GPU parity of the tiled (compute-once) value pass (PFEM_ASM=tiled; csrc/assembly_tiled.cuh, tiles.hpp):
bit-identical to the default row-gather kernel and to the no-FMA oracle, through the C ABI."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from pfemfort_b200 import driver as D, mesh as M, solver as S

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tiled_env():
    old = {k: os.environ.get(k) for k in ("PFEM_ASM", "PFEM_TILE_ROWS", "PFEM_TILE_THREADS", "PFEM_TILE_SMEM_KB")}
    os.environ["PFEM_ASM"] = "tiled"
    yield os.environ
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _assemble(m, kind, num, elemData=None, timeData=None, twice=False):
    s = S.SolverB200(0)
    D.run_rank(s, m, num, elemData=elemData, timeData=timeData, do_solve=False)
    if twice:
        s.assemble(D.DEFAULT_ELEMDATA[kind] if elemData is None else elemData, D.DEFAULT_TIMEDATA if timeData is None else timeData)
    rp, col, val = s.get_csr()
    rhs = s.get_rhs()
    mode = s.assembly_mode()
    s.free()
    return rp, col, val, rhs, mode


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
    "gen_tria_61": lambda d: (M.gen_tria_poisson(61), S.POISSON_TRIA),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("threads,rows", [(256, 96), (128, 32), (512, 192)])
def test_tiled_value_pass_bit_identical(gpu, input_dir, tiled_env, name, threads, rows):
    m, kind = CASES[name](input_dir)
    num = D.number(m, kind)
    tiled_env["PFEM_TILE_THREADS"] = str(threads)
    tiled_env["PFEM_TILE_ROWS"] = str(rows)
    rp, col, val, rhs, mode = _assemble(m, kind, num)
    assert mode[0] == 2 and mode[1] >= (num.size_global + rows - 1) // rows and mode[2] >= 1.0
    orp, ocol, oval, orhs = _oracle(m, kind, num)
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
    assert np.array_equal(val, oval), "tiled values not bit-identical to the no-FMA oracle"
    assert np.array_equal(rhs, orhs)
    # and to the row-gather kernel
    tiled_env["PFEM_ASM"] = "rows"
    _, _, dval, drhs, dmode = _assemble(m, kind, num)
    assert dmode[0] == 1 and np.array_equal(val, dval) and np.array_equal(rhs, drhs)


def test_tiled_non_unit_coefficients_dirichlet_interior_and_accumulate(gpu, input_dir, tiled_env):
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    rng = np.random.default_rng(5)
    extra = np.setdiff1d(rng.choice(m.nNode, 150, replace=False) + 1, m.dbc_node)
    m.dbc_node = np.concatenate([m.dbc_node, extra.astype(np.int32)])
    m.dbc_dof = np.ones(m.dbc_node.size, np.int32)
    m.dbc_val = np.concatenate([m.dbc_val, rng.standard_normal(extra.size)])
    kind = S.POISSON_TETRA
    num = D.number(m, kind)
    ed, td = [1.3, 0.7, 2.1], [0.0, 0.9, 0.0]
    _, _, val, rhs, mode = _assemble(m, kind, num, ed, td)
    _, _, oval, orhs = _oracle(m, kind, num, ed, td)
    assert mode[0] == 2 and np.array_equal(val, oval) and np.array_equal(rhs, orhs)
    # second sweep without setZero: ADD on top, entry by entry in the same order
    _, _, v2, r2, _ = _assemble(m, kind, num, ed, td, twice=True)
    grp, gcol = O.pattern(num.elemDof, num.size_global)
    o2, or2, _ = O.assemble(kind, num.conn_new, m.coords, num.node_map_get_old, num.elemDof, num.solnApplied, ed, td, grp, gcol,
                            val=oval.copy(), rhs=orhs.copy())
    assert np.array_equal(v2, o2) and np.array_equal(r2, or2)


def test_tiled_solve_and_negative_jacobian(gpu, input_dir, tiled_env):
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    num = D.number(m, S.POISSON_TETRA)
    s = S.SolverB200(0)
    info = D.run_rank(s, m, num, rtol=1e-10)
    assert s.assembly_mode()[0] == 2 and info["reason"] == 2
    u = D.nodal_solution(num, s.get_solution())[0]
    assert np.abs(u - (m.coords ** 2).sum(0)).max() < 2e-7
    s.free()
    e = int(np.flatnonzero((num.elemDof >= 0).all(axis=0))[5])
    m.conn[[0, 1], e] = m.conn[[1, 0], e]
    num = D.number(m, S.POISSON_TETRA)
    s = S.SolverB200(0)
    with pytest.raises(S.PfemError) as ei:
        D.run_rank(s, m, num)
    assert ei.value.status == S.ERR_NEG_JACOBIAN
    s.free()


def test_tiled_request_keeps_row_gather_for_elasticity(gpu, input_dir, tiled_env):
    m = M.read_mesh(os.path.join(input_dir, "cookmembranetria32"))
    num = D.number(m, S.ELASTICITY_TRIA)
    rp, col, val, rhs, mode = _assemble(m, S.ELASTICITY_TRIA, num)
    assert mode[0] == 1                       # several dofs per node: the row-gather kernel stays
    s = S.SolverB200(0)
    D.run_rank(s, m, num, do_solve=False, apply_force_bc=False)
    s.free()
