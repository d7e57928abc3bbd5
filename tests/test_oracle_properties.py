"""The size-independent properties used by the full-size GPU tests (tests/test_gpu_zz_fullsize.py), checked here on the
CPU oracle at sizes it assembles in seconds: the checks themselves are validated before they are trusted at 48 M elements."""
import numpy as np

from oracle import pyoracle as O
from pfemfort_b200 import driver as D, mesh as M, solver as S
from properties import nnz_tet_poisson, nnz_tria_poisson, symmetric_to_rounding


def _system(m, kind, elemData=None):
    num = D.number(m, kind)
    rp, col = O.pattern(num.elemDof, num.size_global)
    val, rhs, nbad = O.assemble(kind, num.conn_new, m.coords, num.node_map_get_old, num.elemDof, num.solnApplied,
                                D.DEFAULT_ELEMDATA[kind] if elemData is None else elemData, D.DEFAULT_TIMEDATA, rp, col)
    assert nbad == 0
    return num, rp, col, val, rhs


def test_pattern_size_formulas():
    for n in (3, 6, 11):
        m = M.gen_tetra(-1, 1, n, -1, 1, n, -1, 1, n)
        num, rp, col, _, _ = _system(m, S.POISSON_TETRA)
        assert num.size_global == (n - 1) ** 3 and col.size == nnz_tet_poisson(n)
    for n in (3, 20, 57):
        m = M.gen_tria_poisson(n)
        num, rp, col, _, _ = _system(m, S.POISSON_TRIA)
        assert col.size == nnz_tria_poisson(n)
    assert nnz_tet_poisson(200) == 117_260_947 and nnz_tria_poisson(1000) == 6_978_017
    assert nnz_tet_poisson(10) == 9097 and nnz_tria_poisson(20) == 2377                     # the bundled tet10 / tria20x20


def test_doubling_the_material_constant_is_exact():
    m = M.gen_tetra(-1, 1, 7, -1, 1, 6, -1, 1, 5)
    _, rp, col, v1, _ = _system(m, S.POISSON_TETRA)
    _, _, _, v2, _ = _system(m, S.POISSON_TETRA, [2.0, 2.0, 2.0])
    assert np.array_equal(v2, 2.0 * v1)
    m = M.gen_tria_poisson(17)
    _, _, _, v1, _ = _system(m, S.POISSON_TRIA)
    _, _, _, v2, _ = _system(m, S.POISSON_TRIA, [2.0, 2.0, 1.0])
    assert np.array_equal(v2, 2.0 * v1)
    m = M.gen_tetra(-0.5, 0.5, 3, 0.0, 6.0, 9, -0.5, 0.5, 3, dbc="clamp_y0", ndof=3)
    ed = list(D.DEFAULT_ELEMDATA[S.ELASTICITY_TETRA])
    _, _, _, v1, r1 = _system(m, S.ELASTICITY_TETRA)
    _, _, _, v2, r2 = _system(m, S.ELASTICITY_TETRA, [2.0 * ed[0]] + ed[1:])
    assert np.array_equal(v2, 2.0 * v1) and np.array_equal(r1, r2)      # the body-force load does not depend on E


def test_symmetry_probe_detects_asymmetry():
    m = M.gen_tetra(-1, 1, 6, -1, 1, 6, -1, 1, 6)
    _, rp, col, val, _ = _system(m, S.POISSON_TETRA)
    assert symmetric_to_rounding(rp, col, val)
    # K(i,j) and K(j,i) are evaluated with different operand orders: symmetric to rounding, not always bit for bit
    bad = val.copy()
    k = int(np.flatnonzero(col[rp[5]:rp[6]] != 5)[0]) + rp[5]
    bad[k] *= 1.0 + 1e-6
    assert not symmetric_to_rounding(rp, col, bad)
    m = M.gen_tetra(-0.5, 0.5, 3, 0.0, 6.0, 9, -0.5, 0.5, 3, dbc="clamp_y0", ndof=3)
    _, rp, col, val, _ = _system(m, S.ELASTICITY_TETRA)
    assert symmetric_to_rounding(rp, col, val)
