"""Full-size runs of the BASELINE.json configurations C2, C4 and C5 on one GPU, checked

(1) value by value against the CPU oracle: the single-thread `orc_assemble` (sequential reference element order, ~6 s for
    the 48 M tets of C5) on the GPU-built pattern -- matrix values and RHS must be bit-identical (`array_equal`) on the
    default value pass -- and, for C2, the oracle's CG iteration count at full size;
(2) through size-independent properties:

* pattern size equal to the closed-form count of SURVEY.md section 8 (N + 2 x free-node edges);
* symmetry of the assembled operator (x'Ay == y'Ax to rounding for random x, y);
* exact linearity in the material constant: doubling kx,ky,kz (or E) doubles every matrix entry bit for bit, which
  also cross-checks the unit-coefficient kernel against the general one at full size;
* run-to-run determinism (bit-identical second pass);
* the known answers: u = x^2+y^2+z^2 (C5), the analytic Laplace solution (C2), the beam tip displacement (C4);
* CG iteration counts within +-2 % of the PETSc-semantics counts recorded for these systems.

The same property checks run on the CPU oracle at small sizes in tests/test_oracle_properties.py.
"""
import numpy as np
import pytest

from oracle import pyoracle as O
from pfemfort_b200 import driver as D, mesh as M, solver as S
from properties import nnz_tet_poisson, nnz_tria_poisson, symmetric_to_rounding

pytestmark = pytest.mark.gpu
ITS_TOL = 0.02


def _passes(m, kind, num, ed2):
    """Value pass with the drivers' constants (twice) and with the doubled material constant."""
    s = S.SolverB200(0)
    D.run_rank(s, m, num, do_solve=False, apply_force_bc=False)
    rp, col, v1 = s.get_csr()
    r1 = s.get_rhs()
    s.setZero()
    s.assemble(D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA)
    v1b = s.get_csr(values=True)[2]
    assert np.array_equal(v1, v1b) and np.array_equal(r1, s.get_rhs()), "value pass is not run-to-run deterministic"
    s.setZero()
    s.assemble(ed2, D.DEFAULT_TIMEDATA)
    v2 = s.get_csr(values=True)[2]
    assert np.array_equal(v2, 2.0 * v1), "doubling the material constant must double every entry exactly"
    # value level, full size: the sequential oracle on the same pattern (the pattern itself is pinned by its closed-form size,
    # sortedness and the symmetric probe; at small and mid sizes it is compared entry by entry in test_gpu_parity.py)
    oval, orhs, nbad = O.assemble(kind, num.conn_new, m.coords, None, num.elemDof, num.solnApplied, D.DEFAULT_ELEMDATA[kind],
                                  D.DEFAULT_TIMEDATA, rp, col)
    assert nbad == 0
    assert np.array_equal(v1, oval), "full-size matrix values differ from the sequential oracle"
    assert np.array_equal(r1, orhs), "full-size RHS differs from the sequential oracle"
    return s, rp, col, v1


def test_c5_tet200_poisson_full_size(gpu):
    n = 200
    m = M.gen_tetra(-1, 1, n, -1, 1, n, -1, 1, n)
    kind = S.POISSON_TETRA
    num = D.number(m, kind)
    assert m.nElem == 48_000_000 and num.size_global == 7_880_599
    s, rp, col, val = _passes(m, kind, num, [2.0, 2.0, 2.0])
    assert col.size == nnz_tet_poisson(n) == 117_260_947       # N + 2 x (free-node edges of the 6-tet split)
    assert np.all(np.diff(rp) <= 15) and rp[-1] == col.size
    assert symmetric_to_rounding(rp, col, val)
    del rp, col, val
    s.free()
    s = S.SolverB200(0)
    info = D.run_rank(s, m, num, rtol=1e-10)
    assert info["reason"] == 2 and abs(info["its"] - 699) <= ITS_TOL * 699
    u = D.nodal_solution(num, s.get_solution())[0]
    assert np.abs(u - (m.coords ** 2).sum(0)).max() < 2e-7      # nodally exact up to the float32-rounded boundary data
    s.free()


def test_c2_tria1000_poisson_full_size(gpu):
    m = M.gen_tria_poisson(1000)
    kind = S.POISSON_TRIA
    num = D.number(m, kind)
    assert m.nElem == 2_000_000 and num.size_global == 998_001
    s, rp, col, val = _passes(m, kind, num, [2.0, 2.0, 1.0])
    assert col.size == nnz_tria_poisson(1000) == 6_978_017      # explicit zeros of the hypotenuse couplings included
    assert symmetric_to_rounding(rp, col, val)
    s.free()
    s = S.SolverB200(0)
    info = D.run_rank(s, m, num, rtol=1e-10)
    assert info["reason"] == 2 and abs(info["its"] - 1442) <= ITS_TOL * 1442
    _, oits, oreason, _ = O.cg_jacobi(rp, col, val, s.get_rhs(), rtol=1e-10, threads=O.num_threads())
    assert oreason == 2 and abs(info["its"] - oits) <= ITS_TOL * oits, (info["its"], oits)
    u = D.nodal_solution(num, s.get_solution())[0]
    assert np.abs(u - M.exact_poisson_tria(m.coords[0], m.coords[1])).max() < 1e-6
    s.free()


def test_c4_beam_elasticity_full_size(gpu):
    m = M.gen_tetra(-0.5, 0.5, 50, 0.0, 6.0, 300, -0.5, 0.5, 50, dbc="clamp_y0", ndof=3)
    kind = S.ELASTICITY_TETRA
    num = D.number(m, kind)
    assert m.nElem == 4_500_000 and num.size_global == 2_340_900
    ed = list(D.DEFAULT_ELEMDATA[kind])
    ed2 = [2.0 * ed[0]] + ed[1:]
    s, rp, col, val = _passes(m, kind, num, ed2)
    assert col.size == 102_964_482
    assert symmetric_to_rounding(rp, col, val)
    del rp, col, val
    s.free()
    s = S.SolverB200(0)
    info = D.run_rank(s, m, num, rtol=1e-10, max_it=200000)
    # The iteration count of this ill-conditioned system (Jacobi-CG, rtol 1e-10, ~6000+ iterations) depends on the summation
    # ORDER of the dot products, far beyond +-2 %: the CPU oracle needs 8175 iterations with sequential sums (1 thread) and 6402
    # with 4-thread OpenMP reductions (same matrix, same RHS, bit for bit; /tmp run of 2026-10-17 recorded in DESIGN.md 6), the
    # GPU's tree sums 5891 (r01 reduction order) / 6027 (r02 order).  All converge (reason 2) to the same displacement field.
    # So: converged, and never more iterations than the sequential-order reference (+2 %).  C5 and C2, whose counts are stable
    # under the summation order, keep the +-2 % check against the oracle.
    assert info["reason"] == 2 and 5400 <= info["its"] <= 8175 * (1 + ITS_TOL), info["its"]
    u = D.nodal_solution(num, s.get_solution())
    assert abs(np.sqrt((u ** 2).sum(0)).max() - 0.8221) < 5e-3  # README image 0.82, beam theory 0.808
    s.free()
