"""GPU parity of the explicit-dynamics path (csrc/explicit.cu) against the sequential CPU oracle: single-element routines,
lumped mass, and the central-difference time loop -- bit for bit (the node gather adds the element residuals in the
sequential loop's order; the TU is compiled without FMA contraction)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from pfemfort_b200 import driver as D, explicit as X, mesh as M, solver as S

pytestmark = pytest.mark.gpu
ED2 = [200.0, 0.3, 10.0, 1.0, 0.0]            # triaelasticityexplicit.F:870-875
ED3 = [200.0, 0.3, 10.0, 0.5, -0.25, 1.0]
TD = [0.0, 1.0, 0.0]


@pytest.mark.parametrize("kind", [S.ELASTICITY_TRIA, S.ELASTICITY_TETRA])
def test_single_element_routines_bit_exact(gpu, kind):
    rng = np.random.default_rng(17 + kind)
    npe, ndof, ndim = S.KIND_DIMS[kind]
    ed = ED2 if ndim == 2 else ED3
    base = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]) if ndim == 2 else \
        np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]])
    for _ in range(8):
        xyz = base + 0.2 * rng.standard_normal((ndim, npe)) + 3.0 * rng.standard_normal((ndim, 1))
        u = rng.standard_normal(npe * ndof)
        z = xyz[2] if ndim == 3 else None
        Fo, rc = O.residual_elasticity(kind, xyz[0], xyz[1], z, ed, TD, u)
        Mo, rc2 = O.mass_matrix(kind, xyz[0], xyz[1], z, ed)
        assert rc == 0 and rc2 == 0
        assert np.array_equal(X.residual_elasticity(kind, xyz[0], xyz[1], z, ed, TD, u), Fo)
        assert np.array_equal(X.mass_matrix(kind, xyz[0], xyz[1], z, ed), Mo)
    # inverted element: the reference STOPs, the library reports it
    flip = [1, 0, 2] if ndim == 2 else [1, 0, 2, 3]
    with pytest.raises(S.PfemError) as ei:
        X.mass_matrix(kind, base[0][flip], base[1][flip], base[2][flip] if ndim == 3 else None, ed)
    assert ei.value.status == S.ERR_NEG_JACOBIAN


@pytest.mark.parametrize("name,kind,swap,ed,dt", [("cookmembranetria32", S.ELASTICITY_TRIA, False, ED2, 2e-4),
                                                  ("beam3Dtet6366", S.ELASTICITY_TETRA, True, ED3, 1e-3)])
def test_lumped_mass_and_time_loop_bit_exact(gpu, input_dir, name, kind, swap, ed, dt):
    m = M.read_mesh(os.path.join(input_dir, name), swap_34=swap)
    num = D.number(m, kind)
    fs = X.free_slots(num)
    ex = X.ExplicitB200(0)
    ex.set_mesh(kind, num.conn_new, m.coords)
    ex.set_free_dofs(fs)
    ex.lumped_mass(ed)
    Mg, nbad = O.explicit_lumped_mass(kind, num.conn_new, m.coords, ed)
    assert nbad == 0 and np.array_equal(ex.get_state()["mass"], Mg)
    # load phase 1 (body force on), then phase 2 (off): the driver's switch at timeNow <= 0.1 (triaelasticityexplicit.F:976-979)
    ed_off = list(ed)
    ed_off[3] = 0.0
    ex.advance(25, dt, ed, TD)
    st = O.explicit_advance(kind, num.conn_new, m.coords, fs, ed, TD, dt, 25, Mg)
    g = ex.get_state()
    for k in ("disp", "dispPrev2", "velo", "acce"):
        assert np.array_equal(g[k], st[k]), k
    saved = (g["disp"].copy(), g["dispPrev2"].copy())
    ex.advance(15, dt, ed_off, TD)
    st = O.explicit_advance(kind, num.conn_new, m.coords, fs, ed_off, TD, dt, 15, Mg, state=st)
    g2 = ex.get_state()
    for k in ("disp", "dispPrev2", "velo", "acce"):
        assert np.array_equal(g2[k], st[k]), k
    assert np.abs(g2["disp"]).max() > 0 and ex.info()["steps"] == 40 and ex.info()["launches"] >= 40
    # checkpoint / resume: restart from the saved pair, bit-identical continuation
    ex.set_state(*saved)
    ex.advance(15, dt, ed_off, TD)
    g3 = ex.get_state()
    for k in ("disp", "dispPrev2", "velo", "acce"):
        assert np.array_equal(g3[k], g2[k]), k
    ex.free()


def test_explicit_state_machine_and_errors(gpu, input_dir):
    m = M.read_mesh(os.path.join(input_dir, "cookmembranetria32"))
    num = D.number(m, S.ELASTICITY_TRIA)
    ex = X.ExplicitB200(0)
    with pytest.raises(S.PfemError) as ei:
        ex.lumped_mass(ED2)
    assert ei.value.status == S.ERR_STATE
    with pytest.raises(S.PfemError):
        ex.set_mesh(S.POISSON_TRIA, num.conn_new, m.coords)        # elasticity kinds only
    ex.set_mesh(S.ELASTICITY_TRIA, num.conn_new, m.coords)
    with pytest.raises(S.PfemError) as ei:
        ex.advance(1, 1e-4, ED2)                                   # no mass, no free dofs yet
    assert ei.value.status == S.ERR_STATE
    with pytest.raises(S.PfemError) as ei:
        ex.set_free_dofs(np.array([0, 5], np.int32))               # slot 0 is out of range (1-based)
    assert ei.value.status == S.ERR_NUMBERING
    bad = num.conn_new.copy()
    bad[[0, 1], 7] = bad[[1, 0], 7]                                # one inverted triangle
    ex.set_mesh(S.ELASTICITY_TRIA, bad, m.coords)
    with pytest.raises(S.PfemError) as ei:
        ex.lumped_mass(ED2)
    assert ei.value.status == S.ERR_NEG_JACOBIAN
    ex.free()


def test_explicit_mid_size_generated_mesh(gpu):
    """24^3 x 6 tets clamped at y = 0 (C4's recipe at small size), 10 steps: bit-identical state, run-to-run deterministic."""
    m = M.gen_tetra(-0.5, 0.5, 8, 0.0, 6.0, 48, -0.5, 0.5, 8, dbc="clamp_y0", ndof=3)
    kind = S.ELASTICITY_TETRA
    num = D.number(m, kind)
    fs = X.free_slots(num)
    outs = []
    for _ in range(2):
        ex = X.ExplicitB200(0)
        ex.set_mesh(kind, num.conn_new, m.coords)
        ex.set_free_dofs(fs)
        ex.lumped_mass(ED3)
        ex.advance(10, 1e-3, ED3, TD)
        outs.append(ex.get_state())
        ex.free()
    Mg, _ = O.explicit_lumped_mass(kind, num.conn_new, m.coords, ED3)
    st = O.explicit_advance(kind, num.conn_new, m.coords, fs, ED3, TD, 1e-3, 10, Mg)
    for k in ("disp", "dispPrev2", "velo", "acce"):
        assert np.array_equal(outs[0][k], st[k]) and np.array_equal(outs[0][k], outs[1][k]), k
