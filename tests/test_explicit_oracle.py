"""CPU tests of the oracle's explicit-dynamics restatement (ResidualElasticityLinear{Tria,Tetra}, MassMatrixLinear{Tria,
Tetra}, the lumped-mass loop and the central-difference time loop of triaelasticityexplicit.F): closed forms, physical
identities, and an independent vectorised numpy implementation of the same scheme."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from pfemfort_b200 import driver as D, explicit as X, mesh as M, solver as S

THIRD_F = float(np.float32(1.0) / np.float32(3.0))
SIXTH_F = float(np.float32(1.0) / np.float32(6.0))
ED2 = [200.0, 0.3, 10.0, 1.0, 0.0]            # triaelasticityexplicit.F:870-875
ED3 = [200.0, 0.3, 10.0, 0.5, -0.25, 1.0]
TD = [0.0, 1.0, 0.0]


def test_mass_routines_closed_form():
    x, y = np.array([0.0, 2.0, 0.5]), np.array([0.0, 0.25, 1.5])
    area = 0.5 * ((x[1] - x[0]) * (y[2] - y[0]) - (y[1] - y[0]) * (x[2] - x[0]))
    Ml, rc = O.mass_matrix(O.ELASTICITY_TRIA, x, y, None, ED2)
    assert rc == 0
    N = np.array([1.0 - THIRD_F - THIRD_F, THIRD_F, THIRD_F])
    want = np.repeat(ED2[2] * area * N * N.sum(), 2)
    assert np.allclose(Ml, want, rtol=1e-14) and abs(Ml[0::2].sum() - ED2[2] * area) < 1e-6 * ED2[2] * area
    xt, yt, zt = np.array([1.0, 0.0, 0.0, 0.0]), np.array([0.0, 1.0, 0.0, 0.0]), np.array([0.0, 0.0, 0.0, 1.0])   # node 3 = origin
    Mt, rc = O.mass_matrix(O.ELASTICITY_TETRA, xt, yt, zt, ED3)
    assert rc == 0 and np.allclose(Mt, ED3[2] * SIXTH_F * 0.25, rtol=1e-14)       # volume 1/6 (float literal), quarter per node
    _, rc = O.mass_matrix(O.ELASTICITY_TRIA, x[::-1].copy(), y[::-1].copy(), None, ED2)
    assert rc == 1                                                              # clockwise triangle: the reference STOPs


@pytest.mark.parametrize("kind", [O.ELASTICITY_TRIA, O.ELASTICITY_TETRA])
def test_residual_identities(kind):
    rng = np.random.default_rng(5 + kind)
    if kind == O.ELASTICITY_TRIA:
        x, y, z = np.array([0.1, 1.9, 0.4]), np.array([-0.2, 0.3, 1.6]), None
        ed, npe, ndof = ED2, 3, 2
    else:
        x, y, z = np.array([1.1, 0.1, -0.1, 0.2]), np.array([0.1, 0.9, 0.0, 0.1]), np.array([0.0, 0.2, -0.1, 1.3])
        ed, npe, ndof = ED3, 4, 3
    xyz = np.stack([x, y] + ([z] if z is not None else []))
    # zero displacement or rigid translation: body force only, sum = density-weighted volume x b (2-D) / volume x b (3-D)
    F0, rc = O.residual_elasticity(kind, x, y, z, ed, TD, np.zeros(npe * ndof))
    Ft, _ = O.residual_elasticity(kind, x, y, z, ed, TD, np.tile(rng.standard_normal(ndof), npe))
    assert rc == 0 and np.allclose(F0, Ft, atol=1e-12)
    Ml, _ = O.mass_matrix(kind, x, y, z, ed)
    b = np.array(ed[3:3 + ndof])
    scale = 1.0 if kind == O.ELASTICITY_TRIA else 1.0 / ed[2]                   # the 3-D residual has no density factor (:712)
    assert np.allclose(F0.reshape(npe, ndof).sum(0), Ml.reshape(npe, ndof).sum(0)[0] * scale * b, rtol=1e-6)
    # linear displacement field u = G x: constant strain; internal forces are self-equilibrated and match -V B^T sigma
    G = rng.standard_normal((ndof, ndof))
    u = (G @ xyz).T.ravel()
    F, _ = O.residual_elasticity(kind, x, y, z, ed, TD, u)
    Fint = (F - F0).reshape(npe, ndof)
    assert np.abs(Fint.sum(0)).max() < 1e-10 * np.abs(Fint).max()
    E, nu = ed[0], ed[1]
    b1 = E / ((1 + nu) * (1 - 2 * nu))
    eps = 0.5 * (G + G.T)
    # the reference multiplies the TENSOR shear strain 0.5 (g_ij + g_ji) by the engineering shear modulus b1 (1-2nu)/2
    # (elasticity2D.F:259,206; elasticity3D.F:685-687,626-628): half the textbook shear stress.  Replicated, not fixed.
    shear_half = np.where(np.eye(ndof) > 0, 1.0, 0.5)
    sig = b1 * (1 - 2 * nu) * eps * shear_half + b1 * nu * np.trace(eps) * np.eye(ndof)
    # nodal forces of a constant stress: -V sigma grad N_i ; check through the virtual work of a second linear field
    H = rng.standard_normal((ndof, ndof))
    v = (H @ xyz).T
    vol = abs(np.linalg.det(np.stack([xyz[:, k] - xyz[:, 0] for k in range(1, npe)]))) / (2 if ndof == 2 else 6)
    if kind == O.ELASTICITY_TETRA:
        vol *= SIXTH_F * 6                                                       # the routine's float(1/6) weight
    assert np.isclose((Fint * v).sum(), -vol * (sig * (0.5 * (H + H.T))).sum(), rtol=1e-9)


def _numpy_explicit(kind, conn, coords, free_slots, ed, dt, nsteps):
    """Independent vectorised implementation of the scheme (no attention to evaluation order)."""
    npe, ndof, ndim = S.KIND_DIMS[kind]
    c = conn - 1
    nE, nN = c.shape[1], coords.shape[1]
    P = coords[:, c]                                         # [ndim, npe, nE]
    if ndim == 2:
        Nq = np.array([1 - 2 * THIRD_F, THIRD_F, THIRD_F]); w = 0.5
        dNxi = np.array([[-1.0, 1, 0], [-1, 0, 1]])
    else:
        Nq = np.array([0.25, 0.25, 0.25, 0.25]); w = SIXTH_F
        dNxi = np.array([[1.0, 0, -1, 0], [0, 1, -1, 0], [0, 0, -1, 1]])
    J = np.einsum("kn,cne->kce", dNxi, P)                    # J[k,c] = sum_n dN_n/dxi_k x_c(n)
    detJ = np.linalg.det(J.transpose(2, 0, 1))
    Jinv = np.linalg.inv(J.transpose(2, 0, 1))               # [nE, c, k]
    dN = np.einsum("eck,kn->ecn", Jinv, dNxi)                # [nE, c, n]
    E, nu, dens = ed[:3]
    b = np.array(ed[3:3 + ndof])
    b1 = E / ((1 + nu) * (1 - 2 * nu))
    dvol = w * detJ
    mnode = (dens * dvol)[:, None] * Nq[None, :] * Nq.sum()
    Mg = np.zeros((nN, ndof))
    for n in range(npe):
        np.add.at(Mg, c[n], mnode[:, n, None])
    Mg = Mg.ravel()
    fs = free_slots - 1
    d1 = np.zeros(nN * ndof); d2 = np.zeros(nN * ndof)
    for _ in range(nsteps):
        U = d1.reshape(nN, ndof)[c]                          # [npe, nE, ndof]
        G = np.einsum("ned,ecn->edc", U, dN)                 # grad[d][c]
        eps = 0.5 * (G + G.transpose(0, 2, 1))
        shear_half = np.where(np.eye(ndof) > 0, 1.0, 0.5)                      # the reference's half shear stress (see above)
        sig = b1 * (1 - 2 * nu) * eps * shear_half + (b1 * nu) * np.trace(eps, axis1=1, axis2=2)[:, None, None] * np.eye(ndof)
        Fint = -np.einsum("e,edc,ecn->ned", dvol, sig, dN)
        bodyscale = dens if ndim == 2 else 1.0
        Fb = (bodyscale * dvol)[None, :, None] * Nq[:, None, None] * b[None, None, :]
        rhs = np.zeros((nN, ndof))
        for n in range(npe):
            np.add.at(rhs, c[n], Fint[n] + Fb[n])
        rhs = rhs.ravel()
        new = d1.copy()
        new[fs] = (dt * dt) * (rhs[fs] + Mg[fs] / (dt * dt) * (2 * d1[fs] - d2[fs])) / Mg[fs]
        d2, d1 = d1, new
    return Mg, d1, d2


@pytest.mark.parametrize("name,kind,swap,ed,dt", [("cookmembranetria32", S.ELASTICITY_TRIA, False, ED2, 2e-4),
                                                  ("beam3Dtet6366", S.ELASTICITY_TETRA, True, ED3, 1e-3)])
def test_time_loop_against_numpy_and_restart(input_dir, name, kind, swap, ed, dt):
    m = M.read_mesh(os.path.join(input_dir, name), swap_34=swap)
    num = D.number(m, kind)
    fs = X.free_slots(num)
    assert fs.size == num.size_global and np.all(np.diff(fs) > 0)
    Mg, nbad = O.explicit_lumped_mass(kind, num.conn_new, m.coords, ed)
    assert nbad == 0
    st = O.explicit_advance(kind, num.conn_new, m.coords, fs, ed, TD, dt, 12, Mg)
    Mn, d1, d2 = _numpy_explicit(kind, num.conn_new, m.coords, fs, ed, dt, 12)
    assert np.allclose(Mg, Mn, rtol=1e-13)
    scale = np.abs(d1).max()
    assert scale > 0 and np.abs(st["disp"] - d1).max() < 1e-10 * scale and np.abs(st["dispPrev2"] - d2).max() < 1e-10 * scale
    assert np.array_equal(st["disp"], st["dispPrev"])                           # loop invariant (:1118-1121)
    fixed = np.setdiff1d(np.arange(Mg.size), fs - 1)
    assert fixed.size > 0 and np.all(st["disp"][fixed] == 0.0)                  # Dirichlet dofs are never touched
    # restart: 5 + 7 steps from the saved state == 12 steps
    a = O.explicit_advance(kind, num.conn_new, m.coords, fs, ed, TD, dt, 5, Mg)
    a = O.explicit_advance(kind, num.conn_new, m.coords, fs, ed, TD, dt, 7, Mg, state=a)
    for k in ("disp", "dispPrev2", "velo", "acce"):
        assert np.array_equal(a[k], st[k]), k
    # total mass = density x domain measure (area 1440 for Cook's membrane, volume 6 for the beam; float-literal weights)
    total = Mg.reshape(-1, S.KIND_DIMS[kind][1])[:, 0].sum()
    assert np.isclose(total, ed[2] * (1440.0 if kind == S.ELASTICITY_TRIA else 6.0), rtol=1e-6)
