"""Host-side mirror of the explicit-dynamics part of the C ABI (include/pfem_b200.h, csrc/explicit.cu): the element
routines ResidualElasticityLinear{Tria,Tetra} / MassMatrixLinear{Tria,Tetra} and the central-difference time loop of
triaelasticityexplicit.F:881-921, 972-1121.  No CPU fallback: everything runs on the GPU through libpfemb200.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import solver as S

_f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)   # noqa: E731
_i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)     # noqa: E731


# the explicit PROGRAM's own constants (triaelasticityexplicit.F:870-875, 957-958): single-precision literals widened
DRIVER_ELEMDATA_TRIA = [200.0, float(np.float32(0.3)), 10.0, 1.0, 0.0, 0.0]   # E, nu, density, body force x / y
DRIVER_TIMEDATA = [0.0, 1.0, 0.0]
DRIVER_DT = float(np.float32(0.0002))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _chk(rc):
    S._chk(rc)


def residual_elasticity(kind, x, y, z, elemData, timeData, dispC, veloC=None):
    """One element: Flocal (6 or 12 values).  Raises PfemError(ERR_NEG_JACOBIAN) where the reference STOPs."""
    lib = S.load_library()
    npe, ndof, ndim = S.KIND_DIMS[kind]
    F = np.zeros(npe * ndof)
    x, y, ed, td, dc = _f64(x), _f64(y), _f64(elemData), _f64(timeData), _f64(dispC)
    vc = _f64(veloC) if veloC is not None else np.zeros(npe * ndof)
    if kind == S.ELASTICITY_TRIA:
        _chk(lib.pfem_residual_elasticity_linear_tria(_d(x), _d(y), _d(ed), _d(td), _d(dc), _d(vc), _d(F)))
    else:
        zz = _f64(z)
        _chk(lib.pfem_residual_elasticity_linear_tetra(_d(x), _d(y), _d(zz), _d(ed), _d(td), _d(dc), _d(vc), _d(F)))
    return F


def mass_matrix(kind, x, y, z, elemData):
    lib = S.load_library()
    npe, ndof, ndim = S.KIND_DIMS[kind]
    Ml = np.zeros(npe * ndof)
    x, y, ed = _f64(x), _f64(y), _f64(elemData)
    if kind == S.ELASTICITY_TRIA:
        _chk(lib.pfem_mass_matrix_linear_tria(_d(x), _d(y), _d(ed), _d(Ml)))
    else:
        zz = _f64(z)
        _chk(lib.pfem_mass_matrix_linear_tetra(_d(x), _d(y), _d(zz), _d(ed), _d(Ml)))
    return Ml


class ExplicitB200:
    """The plain-array state of an explicit driver (globalM, disp, dispPrev, dispPrev2, velo, acce) resident on one GPU."""

    def __init__(self, device: int = 0):
        self._lib = S.load_library()
        self._h = C.c_void_p()
        _chk(self._lib.pfem_explicit_create(C.byref(self._h), device))
        self.nd = 0

    def free(self):
        if self._h:
            self._lib.pfem_explicit_free(self._h)
            self._h = C.c_void_p()

    def set_mesh(self, kind, conn, coords):
        conn, coords = _i32(conn), _f64(coords)
        self.nd = coords.shape[1] * S.KIND_DIMS[kind][1]
        _chk(self._lib.pfem_explicit_set_mesh(self._h, kind, conn.shape[1], conn.ctypes.data_as(C.POINTER(C.c_int)), coords.shape[1],
                                              _d(coords)))

    def set_free_dofs(self, assyForSoln):
        a = _i32(assyForSoln)
        _chk(self._lib.pfem_explicit_set_free_dofs(self._h, a.size, a.ctypes.data_as(C.POINTER(C.c_int))))

    def lumped_mass(self, elemData):
        _chk(self._lib.pfem_explicit_lumped_mass(self._h, _d(_f64(elemData))))

    def advance(self, nsteps, dt, elemData, timeData=(0.0, 1.0, 0.0)):
        _chk(self._lib.pfem_explicit_advance(self._h, int(nsteps), C.c_double(dt), _d(_f64(elemData)), _d(_f64(timeData))))

    def get_state(self):
        out = {k: np.zeros(self.nd) for k in ("disp", "dispPrev2", "velo", "acce", "mass")}
        _chk(self._lib.pfem_explicit_get_state(self._h, _d(out["disp"]), _d(out["dispPrev2"]), _d(out["velo"]), _d(out["acce"]),
                                               _d(out["mass"])))
        return out

    def set_state(self, disp, dispPrev2):
        _chk(self._lib.pfem_explicit_set_state(self._h, _d(_f64(disp)), _d(_f64(dispPrev2))))

    def info(self):
        st, la, t = C.c_longlong(), C.c_longlong(), C.c_double()
        _chk(self._lib.pfem_explicit_get_info(self._h, C.byref(st), C.byref(la), C.byref(t)))
        return dict(steps=st.value, launches=la.value, t_advance=t.value)


def free_slots(num) -> np.ndarray:
    """assyForSoln of the drivers (tetrapoissonparallelimpl1.F:722-734): 1-based node slot (newnode-1)*ndof + j of every
    free dof, in dof order."""
    ndof, nNode = num.NodeDofArrayNew.shape
    nda = num.NodeDofArrayNew.T.ravel()          # slot order: node-major
    slots = np.flatnonzero(nda > 0)
    out = np.zeros(num.size_global, np.int32)
    out[nda[slots] - 1] = slots + 1
    return out
