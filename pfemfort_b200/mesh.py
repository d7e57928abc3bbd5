"""Host-side mesh input: the reference's text formats and the recipes of its generators.

Mirrors the readers of the *parallelimpl1 drivers (tetrapoissonparallelimpl1.F:216-355: list-directed
``id x y [z]`` / ``id n1..n`` / ``node dof value`` rows, 1-based ids) and the structured generators whose
recipes define the synthetic benchmark inputs (genTetra.cpp:187-216 nodes, :250-334 six tets per cell,
:355-525 Dirichlet rows; the triangle rule of tria20x20-elems).  Host code only (numpy); nothing here is timed.

Array conventions are the drivers': ``coords[c, n]`` and ``conn[i, e]`` are C-contiguous ``(ncol, nrow)``
arrays, i.e. the Fortran column-major ``coords(nNode, ndim)`` / ``elemNodeConn(nElem, npElem)`` in memory.
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Mesh:
    coords: np.ndarray                 # float64 [ndim, nNode]  (OLD numbering)
    conn: np.ndarray                   # int32   [npElem, nElem] 1-based node ids
    dbc_node: np.ndarray               # int32 [nDBC] 1-based
    dbc_dof: np.ndarray                # int32 [nDBC] 1-based
    dbc_val: np.ndarray                # float64 [nDBC]
    fbc_node: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    fbc_dof: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    fbc_val: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float64))
    name: str = ""

    @property
    def ndim(self) -> int:
        return self.coords.shape[0]

    @property
    def nNode(self) -> int:
        return self.coords.shape[1]

    @property
    def npElem(self) -> int:
        return self.conn.shape[0]

    @property
    def nElem(self) -> int:
        return self.conn.shape[1]


def _open(path: str):
    return gzip.open(path, "rt") if path.endswith(".gz") else open(path, "rt")


def _load_table(path: str, ncols: int | None = None) -> np.ndarray:
    """One text table ``id v1 v2 ...``.  Plain files go through the one-pass C parser of the library
    (csrc/host_meshio.cu) when the column count is known; gzipped fixtures through numpy."""
    if ncols is not None and not path.endswith(".gz"):
        import ctypes as C

        from . import solver as S
        lib = S.load_library()
        lib.pfem_host_read_table.restype = C.c_longlong
        lib.pfem_host_read_table.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_double), C.c_longlong]
        n = lib.pfem_host_read_table(path.encode(), ncols, None, 0)
        if n < 0:
            raise FileNotFoundError(path)
        out = np.zeros((ncols, max(n, 1)))
        # the count pass tokenises, the fill pass converts: rows with non-numeric tokens drop out in the second pass only
        n = min(n, lib.pfem_host_read_table(path.encode(), ncols, out.ctypes.data_as(C.POINTER(C.c_double)), n))
        return np.ascontiguousarray(out[:, :n].T)
    with _open(path) as f:
        return np.loadtxt(f, dtype=np.float64, ndmin=2)


def _find(prefix: str, kind: str) -> str | None:
    for ext in (".dat.gz", ".dat"):
        p = f"{prefix}-{kind}{ext}"
        if os.path.exists(p):
            return p
    return None


def read_mesh(prefix: str, swap_34: bool = False) -> Mesh:
    """Read ``<prefix>-nodes|elems|DirichBC|ForceBC.dat[.gz]`` (the drivers' positional arguments).

    swap_34 exchanges local nodes 3 and 4 of every tetrahedron: the bundled beam3Dtet6366 file has
    negative Jacobians under the reference's own basis functions.
    """
    nodes = _load_table(_find(prefix, "nodes"))
    elems = _load_table(_find(prefix, "elems"))
    dbc = _load_table(_find(prefix, "DirichBC"))
    coords = np.ascontiguousarray(nodes[:, 1:].T)
    conn = np.ascontiguousarray(elems[:, 1:].T.astype(np.int32))
    if swap_34:
        conn[[2, 3]] = conn[[3, 2]]
    m = Mesh(coords, conn, dbc[:, 0].astype(np.int32), dbc[:, 1].astype(np.int32), dbc[:, 2].copy(),
             name=os.path.basename(prefix))
    fpath = _find(prefix, "ForceBC")
    if fpath:
        fbc = _load_table(fpath)
        m.fbc_node = fbc[:, 0].astype(np.int32)
        m.fbc_dof = fbc[:, 1].astype(np.int32)
        m.fbc_val = fbc[:, 2].copy()
    return m


def _text_round(v: np.ndarray) -> np.ndarray:
    """Round trip through the generators' ``fixed, precision(8)`` text output (genTetra.cpp:187-189)."""
    return np.array([float(f"{x:.8f}") for x in np.asarray(v, dtype=np.float64).ravel()]).reshape(np.shape(v))


def _accumulate(x0: float, x1: float, n: int) -> np.ndarray:
    """Repeated ``xx += dx`` in double (genTetra.cpp:199-216): sequential accumulation, not linspace."""
    dx = (x1 - x0) / n
    out = np.empty(n + 1, dtype=np.float64)
    xx = x0
    for i in range(n + 1):
        out[i] = xx
        xx += dx
    return out


def tria_elements(nx: int, ny: int) -> np.ndarray:
    """The triangle rule of the bundled tria meshes: cell (i,j), n = j*(nx+1)+i+1 ->
    (n, n+1, n+nx+2), (n, n+nx+2, n+nx+1)."""
    j, i = np.meshgrid(np.arange(ny, dtype=np.int64), np.arange(nx, dtype=np.int64), indexing="ij")
    n = (j * (nx + 1) + i + 1).ravel()
    conn = np.empty((3, 2 * nx * ny), dtype=np.int32)
    conn[:, 0::2] = np.stack([n, n + 1, n + nx + 2])
    conn[:, 1::2] = np.stack([n, n + nx + 2, n + nx + 1])
    return conn


def gen_tria_poisson(n: int) -> Mesh:
    """Unit-square n x n tria Poisson case of tria20x20 / tria1000x1000: u(x,0) = sin(pi x), 0 on the other
    sides; nodes written with 8 decimals.  Boundary rows are listed in the order of the bundled files:
    bottom, then right (excluding corner already listed), top, left."""
    xs = _text_round(np.arange(n + 1) / n)
    X, Y = np.meshgrid(xs, xs, indexing="xy")       # row j = y, column i = x ; node id = j*(n+1)+i+1
    coords = np.ascontiguousarray(np.stack([X.ravel(), Y.ravel()]))
    conn = tria_elements(n, n)
    ids = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    bottom = ids[0, :]
    top = ids[n, :]
    left = ids[1:n, 0]
    right = ids[1:n, n]
    nodes = np.concatenate([bottom, top, left, right]) + 1
    vals = np.zeros(nodes.size)
    vals[: n + 1] = _text_round(np.sin(np.pi * coords[0, bottom]))
    return Mesh(coords, conn, nodes.astype(np.int32), np.ones(nodes.size, np.int32), vals, name=f"tria{n}x{n}")


def gen_tetra(x0, x1, nEx, y0, y1, nEy, z0, z1, nEz, dbc: str = "poisson", ndof: int = 1) -> Mesh:
    """genTetra.cpp recipe: structured box, 6 tets per cell.

    Coordinates are accumulated doubles rounded to 8 decimals (the text round trip a driver sees).
    dbc = "poisson": all six faces, value x^2+y^2+z^2 evaluated on float32-rounded coordinates (vtkPoints
    stores floats; genTetra.cpp:497-525), printed with 8 decimals.
    dbc = "clamp_y0": every dof fixed to 0 on the y = y0 face (genTetranovtk.cpp:240 variant, the beam).
    """
    ax, ay, az = _accumulate(x0, x1, nEx), _accumulate(y0, y1, nEy), _accumulate(z0, z1, nEz)
    nNx, nNy, nNz = nEx + 1, nEy + 1, nEz + 1
    rx, ry, rz = _text_round(ax), _text_round(ay), _text_round(az)
    Z, Y, X = np.meshgrid(rz, ry, rx, indexing="ij")            # node index = kk*nNx*nNy + jj*nNx + ii
    coords = np.ascontiguousarray(np.stack([X.ravel(), Y.ravel(), Z.ravel()]))
    # elements (genTetra.cpp:250-334), 0-based corner ids pts[0..7], then +1
    kk, jj, ii = np.meshgrid(np.arange(nEz, dtype=np.int64), np.arange(nEy, dtype=np.int64),
                             np.arange(nEx, dtype=np.int64), indexing="ij")
    nn = nNx * nNy
    p0 = (nn * kk + nNx * jj + ii).ravel()
    p1 = p0 + 1
    p2 = p0 + nNx
    p3 = p2 + 1
    p4 = p0 + nn
    p5 = p4 + 1
    p6 = p4 + nNx
    p7 = p6 + 1
    tets = [(p0, p1, p3, p5), (p0, p3, p2, p5), (p2, p3, p7, p5), (p4, p6, p7, p2), (p4, p7, p5, p2), (p0, p4, p5, p2)]
    ncell = p0.size
    conn = np.empty((4, 6 * ncell), dtype=np.int32)
    for t, quad in enumerate(tets):
        for a in range(4):
            conn[a, t::6] = quad[a] + 1
    if dbc == "poisson":
        K, J, I = np.meshgrid(np.arange(nNz), np.arange(nNy), np.arange(nNx), indexing="ij")
        on = (I == 0) | (I == nNx - 1) | (J == 0) | (J == nNy - 1) | (K == 0) | (K == nNz - 1)
        nodes = np.flatnonzero(on.ravel())                      # sorted unique, genTetra.cpp:510-511
        fx = ax.astype(np.float32).astype(np.float64)
        fy = ay.astype(np.float32).astype(np.float64)
        fz = az.astype(np.float32).astype(np.float64)
        i_ = nodes % nNx
        j_ = (nodes // nNx) % nNy
        k_ = nodes // nn
        val = fx[i_] * fx[i_] + fy[j_] * fy[j_] + fz[k_] * fz[k_]
        # 8-decimal text round trip, vectorised: the distinct values are few
        uniq, inv = np.unique(val, return_inverse=True)
        val = _text_round(uniq)[inv]
        dn = (nodes + 1).astype(np.int32)
        dd = np.ones(nodes.size, np.int32)
        dv = val
    elif dbc == "clamp_y0":
        K, I = np.meshgrid(np.arange(nNz), np.arange(nNx), indexing="ij")
        nodes = (K * nn + I).ravel() + 1
        nodes.sort()
        dn = np.repeat(nodes, ndof).astype(np.int32)
        dd = np.tile(np.arange(1, ndof + 1), nodes.size).astype(np.int32)
        dv = np.zeros(dn.size)
    else:
        raise ValueError(dbc)
    return Mesh(coords, conn, dn, dd, dv, name=f"tet{nEx}x{nEy}x{nEz}")


def gen_tetra_gpu(x0, x1, nEx, y0, y1, nEy, z0, z1, nEz, dbc: str = "poisson", ndof: int = 1, device: int = 0) -> Mesh:
    """gen_tetra on the GPU (csrc/gpu_setup.cu, pfem_gpu_gen_tetra): only the per-axis accumulation `xx += dx` (n+1 values per
    axis) runs on the host; nodes, tets and Dirichlet rows are formed by kernels.  Bit-identical to gen_tetra."""
    import ctypes as C

    from . import solver as S
    lib = S.load_library()
    lib.pfem_gpu_gen_tetra.restype = C.c_longlong
    ax, ay, az = _accumulate(x0, x1, nEx), _accumulate(y0, y1, nEy), _accumulate(z0, z1, nEz)
    mode = {"poisson": 0, "clamp_y0": 1}[dbc]
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None   # noqa: E731
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None      # noqa: E731
    nrows = lib.pfem_gpu_gen_tetra(device, nEx, nEy, nEz, dp(ax), dp(ay), dp(az), mode, ndof, None, None, None, None, None)
    if nrows < 0:
        raise S.PfemError(int(-nrows), lib.pfem_last_error().decode())
    nN, nE = (nEx + 1) * (nEy + 1) * (nEz + 1), 6 * nEx * nEy * nEz
    coords = np.zeros((3, nN))
    conn = np.zeros((4, nE), np.int32)
    dn, dd, dv = np.zeros(nrows, np.int32), np.zeros(nrows, np.int32), np.zeros(nrows)
    rc = lib.pfem_gpu_gen_tetra(device, nEx, nEy, nEz, dp(ax), dp(ay), dp(az), mode, ndof, dp(coords), ip(conn), ip(dn), ip(dd), dp(dv))
    if rc < 0:
        raise S.PfemError(int(-rc), lib.pfem_last_error().decode())
    return Mesh(coords, conn, dn, dd, dv, name=f"tet{nEx}x{nEy}x{nEz}")


def exact_poisson_tria(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Analytic Laplace solution left in the comments of triapoissonparallelimpl1.F:954-955."""
    return (np.cosh(np.pi * y) - np.sinh(np.pi * y) / np.tanh(np.pi)) * np.sin(np.pi * x)


# ---- binary container (.pfemb): the arrays the drivers build from the text files, parsed once -----------------------
PFEMB_MAGIC = b"PFEMB1\0\0"


def write_binary(m: Mesh, path: str) -> None:
    """Write the PFEMB1 container (layout: csrc/host_meshio.cu).  Little endian, SoA arrays as the drivers hold them."""
    def i32(a):
        b = np.ascontiguousarray(a, "<i4").tobytes()
        return b + (b"\0\0\0\0" if (len(b) // 4) % 2 else b"")
    with open(path, "wb") as f:
        f.write(PFEMB_MAGIC)
        f.write(np.array([m.ndim, m.npElem, m.nNode, m.nElem, m.dbc_node.size, m.fbc_node.size], "<i8").tobytes())
        f.write(np.ascontiguousarray(m.coords, "<f8").tobytes())
        f.write(i32(m.conn))
        f.write(i32(m.dbc_node)); f.write(i32(m.dbc_dof)); f.write(np.ascontiguousarray(m.dbc_val, "<f8").tobytes())
        f.write(i32(m.fbc_node)); f.write(i32(m.fbc_dof)); f.write(np.ascontiguousarray(m.fbc_val, "<f8").tobytes())


def read_binary(path: str) -> Mesh:
    """Read a PFEMB1 container (memory-mapped: no parsing, no copy until the arrays are used)."""
    raw = np.memmap(path, dtype=np.uint8, mode="r")
    if raw.size < 56 or bytes(raw[:8]) != PFEMB_MAGIC:
        raise ValueError(f"{path} is not a PFEMB1 mesh container")
    ndim, npe, nN, nE, nD, nF = (int(v) for v in np.frombuffer(raw[8:56], "<i8"))
    off = 56

    def take(dtype, count, shape=None):
        nonlocal off
        nbytes = count * np.dtype(dtype).itemsize
        a = np.frombuffer(raw[off:off + nbytes], dtype)
        off += nbytes + (4 if (np.dtype(dtype).itemsize == 4 and count % 2) else 0)
        return a.reshape(shape) if shape else a

    coords = take("<f8", ndim * nN, (ndim, nN))
    conn = take("<i4", npe * nE, (npe, nE))
    dn, dd, dv = take("<i4", nD), take("<i4", nD), take("<f8", nD)
    fn, fd, fv = take("<i4", nF), take("<i4", nF), take("<f8", nF)
    if off != raw.size:
        raise ValueError(f"{path}: size does not match its header")
    return Mesh(np.array(coords), np.array(conn), np.array(dn), np.array(dd), np.array(dv), np.array(fn), np.array(fd),
                np.array(fv), name=os.path.basename(path))
