"""Size-independent properties of the assembled systems, shared by the CPU (oracle, small sizes) and GPU (full size) tests."""
import numpy as np
import scipy.sparse as sp


def nnz_tet_poisson(n: int) -> int:
    """Pattern size of the genTetra n^3 x 6 Poisson system with all six faces fixed: N + 2 x (edges between free nodes).
    Free nodes: (n-1)^3.  Edges of the 6-tet split: axis edges, one diagonal per cell face, one body diagonal per cell."""
    N = (n - 1) ** 3
    edges = 3 * (n - 2) * (n - 1) ** 2 + 3 * (n - 2) ** 2 * (n - 1) + (n - 2) ** 3
    return N + 2 * edges


def nnz_tria_poisson(n: int) -> int:
    """Same for the n x n right-triangle mesh (the hypotenuse couplings are exact zeros but belong to the pattern)."""
    N = (n - 1) ** 2
    edges = 2 * (n - 2) * (n - 1) + (n - 2) ** 2
    return N + 2 * edges


def symmetric_to_rounding(rp, col, val, seed=0, tol=1e-12) -> bool:
    """x'Ay vs y'Ax for random vectors: equal to rounding iff A is symmetric (to rounding)."""
    n = rp.size - 1
    A = sp.csr_matrix((val, col, rp), shape=(n, n))
    rng = np.random.default_rng(seed)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    a, b = float(x @ (A @ y)), float(y @ (A @ x))
    scale = float(np.abs(x) @ (abs(A) @ np.abs(y)))
    return abs(a - b) <= tol * scale


def values_within(rp, val, ref, tol=1e-12) -> bool:
    """Every entry within tol of the reference RELATIVE TO THE LARGEST ENTRY OF ITS ROW (the diagonal for these
    operators): the scale-aware reading of "within 1e-12 relative" that stays meaningful for entries that cancel to
    (nearly) zero, such as the hypotenuse couplings of the right-triangle meshes."""
    val, ref = np.asarray(val), np.asarray(ref)
    if val.shape != ref.shape:
        return False
    if ref.size == 0:
        return True
    lens = np.diff(rp)
    rows = np.repeat(np.arange(lens.size), lens)
    rowmax = np.zeros(lens.size)
    np.maximum.at(rowmax, rows, np.abs(ref))
    scale = np.where(rowmax[rows] > 0, rowmax[rows], 1.0)
    return bool((np.abs(val - ref) <= tol * scale).all())


def vector_within(v, ref, tol=1e-12) -> bool:
    v, ref = np.asarray(v), np.asarray(ref)
    scale = float(np.abs(ref).max()) if ref.size else 1.0
    return bool(np.abs(v - ref).max() <= tol * (scale or 1.0)) if ref.size else True


def same_system(s, rp, val, oval, rhs, orhs) -> bool:
    """Assembled values / RHS against the oracle, at the bar of the kernel that ran: the reference-order kernels (modes 0-2)
    sum in the reference's sequential order without FMAs => bit-identical; the FMA kernels (mode 4: default row gather
    with the FMA operators, mode 3: colour-scheduled tiles) => 1e-12 relative, the north-star contract."""
    if s.assembly_mode()[0] < 3:
        return bool(np.array_equal(val, oval) and np.array_equal(rhs, orhs))
    return values_within(rp, val, oval) and vector_within(rhs, orhs)
