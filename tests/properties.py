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
