"""The C++ counterpart of the Fortran driver PROGRAMs (drivers/pfem_driver.cpp): same command line, same text input
files, same temp.dat output; linked against libpfemb200.so only."""
import gzip
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle as O
from pfemfort_b200 import driver as D, mesh as M, solver as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "pfemfort_b200", "bin", "pfem_driver")


def _unpack(name, input_dir, tmp):
    out = []
    for part in ("nodes", "elems", "DirichBC", "ForceBC"):
        src = os.path.join(input_dir, f"{name}-{part}.dat.gz")
        if not os.path.exists(src):
            continue
        dst = os.path.join(tmp, f"{name}-{part}.dat")
        with gzip.open(src, "rb") as f, open(dst, "wb") as g:
            shutil.copyfileobj(f, g)
        out.append(dst)
    return out


def _oracle_solution(m, kind, rtol):
    num = D.number(m, kind)
    npe, ndof, ndim = S.KIND_DIMS[kind]
    rp, col = O.pattern(num.elemDof, num.size_global)
    val, rhs, _ = O.assemble(kind, num.conn_new, m.coords, None, num.elemDof, num.solnApplied, D.DEFAULT_ELEMDATA[kind],
                             D.DEFAULT_TIMEDATA, rp, col)
    if m.fbc_node.size:
        O.add_force_bc(rhs, m.fbc_node, m.fbc_dof, m.fbc_val, ndof, num.node_map_get_new, num.NodeDofArrayNew, num.size_global)
    x, its, reason, _ = O.cg_jacobi(rp, col, val, rhs, rtol=rtol)
    return num, x, its


@pytest.mark.parametrize("name,phys,kind", [("tet10", "tetrapoisson", S.POISSON_TETRA), ("tria20x20", "triapoisson", S.POISSON_TRIA),
                                            ("cookmembranetria32", "triaelasticity", S.ELASTICITY_TRIA)])
def test_cpp_driver_single_rank(gpu, input_dir, tmp_path, name, phys, kind):
    files = _unpack(name, input_dir, str(tmp_path))
    env = dict(os.environ, PFEM_KSP_RTOL="1e-10", PFEM_PC_TYPE="jacobi")
    r = subprocess.run([DRIVER, phys] + files, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = M.read_mesh(os.path.join(input_dir, name))
    num, ox, oits = _oracle_solution(m, kind, 1e-10)
    assert f"Convergence in {oits} iterations." in r.stdout or f"Convergence in {oits + 1} iterations." in r.stdout \
        or f"Convergence in {oits - 1} iterations." in r.stdout, r.stdout
    t = np.loadtxt(os.path.join(str(tmp_path), "temp.dat"))
    npe, ndof, ndim = S.KIND_DIMS[kind]
    # the Poisson PROGRAMs write "ii  node  value" (tetrapoissonparallelimpl1.F:940), the elasticity PROGRAMs the value only
    # (tetraelasticityparallelimpl1.F:1046)
    assert t.ndim == (2 if ndof == 1 else 1) and t.shape[0] == num.size_global
    vals = t[:, 2] if ndof == 1 else t
    assert np.abs(vals - ox).max() <= 1e-7 * np.abs(ox).max()
    # ... and against the temp.dat records of the reference's own PROGRAM executed on the same files (tests/golden/ref_driver_*)
    g = np.load(os.path.join(ROOT, "tests", "golden", f"ref_driver_{name}_p1.npz"))
    assert np.abs(vals - g["temp_dat_value"]).max() <= 1e-7 * np.abs(g["temp_dat_value"]).max()
    if ndof == 1:
        # the second column is the (old-numbering) node slot of every free dof, like assyForSoln
        free = np.flatnonzero(num.NodeDofArrayNew.T.ravel() > 0) + 1
        assert np.array_equal(t[:, 1].astype(int), free)
        assert np.array_equal(t[:, :2].astype(int), g["temp_dat_index"])


def test_cpp_driver_two_ranks(gpu, input_dir, tmp_path):
    if S.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    files = _unpack("tet10", input_dir, str(tmp_path))
    env = dict(os.environ, PFEM_KSP_RTOL="1e-10", PFEM_PC_TYPE="jacobi", PFEM_NCCL_ID_FILE=os.path.join(str(tmp_path), "nccl.id"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", DRIVER, "tetrapoisson"] + files
    r = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    t = np.loadtxt(os.path.join(str(tmp_path), "temp.dat"))
    # temp.dat: (dof index in the partition-renumbered system, old node id, value): compare nodally with the exact solution
    u = np.zeros(m.nNode)
    u[t[:, 1].astype(int) - 1] = t[:, 2]
    free = t[:, 1].astype(int) - 1
    assert np.abs(u[free] - (m.coords[:, free] ** 2).sum(0)).max() < 2e-7


def test_cpp_driver_reference_defaults_and_options_file(gpu, input_dir, tmp_path):
    """No options at all: the reference's coded defaults, CG + PCBJACOBI/ILU(0) at rtol 1e-5 (solverpetsc.F:187,206);
    with a petsc_options.dat in the working directory (PetscInitialize, tetrapoissonparallelimpl1.F:168): its options."""
    files = _unpack("tet10", input_dir, str(tmp_path))
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    num = D.number(m, S.POISSON_TETRA)
    rp, col = O.pattern(num.elemDof, num.size_global)
    val, rhs, _ = O.assemble(S.POISSON_TETRA, num.conn_new, m.coords, None, num.elemDof, num.solnApplied,
                             D.DEFAULT_ELEMDATA[S.POISSON_TETRA], D.DEFAULT_TIMEDATA, rp, col)
    env = {k: v for k, v in os.environ.items() if not k.startswith("PFEM_")}
    r = subprocess.run([DRIVER, "tetrapoisson"] + files, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    ox, oits, _, _ = O.cg_bjacobi_ilu0(rp, col, val, rhs, rtol=1e-5)
    assert f"Convergence in {oits} iterations." in r.stdout, r.stdout
    t = np.loadtxt(os.path.join(str(tmp_path), "temp.dat"))
    assert np.abs(t[:, 2] - ox).max() <= 1e-6 * np.abs(ox).max()
    with open(os.path.join(str(tmp_path), "petsc_options.dat"), "w") as f:
        f.write("-ksp_type cg\n-pc_type jacobi\n-ksp_rtol 1e-10\n")
    r = subprocess.run([DRIVER, "tetrapoisson"] + files, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    ox, oits, _, _ = O.cg_jacobi(rp, col, val, rhs, rtol=1e-10)
    assert any(f"Convergence in {k} iterations." in r.stdout for k in (oits - 1, oits, oits + 1)), r.stdout
