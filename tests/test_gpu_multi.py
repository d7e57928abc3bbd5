"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped on the single-GPU tier)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    from pfemfort_b200 import solver
    return solver.device_count()


@pytest.mark.parametrize("mesh,nproc,port", [("tet10", 2, 29621), ("beam3Dtet6366", 2, 29622), ("cookmembranetria32", 2, 29623),
                                             ("gen_tet24", 2, 29624), ("gen_tet24", 4, 29625), ("gen_tet24", 8, 29626)])
@pytest.mark.parametrize("comm,asm", [("p2p", "rows"), ("p2p", "fast"), ("nccl", "rows")])
def test_multi_gpu_matches_oracle(gpu, tmp_path, mesh, nproc, port, comm, asm):
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    os.environ["PFEM_COMM"] = comm          # p2p: peer-memory kernels over NVLink (default); nccl: the NCCL path
    os.environ.pop("PFEM_ASM", None)
    if asm == "fast":
        os.environ["PFEM_ASM"] = "fast"     # FMA element operators: 1e-12 contract instead of bit-identity
    out = os.path.join(str(tmp_path), "result.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mp_worker.py"), "--mode", "gpu", "--mesh", mesh, "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    assert res["comm_mode"] == (2 if comm == "p2p" else 1)
    os.environ.pop("PFEM_ASM", None)
    assert res["pattern_bit_identical"] and res["values_within_1e-12"] and res["rhs_within_1e-12"]
    if asm == "rows":
        assert res["assembly_mode"] in (0, 1) and res["values_bit_identical"] and res["rhs_bit_identical"]
    else:
        assert res["assembly_mode"] == 4
    assert all(x == res["oracle_reason"] == 2 for x in res["reason"])
    assert len(set(res["its"])) == 1
    assert abs(res["its"][0] - res["oracle_its"]) <= max(1, 0.02 * res["oracle_its"])
    assert res["solution_rel_err"] < 1e-7


@pytest.mark.parametrize("sync,halo,port", [("last", "flag", 29631), ("lean", "flag", 29632), ("last", "tag", 29633)])
@pytest.mark.parametrize("mesh,nproc", [("gen_tet24", 2), ("beam3Dtet6366", 2), ("gen_tet24", 4)])
def test_multi_gpu_persistent_variants(gpu, tmp_path, mesh, nproc, sync, halo, port):
    """The non-default flavours of the persistent kernel's grid barrier (PFEM_PCG_SYNC) and halo (PFEM_PCG_HALO: values +
    per-neighbour flags instead of tag-validated entries) give the same result as the defaults checked above."""
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    env = dict(os.environ, PFEM_COMM="p2p", PFEM_PCG_SYNC=sync, PFEM_PCG_HALO=halo)
    env.pop("PFEM_ASM", None)
    out = os.path.join(str(tmp_path), "result.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port + 10 * nproc), os.path.join(ROOT, "tests", "mp_worker.py"), "--mode", "gpu", "--mesh", mesh, "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    assert res["comm_mode"] == 2 and res["pattern_bit_identical"] and res["values_bit_identical"] and res["rhs_bit_identical"]
    assert all(x == res["oracle_reason"] == 2 for x in res["reason"]) and len(set(res["its"])) == 1
    assert abs(res["its"][0] - res["oracle_its"]) <= max(1, 0.02 * res["oracle_its"])
    assert res["solution_rel_err"] < 1e-7


@pytest.mark.parametrize("mesh,nproc,port", [("gen_tet24", 2, 29661), ("beam3Dtet6366", 2, 29662), ("gen_tet24", 4, 29663)])
@pytest.mark.parametrize("comm", ["nccl", "p2p"])
def test_multi_gpu_bjacobi_ilu0(gpu, tmp_path, mesh, nproc, port, comm):
    """The reference's default PC under mpirun -np N: block Jacobi with one ILU(0) block per rank (solverpetsc.F:206).
    Iteration counts must equal the oracle's CG with the same blocks."""
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    env = dict(os.environ, PFEM_COMM=comm, PFEM_KERNELS_P2P="1" if comm == "p2p" else "0")
    env.pop("PFEM_ASM", None)
    out = os.path.join(str(tmp_path), "result.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port + (5 if comm == "p2p" else 0)), os.path.join(ROOT, "tests", "mp_worker.py"), "--mode", "gpu", "--mesh", mesh,
           "--pc", "bjacobi", "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    assert res["pattern_bit_identical"] and res["values_bit_identical"] and res["rhs_bit_identical"]
    assert all(x == res["oracle_reason"] == 2 for x in res["reason"]) and len(set(res["its"])) == 1
    assert abs(res["its"][0] - res["oracle_its"]) <= max(1, 0.02 * res["oracle_its"])
    assert res["solution_rel_err"] < 1e-7


@pytest.mark.parametrize("mesh,nproc,port", [("tet10", 2, 29671), ("cookmembranetria32", 2, 29672), ("gen_tet24", 4, 29673)])
def test_multi_gpu_slow_path_stash(gpu, tmp_path, mesh, nproc, port):
    """MatSetValues / VecSetValues mirrors on rows owned by ANOTHER rank are stashed and shipped to the owner at the next
    assembly point, like PETSc's stash (solverpetsc.F:447-468): nothing is dropped."""
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    out = os.path.join(str(tmp_path), "result.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mp_worker.py"), "--mode", "stash", "--mesh", mesh, "--out", out]
    env = dict(os.environ)
    env.pop("PFEM_ASM", None)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    assert res["mixed_elements"] > 0 and res["values_match"] and res["rhs_match"], res
