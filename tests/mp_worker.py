"""Multi-rank worker (launched by torchrun from tests/test_gpu_multi.py and tests/test_multirank_cpu.py).

mode gpu : every rank drives one GPU through the C ABI (NCCL halo + all-reduce inside the library) and rank 0
           checks the gathered result against the sequential oracle.
mode cpu : gloo only, no GPU: every rank assembles its own row block from its owned + overlap elements with the
           oracle (the host-side decomposition the GPU path relies on) and rank 0 checks that the union is
           bit-identical to the global assembly: the claim that assembly needs no collective.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import pyoracle as O  # noqa: E402
from pfemfort_b200 import driver as D, mesh as M, solver as S  # noqa: E402


def load(name):
    inp = os.path.join(ROOT, "tests", "golden", "input")
    if name == "beam3Dtet6366":
        return M.read_mesh(os.path.join(inp, name), swap_34=True), S.ELASTICITY_TETRA
    if name.startswith("gen_tet"):
        n = int(name[7:])
        return M.gen_tetra(-1, 1, n, -1, 1, n, -1, 1, n), S.POISSON_TETRA
    if name.startswith("gen_tria"):
        return M.gen_tria_poisson(int(name[8:])), S.POISSON_TRIA
    kind = {"tria20x20": S.POISSON_TRIA, "tet10": S.POISSON_TETRA, "cookmembranetria32": S.ELASTICITY_TRIA}[name]
    return M.read_mesh(os.path.join(inp, name)), kind


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="gpu")
    ap.add_argument("--mesh", default="tet10")
    ap.add_argument("--partition", default="metis")
    ap.add_argument("--out", default="")
    ap.add_argument("--pc", default="jacobi", choices=["jacobi", "bjacobi"])
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    m, kind = load(a.mesh)
    npe, ndof, ndim = S.KIND_DIMS[kind]
    # rank 0 partitions and broadcasts (tetrapoissonparallelimpl1.F:457-484)
    npart = torch.zeros(m.nNode, dtype=torch.int32)
    if rank == 0:
        if a.partition == "metis":
            _, p = D.partition(m, kind, world)
        else:   # contiguous blocks of old node ids
            p = (np.arange(m.nNode, dtype=np.int64) * world // m.nNode).astype(np.int32)
        npart = torch.from_numpy(np.ascontiguousarray(p, dtype=np.int32))
    dist.broadcast(npart, 0)
    num = D.number(m, kind, world, npart.numpy())
    lo, hi = num.row_range(rank)
    result = {"rank": rank, "rows": [lo, hi]}
    if a.mode == "cpu":
        lst = D.local_elements(num, rank)
        conn = np.ascontiguousarray(num.conn_new[:, lst])
        edof = np.ascontiguousarray(num.elemDof[:, lst])
        grp, gcol = O.pattern(num.elemDof, num.size_global)        # every rank can build the global pattern on CPU
        val, rhs, nbad = O.assemble(kind, conn, m.coords, num.node_map_get_old, edof, num.solnApplied,
                                    D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA, grp, gcol, row_lo=lo, row_hi=hi)
        vals = [None] * world
        rhss = [None] * world
        dist.all_gather_object(vals, val[grp[lo]:grp[hi]])
        dist.all_gather_object(rhss, rhs[lo:hi])
        if rank == 0:
            gval, grhs, _ = O.assemble(kind, num.conn_new, m.coords, num.node_map_get_old, num.elemDof, num.solnApplied,
                                       D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA, grp, gcol)
            result["values_bit_identical"] = bool(np.array_equal(np.concatenate(vals), gval))
            result["rhs_bit_identical"] = bool(np.array_equal(np.concatenate(rhss), grhs))
            result["nnz"] = int(gcol.size)
            result["local_elements"] = int(lst.size)
    elif a.mode == "stash":
        # slow-path adds on rows of OTHER ranks (MatSetValues / VecSetValues on off-process rows): stashed, shipped to the
        # owner at the next assembly point (solve), added there.  Every rank adds the block (rank+1) * ones on the dofs of
        # a few interface elements; the gathered system must equal the batched assembly + the sum of all ranks' blocks.
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.frombuffer(bytearray(S.comm_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(idt, 0)
        s = S.SolverB200(device=local_rank, rank=rank, nranks=world, nccl_id=bytes(idt.numpy().tobytes()))
        D.run_rank(s, m, num, rank=rank, rtol=1e-10, do_solve=False)
        owner = np.searchsorted(np.array([num.row_range(q)[1] for q in range(world)]), np.maximum(num.elemDof, 0), side="right")
        free = num.elemDof >= 0
        mixed = [e for e in range(num.elemDof.shape[1]) if free[:, e].all() and len(set(owner[:, e])) > 1][:5]
        for e in mixed:
            dofs = num.elemDof[:, e]
            n = dofs.size
            s.add_matrix(dofs, dofs, np.full((n, n), float(rank + 1)))
            s.add_vector(dofs, np.full(n, 10.0 * (rank + 1)))
        s.factoriseAndSolve()                       # assembly point: the stash travels here
        rp, col, val = s.get_csr()
        rhs = s.get_rhs()
        parts = [None] * world
        dist.all_gather_object(parts, (rp, col, val, rhs))
        if rank == 0:
            grp, gcol = O.pattern(num.elemDof, num.size_global)
            gval, grhs, _ = O.assemble(kind, num.conn_new, m.coords, num.node_map_get_old, num.elemDof, num.solnApplied,
                                       D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA, grp, gcol)
            if m.fbc_node.size:
                O.add_force_bc(grhs, m.fbc_node, m.fbc_dof, m.fbc_val, ndof, num.node_map_get_new, num.NodeDofArrayNew, num.size_global)
            tot = float(sum(range(1, world + 1)))
            for e in mixed:
                dofs = num.elemDof[:, e]
                for i in dofs:
                    grhs[i] += 10.0 * tot
                    for j in dofs:
                        k = grp[i] + int(np.searchsorted(gcol[grp[i]:grp[i + 1]], j))
                        gval[k] += tot
            result["mixed_elements"] = len(mixed)
            result["values_match"] = bool(np.allclose(np.concatenate([p[2] for p in parts]), gval, rtol=1e-13, atol=1e-13))
            result["rhs_match"] = bool(np.allclose(np.concatenate([p[3] for p in parts]), grhs, rtol=1e-13, atol=1e-13))
        s.free()
    else:
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.frombuffer(bytearray(S.comm_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(idt, 0)
        s = S.SolverB200(device=local_rank, rank=rank, nranks=world, nccl_id=bytes(idt.numpy().tobytes()))
        info = D.run_rank(s, m, num, rank=rank, rtol=1e-10, pc_type=S.PC_BJACOBI_ILU0 if a.pc == "bjacobi" else S.PC_JACOBI)
        rp, col, val = s.get_csr()
        rhs = s.get_rhs()
        x = s.get_solution()
        result["comm_mode"] = s.comm_mode()
        result["assembly_mode"] = s.assembly_mode()[0]
        parts = [None] * world
        dist.all_gather_object(parts, (rp, col, val, rhs, info["its"], info["reason"]))
        if rank == 0:
            grp, gcol = O.pattern(num.elemDof, num.size_global)
            gval, grhs, _ = O.assemble(kind, num.conn_new, m.coords, num.node_map_get_old, num.elemDof, num.solnApplied,
                                       D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA, grp, gcol)
            if m.fbc_node.size:
                O.add_force_bc(grhs, m.fbc_node, m.fbc_dof, m.fbc_val, ndof, num.node_map_get_new, num.NodeDofArrayNew, num.size_global)
            if a.pc == "bjacobi":      # one ILU(0) block per rank, like PCBJACOBI under mpirun -np world
                starts = [num.row_range(q)[0] for q in range(world)] + [num.size_global]
                ox, oits, oreason, _ = O.cg_bjacobi_ilu0(grp, gcol, gval, grhs, block_start=starts, rtol=1e-10)
            else:
                ox, oits, oreason, _ = O.cg_jacobi(grp, gcol, gval, grhs, rtol=1e-10)
            rowlens = np.concatenate([np.diff(p[0]) for p in parts])
            result["pattern_bit_identical"] = bool(np.array_equal(rowlens, np.diff(grp)) and
                                                   np.array_equal(np.concatenate([p[1] for p in parts]), gcol))
            result["values_bit_identical"] = bool(np.array_equal(np.concatenate([p[2] for p in parts]), gval))
            result["rhs_bit_identical"] = bool(np.array_equal(np.concatenate([p[3] for p in parts]), grhs))
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from properties import values_within, vector_within
            result["values_within_1e-12"] = bool(result["pattern_bit_identical"] and
                                                 values_within(grp, np.concatenate([p[2] for p in parts]), gval))
            result["rhs_within_1e-12"] = bool(vector_within(np.concatenate([p[3] for p in parts]), grhs))
            result["its"] = [int(p[4]) for p in parts]
            result["reason"] = [int(p[5]) for p in parts]
            result["oracle_its"] = int(oits)
            result["oracle_reason"] = int(oreason)
            result["solution_rel_err"] = float(np.abs(x - ox).max() / np.abs(ox).max())
        s.free()
    if rank == 0 and a.out:
        with open(a.out, "w") as f:
            json.dump(result, f)
    if rank == 0:
        print(json.dumps(result), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
