"""N > 1 host-side logic on CPU: world_size-2 (and 3) gloo runs of the domain decomposition the GPU path uses."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle as O
from pfemfort_b200 import driver as D, mesh as M, solver as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, args, tmp_path, port):
    out = os.path.join(tmp_path, "result.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mp_worker.py"), "--out", out] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    with open(out) as f:
        return json.load(f)


@pytest.mark.parametrize("mesh,nproc,port", [("tet10", 2, 29611), ("beam3Dtet6366", 2, 29612), ("tria20x20", 3, 29613)])
def test_rank_blocks_reproduce_global_assembly(tmp_path, mesh, nproc, port):
    res = _torchrun(nproc, ["--mode", "cpu", "--mesh", mesh], str(tmp_path), port)
    assert res["values_bit_identical"] and res["rhs_bit_identical"]


def test_partition_numbering_invariants(input_dir):
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    for nparts in (2, 4, 8):
        _, npart = D.partition(m, S.POISSON_TETRA, nparts)
        assert set(np.unique(npart)) == set(range(nparts))
        num = D.number(m, S.POISSON_TETRA, nparts, npart)
        # maps are inverse permutations; each part owns a contiguous NEW node range in ascending OLD order
        assert np.array_equal(num.node_map_get_new[num.node_map_get_old - 1], np.arange(1, m.nNode + 1))
        for p in range(nparts):
            ns, ne = num.part_info[p, 0], num.part_info[p, 1]
            olds = num.node_map_get_old[ns - 1:ne]
            assert np.all(npart[olds - 1] == p) and np.all(np.diff(olds) > 0)
        assert num.part_info[:, 4].sum() == num.size_global == 729
        # row blocks tile [0, N) and every element is handed to at least one rank
        covered = np.zeros(m.nElem, bool)
        edges = [num.row_range(p) for p in range(nparts)]
        assert edges[0][0] == 0 and edges[-1][1] == num.size_global
        for p in range(nparts):
            covered[D.local_elements(num, p)] = True
            if p:
                assert edges[p][0] == edges[p - 1][1]
        has_free = (num.elemDof >= 0).any(axis=0)
        assert np.array_equal(covered, has_free)
        # the oracle's numbering agrees bit for bit
        o = O.number_dofs(m.nNode, 1, m.dbc_node, m.dbc_dof, m.dbc_val, nparts, npart)
        for k in ("node_map_get_old", "node_map_get_new", "NodeDofArrayNew", "solnApplied"):
            assert np.array_equal(getattr(num, k), o[k]), k
        assert np.array_equal(num.part_info[:, 2], o["row_start"]) and np.array_equal(num.part_info[:, 3], o["row_end"])
