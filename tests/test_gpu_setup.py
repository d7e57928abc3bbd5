"""GPU versions of the drivers' set-up loops (csrc/gpu_setup.cu): numbering / renumbering / ElemDofArray / assyForSoln /
element selection (tetrapoissonparallelimpl1.F:357-734) and the genTetra.cpp mesh recipe.  Integer work: every output must
be bit-identical to the sequential host loops (which tests/test_oracle.py pins against the oracle's orc_number_dofs)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from pfemfort_b200 import driver as D, explicit as X, mesh as M, solver as S

pytestmark = pytest.mark.gpu


def _same_numbering(a, b):
    for f in ("size_global", "nparts"):
        assert getattr(a, f) == getattr(b, f), f
    for f in ("node_map_get_old", "node_map_get_new", "NodeDofArrayNew", "solnApplied", "part_info", "conn_new", "elemDof"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


@pytest.mark.parametrize("name,kind,swap", [("tria20x20", S.POISSON_TRIA, False), ("tet10", S.POISSON_TETRA, False),
                                            ("beam3Dtet6366", S.ELASTICITY_TETRA, True), ("cookmembranetria32", S.ELASTICITY_TRIA, False)])
@pytest.mark.parametrize("nparts", [1, 2, 3, 8])
def test_gpu_numbering_equals_host_and_oracle(gpu, input_dir, name, kind, swap, nparts):
    m = M.read_mesh(os.path.join(input_dir, name), swap_34=swap)
    npart = None
    if nparts > 1:
        _, npart = D.partition(m, kind, nparts)
    host = D.number(m, kind, nparts, npart)
    dev = D.number(m, kind, nparts, npart, on_gpu=True)
    _same_numbering(host, dev)
    # and the oracle's restatement of the same driver block
    ndof = S.KIND_DIMS[kind][1]
    o = O.number_dofs(m.nNode, ndof, m.dbc_node, m.dbc_dof, m.dbc_val, nparts, npart)
    assert o["size_global"] == dev.size_global and np.array_equal(o["node_map_get_old"], dev.node_map_get_old)
    assert np.array_equal(o["NodeDofArrayNew"], dev.NodeDofArrayNew) and np.array_equal(o["solnApplied"], dev.solnApplied)
    assert np.array_equal(o["size_local"], dev.part_info[:, 4]) and np.array_equal(o["row_start"], dev.part_info[:, 2])
    for rank in range(nparts):
        lst, assy, edof = D.gpu_local_elements_and_assy(dev, rank)
        assert np.array_equal(lst, D.local_elements(host, rank)) and np.array_equal(edof, host.elemDof)
        assert np.array_equal(assy, X.free_slots(host))


def test_gpu_numbering_edge_cases(gpu):
    """Duplicate Dirichlet rows (the last one wins, like the sequential loop), an empty part, a random partition."""
    rng = np.random.default_rng(3)
    m = M.gen_tetra(-1, 1, 9, -1, 1, 7, -1, 1, 5, dbc="clamp_y0", ndof=3)
    dup = np.arange(0, m.dbc_node.size, 7)
    m.dbc_node = np.concatenate([m.dbc_node, m.dbc_node[dup]]).astype(np.int32)
    m.dbc_dof = np.concatenate([m.dbc_dof, m.dbc_dof[dup]]).astype(np.int32)
    m.dbc_val = np.concatenate([m.dbc_val, rng.standard_normal(dup.size)])
    npart = rng.integers(0, 5, m.nNode).astype(np.int32)
    npart[npart == 2] = 3                                       # part 2 owns nothing
    for nparts, p in ((1, None), (5, npart)):
        _same_numbering(D.number(m, S.ELASTICITY_TETRA, nparts, p), D.number(m, S.ELASTICITY_TETRA, nparts, p, on_gpu=True))
    bad = npart.copy()
    bad[5] = 9
    with pytest.raises(S.PfemError) as ei:
        D.number(m, S.ELASTICITY_TETRA, 5, bad, on_gpu=True)
    assert ei.value.status == S.ERR_NUMBERING


@pytest.mark.parametrize("args,kw", [((-2, 2, 10, -1, 1, 10, -1, 1, 10), {}), ((-1, 1, 17, -1, 1, 13, -1, 1, 11), {}),
                                     ((-0.5, 0.5, 8, 0.0, 6.0, 48, -0.5, 0.5, 8), dict(dbc="clamp_y0", ndof=3)),
                                     ((-1, 1, 100, -1, 1, 100, -1, 1, 100), {})])
def test_gpu_gen_tetra_equals_recipe(gpu, input_dir, args, kw):
    """pfem_gpu_gen_tetra against the numpy recipe (which reproduces the bundled tet10 / tet100 files string for string)."""
    host = M.gen_tetra(*args, **kw)
    dev = M.gen_tetra_gpu(*args, **kw)
    assert np.array_equal(host.coords, dev.coords) and np.array_equal(host.conn, dev.conn)
    assert np.array_equal(host.dbc_node, dev.dbc_node) and np.array_equal(host.dbc_dof, dev.dbc_dof)
    assert np.array_equal(host.dbc_val, dev.dbc_val)
    if args[2] == 10 and not kw:                                # the bundled fixture itself
        f = M.read_mesh(os.path.join(input_dir, "tet10"))
        assert np.array_equal(f.coords, dev.coords) and np.array_equal(f.conn, dev.conn) and np.array_equal(f.dbc_val, dev.dbc_val)
