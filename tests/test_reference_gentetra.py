"""Mesh generation against the reference's OWN generator, compiled here: oracle/_ref/genTetranovtk is
/root/reference/src/genTetranovtk.cpp built by `make -C oracle ref` (g++ alone; genTetra.cpp itself needs VTK).
tests/golden/ref_gentetra.json holds the sha-256 of the files that binary writes for six grids
(tests/golden/make_reference_vectors.py); where the binary is present (build container, GPU box) it is also run live.

Node and element files: byte for byte, any grid.  Dirichlet node list: cubic grids (every BASELINE configuration).  On
non-cubic grids the reference's two generator variants disagree with each other and with the geometry (genTetranovtk.cpp:281
runs the x = x0 face loop to nNx instead of nNy -- listing a node id beyond nNode when nNx > nNy; genTetra.cpp:414 does
the same on the x = x1 face): `gen_tetra` lists the geometric faces there."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import ref_gentetra as G
from pfemfort_b200 import mesh as M

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_gentetra.json")
with open(GOLDEN) as _f:
    GRIDS = json.load(_f)


def _sha(b):
    return hashlib.sha256(b).hexdigest()


def check_against_golden(name, mesh):
    g = GRIDS[name]
    assert mesh.coords.shape[1] == g["nNode"] and mesh.conn.shape[1] == g["nElem"]
    nodes, elems = G.mesh_text(mesh.coords, mesh.conn)
    assert _sha(nodes) == g["sha_nodes"], "mesh-nodes.dat differs from the reference generator's"
    assert _sha(elems) == g["sha_elems"], "mesh-elems.dat differs from the reference generator's"
    nEx, nEy, nEz = g["grid"][2], g["grid"][5], g["grid"][8]
    if nEx == nEy == nEz:
        assert mesh.dbc_node.size == g["nDBC"]
        assert _sha(mesh.dbc_node.astype("<i4").tobytes()) == g["sha_dbc_nodes"]


@pytest.mark.parametrize("name", sorted(GRIDS))
def test_host_gen_tetra_writes_the_reference_generators_files(name):
    check_against_golden(name, M.gen_tetra(*GRIDS[name]["grid"]))


@pytest.mark.skipif(not G.available(), reason="oracle/_ref/genTetranovtk is built where the reference tree is present")
@pytest.mark.parametrize("grid", [(-1, 1, 12, -1, 1, 12, -1, 1, 12), (0.0, 3.0, 4, -1.0, 1.0, 9, 5.0, 5.5, 2)])
def test_host_gen_tetra_against_a_live_run_of_the_compiled_reference(grid):
    r = G.run(*grid)
    m = M.gen_tetra(*grid)
    assert np.array_equal(m.coords, r["coords"]) and np.array_equal(m.conn, r["conn"])
    if grid[2] == grid[5] == grid[8]:
        assert np.array_equal(m.dbc_node, r["dbc_node"]) and np.array_equal(m.dbc_dof, r["dbc_dof"])


@pytest.mark.skipif(not os.path.exists(G.REF_SOURCE), reason="the reference tree exists only in the build container")
def test_golden_digests_are_what_the_compiled_reference_writes_today():
    assert G.build()
    for name, g in GRIDS.items():
        r = G.run(*g["grid"])
        assert (r["sha_nodes"], r["sha_elems"]) == (g["sha_nodes"], g["sha_elems"]), name
