#!/usr/bin/env python
"""Convert the reference's text input (``<prefix>-nodes|elems|DirichBC[|ForceBC].dat[.gz]``) into one PFEMB1 binary
container (layout: pfemfort_b200/csrc/host_meshio.cu), or generate one of the synthetic benchmark meshes straight into it.

    mesh_convert.py <prefix> out.pfemb [--swap34]
    mesh_convert.py --gen-tetra N out.pfemb          (genTetra recipe, N^3 x 6 Poisson on [-1,1]^3: config C5 at N = 200)
    mesh_convert.py --gen-tria N out.pfemb           (n x n right-triangle Poisson mesh: config C2 at N = 1000)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfemfort_b200 import mesh as M  # noqa: E402

a = sys.argv[1:]
if len(a) >= 3 and a[0] == "--gen-tetra":
    n = int(a[1])
    m, out = M.gen_tetra(-1, 1, n, -1, 1, n, -1, 1, n), a[2]
elif len(a) >= 3 and a[0] == "--gen-tria":
    m, out = M.gen_tria_poisson(int(a[1])), a[2]
elif len(a) >= 2:
    m, out = M.read_mesh(a[0], swap_34="--swap34" in a), a[1]
else:
    sys.exit(__doc__)
M.write_binary(m, out)
print(f"{out}: {m.nNode} nodes, {m.nElem} elements, {m.dbc_node.size} Dirichlet rows, {m.fbc_node.size} force rows, "
      f"{os.path.getsize(out)} bytes")
