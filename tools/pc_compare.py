#!/usr/bin/env python
"""Jacobi vs the reference's default PCBJACOBI/ILU(0) on one GPU: iterations and solve time (genTetra n^3 Poisson, rtol 1e-10).
usage: python tools/pc_compare.py [cells ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfemfort_b200 import driver as D, mesh as M, solver as S  # noqa: E402

for cells in [int(a) for a in sys.argv[1:]] or [40, 100]:
    m = M.gen_tetra(-1, 1, cells, -1, 1, cells, -1, 1, cells)
    num = D.number(m, S.POISSON_TETRA)
    s = S.SolverB200(0)
    for pc, name in ((S.PC_JACOBI, "jacobi"), (S.PC_BJACOBI_ILU0, "bjacobi/ilu0")):
        best = None
        for _ in range(2):
            info = D.run_rank(s, m, num, rtol=1e-10, pc_type=pc)
            best = info if best is None or info["t_solve"] < best["t_solve"] else best
        print(json.dumps(dict(cells=cells, rows=num.size_global, pc=name, its=best["its"], reason=best["reason"],
                              solve_ms=1e3 * best["t_solve"], ms_per_iteration=1e3 * best["t_solve"] / max(best["its"], 1))), flush=True)
    s.free()
