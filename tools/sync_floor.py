#!/usr/bin/env python
"""Synchronisation floor of the persistent CG kernel: us per iteration on a problem so small (cells=16: 3375 rows) that
bandwidth time is nil, so what remains is the cost of the grid barriers / all-reduces of one iteration.
usage: python tools/sync_floor.py [cells ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfemfort_b200 import driver as D, mesh as M, solver as S  # noqa: E402

for cells in [int(a) for a in sys.argv[1:]] or [16, 32]:
    m = M.gen_tetra(-1, 1, cells, -1, 1, cells, -1, 1, cells)
    num = D.number(m, S.POISSON_TETRA)
    for sync in ("last", "lean"):                      # barrier flavour of the persistent kernel (cg.cu: pcg_sync / pcg_sync_ctr)
        os.environ["PFEM_PCG_SYNC"] = sync
        s = S.SolverB200(0)
        best = None
        for _ in range(4):
            info = D.run_rank(s, m, num, rtol=1e-30, max_it=2000)
            us = 1e6 * info["t_solve"] / max(info["its"], 1)
            best = us if best is None else min(best, us)
        print(json.dumps(dict(cells=cells, rows=num.size_global, sync=sync, its=info["its"], us_per_iteration=best)), flush=True)
        s.free()
