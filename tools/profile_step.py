#!/usr/bin/env python
"""One short pass of the hot path for ncu: set-up, one warm step, one profiled step with a capped
iteration count.  Usage: ncu ... python tools/profile_step.py --cells 200 --max-it 12"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfemfort_b200 import driver as D, mesh as M, solver as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=200)
ap.add_argument("--max-it", type=int, default=12)
ap.add_argument("--kind", default="poisson", choices=["poisson", "elasticity"])
a = ap.parse_args()
if a.kind == "poisson":
    m = M.gen_tetra(-1, 1, a.cells, -1, 1, a.cells, -1, 1, a.cells)
    kind = S.POISSON_TETRA
else:   # the beam of config C4, scaled by --cells (50 -> 50x300x50)
    c = a.cells
    m = M.gen_tetra(-0.5, 0.5, c, 0.0, 6.0, 6 * c, -0.5, 0.5, c, dbc="clamp_y0", ndof=3)
    kind = S.ELASTICITY_TETRA
num = D.number(m, kind)
s = S.SolverB200(0)
for _ in range(2):
    info = D.run_rank(s, m, num, rtol=1e-10, max_it=a.max_it)
print(info, s.launch_count())
s.free()
