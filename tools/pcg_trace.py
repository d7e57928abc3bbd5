#!/usr/bin/env python
"""Stage timing of the persistent CG kernel's last iteration (CTA 0, thread 0, clock64), from a -DPFEM_PCG_TRACE build:
    PFEM_EXTRA_NVCC=-DPFEM_PCG_TRACE python pfemfort_b200/build.py --force ; python tools/pcg_trace.py [cells]
Prints cycles between consecutive stamps: direction loop | barrier pieces | SpMV loop | barrier+reduce | update | ..."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfemfort_b200 import driver as D, mesh as M, solver as S  # noqa: E402

NAMES = {0: "iteration top", 1: "direction loop done", 2: "  CTA barrier #1", 3: "  release fence", 4: "  red + poll", 5: "  acquire fence",
         6: "  (partials) done", 7: "plain barrier left", 10: "SpMV loop done", 11: "  CTA barrier #1", 12: "  block sum + stcg + release fence",
         13: "  red + poll", 14: "  acquire fence", 15: "  partial sums (+ mailbox) done", 16: "scalar step done", 17: "CTA barrier #2 left",
         20: "update loop done", 21: "  CTA barrier #1", 22: "  block sums + stcg + release fence", 23: "  red + poll", 24: "  acquire fence",
         25: "  partial sums (+ mailbox) done", 26: "scalar step done", 27: "CTA barrier #2 left"}
for cells in ([int(a) for a in sys.argv[1:]] or [16]) if __name__ == "__main__" else []:
    m = M.gen_tetra(-1, 1, cells, -1, 1, cells, -1, 1, cells)
    num = D.number(m, S.POISSON_TETRA)
    s = S.SolverB200(0)
    for _ in range(2):
        info = D.run_rank(s, m, num, rtol=1e-30, max_it=300)
    out = np.zeros(64, np.int64)
    rc = S.load_library().pfem_debug_pcg_trace(out.ctypes.data_as(C.POINTER(C.c_longlong)))
    assert rc == 0, "not a PFEM_PCG_TRACE build"
    print(f"cells {cells}: {1e6 * info['t_solve'] / info['its']:.2f} us/iteration over {info['its']} iterations")
    keys = sorted(k for k in NAMES if out[k] > 0)
    t0 = out[0]
    prev = t0
    for k in keys:
        print(f"  [{k:2d}] {NAMES[k]:42s} +{out[k] - prev:6d} cyc   (t = {out[k] - t0:6d})")
        prev = out[k]
    s.free()
