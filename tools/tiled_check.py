#!/usr/bin/env python
"""Compare the tiled (PFEM_ASM=tiled / tiled2) value passes with the default row-gather kernel on one GPU: bit-identity of
the CSR values / RHS and the device time of one pass (CUDA events inside the library).  numpy + ctypes only.

usage: tiled_check.py [n ...]   (genTetra n^3 x 6 Poisson meshes; default 100 200)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pfemfort_b200 import driver as D, mesh as M, solver as S  # noqa: E402

# mode (tiled = staged columns + gather, tiled2 = sorted scatter + run sums), threads, rows per tile, KB of smem per CTA
VARIANTS = [("tiled", 256, 96, 110), ("tiled2", 256, 96, 112), ("tiled2", 512, 192, 224), ("tiled2", 128, 64, 74)]


def one_pass(m, kind, num, reps=3):
    s = S.SolverB200(0)
    t0 = time.perf_counter()
    D.run_rank(s, m, num, do_solve=False)
    first = time.perf_counter() - t0
    ts = []
    for _ in range(reps):
        s.setZero()
        s.assemble(D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA)
        ts.append(s.info()["t_assemble"])
    rp, col, val = s.get_csr()
    rhs = s.get_rhs()
    mode = s.assembly_mode()
    s.free()
    return val, rhs, mode, min(ts), first


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [100, 200]
    out = []
    for n in sizes:
        m = M.gen_tetra(-1, 1, n, -1, 1, n, -1, 1, n)
        kind = S.POISSON_TETRA
        num = D.number(m, kind)
        os.environ.pop("PFEM_ASM", None)
        val0, rhs0, mode0, t0, w0 = one_pass(m, kind, num)
        rec = dict(n=n, elements=m.nElem, dof=num.size_global, default_ms=1e3 * t0, default_mode=mode0[0], default_first_wall_s=w0)
        print(json.dumps(rec), flush=True)
        for asm, threads, rows, kb in VARIANTS:
            os.environ.update(PFEM_ASM=asm, PFEM_TILE_THREADS=str(threads), PFEM_TILE_ROWS=str(rows), PFEM_TILE_SMEM_KB=str(kb))
            try:
                val1, rhs1, mode1, t1, w1 = one_pass(m, kind, num)
                r = dict(n=n, asm=asm, threads=threads, rows=rows, smem_kb=kb, tiled_ms=1e3 * t1, mode=mode1[0], ntiles=mode1[1],
                         visits_per_element=mode1[2], first_wall_s=w1, values_bit_identical=bool(np.array_equal(val0, val1)),
                         rhs_bit_identical=bool(np.array_equal(rhs0, rhs1)),
                         max_abs_diff=float(np.abs(val0 - val1).max()), melem_per_s=m.nElem / t1 / 1e6)
            except Exception as ex:   # keep going: the other variants still tell something
                r = dict(n=n, asm=asm, threads=threads, rows=rows, error=str(ex))
            print(json.dumps(r), flush=True)
            out.append(r)
        out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tiled_check.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
