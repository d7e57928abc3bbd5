#!/usr/bin/env python
"""The colour-scheduled tile value pass (default) against the row-gather kernels (PFEM_ASM=rows) on one GPU:
agreement (1e-12 relative to the row maximum), run-to-run bit-identity, device time per pass, tile statistics.
numpy + ctypes only.  usage: ctile_check.py [tetN ...|triaN ...]  (default: tet12 tria40 tet100 tet200)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pfemfort_b200 import driver as D, mesh as M, solver as S  # noqa: E402
from properties import values_within, vector_within  # noqa: E402


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
    s.setZero()
    s.assemble(D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA)
    val2 = s.get_csr()[2]
    info = s.assembly_info()
    mode = s.assembly_mode()
    s.free()
    return rp, val, rhs, mode, info, min(ts), first, bool(np.array_equal(val, val2))


def main():
    cases = sys.argv[1:] or ["tet12", "tria40", "tet100", "tet200"]
    for c in cases:
        if c.startswith("tet"):
            n = int(c[3:])
            m, kind = M.gen_tetra(-1, 1, n, -1, 1, n, -1, 1, n), S.POISSON_TETRA
        else:
            n = int(c[4:])
            m, kind = M.gen_tria_poisson(n), S.POISSON_TRIA
        num = D.number(m, kind)
        os.environ["PFEM_ASM"] = "rows"
        rp, val0, rhs0, mode0, info0, t0, w0, _ = one_pass(m, kind, num)
        os.environ.pop("PFEM_ASM")
        variants = [("fast", {"PFEM_ASM": "fast"}), ("ctile", {"PFEM_ASM": "ctile"})] + [(f"{rule}_B{b}_TR{tr}", {"PFEM_ASM": "ctile", "PFEM_TILE_THREADS": str(b), "PFEM_TILE_ROWS": str(tr), "PFEM_TILE_RULE": rule})
                                        for rule, b, tr in (("full", 512, 1024), ("full", 256, 1024), ("full", 384, 768), ("position", 512, 1024),
                                                            ("position", 1024, 1024), ("position", 384, 1024))]
        if os.environ.get("CTILE_VARIANTS") == "0":
            variants = variants[:2]
        for name, env in variants:
            os.environ.update(env)
            try:
                _, val1, rhs1, mode1, info1, t1, w1, rep = one_pass(m, kind, num)
                r = dict(case=c, variant=name, elements=m.nElem, dof=num.size_global, rows_ms=1e3 * t0, rows_mode=mode0[0],
                         ctile_ms=1e3 * t1, mode=mode1[0], kernel=info1["kernel"], first_wall_s=w1, rows_first_wall_s=w0,
                         values_within_1e12=values_within(rp, val1, val0), rhs_within_1e12=vector_within(rhs1, rhs0),
                         max_abs_diff=float(np.abs(val0 - val1).max()), max_abs=float(np.abs(val0).max()),
                         run_to_run_bit_identical=rep, melem_per_s=m.nElem / t1 / 1e6)
            except Exception as ex:
                r = dict(case=c, variant=name, error=str(ex))
            for k in env:
                os.environ.pop(k, None)
            print(json.dumps(r), flush=True)
            if n < 40:
                break


if __name__ == "__main__":
    main()
