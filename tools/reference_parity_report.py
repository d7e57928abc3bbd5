#!/usr/bin/env python
"""Tabulates, on the CPU, how the oracle and the host logic compare with every golden vector obtained by executing the
reference's own source (tests/golden/ref_*; made by tests/golden/make_reference_vectors.py).  Prints markdown; the committed
copy is profiles/r02_reference_parity.md.  (The CUDA path is held to the same files by tests/test_gpu_zzzz_reference_vectors.py.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as O                                   # noqa: E402
from pfemfort_b200 import driver as D, explicit as X               # noqa: E402
import test_reference_vectors as T                                 # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
INPUT = os.path.join(G, "input")


def maxdiff(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max()) if a.size else 0.0


def main():
    print("# Oracle and host logic against the executed reference (CPU)\n")
    print("`python tools/reference_parity_report.py`; golden files: `tests/golden/ref_*` (the reference's own Fortran executed by")
    print("`oracle/refrun`, its mesh generator compiled). `0` = bit-identical.\n")
    print("## Element routines (96 random elements per kind; negative-Jacobian STOPs must coincide)\n")
    print("| routine | elements compared | STOPs (ref = oracle) | max abs diff K | max abs diff F |")
    print("|---|---|---|---|---|")
    g = np.load(os.path.join(G, "ref_elements.npz"))
    names = ["StiffnessResidualPoissonLinearTria", "StiffnessResidualPoissonLinearTetra", "StiffnessResidualElasticityLinearTria",
             "StiffnessResidualElasticityLinearTetra"]
    for kind in range(4):
        xyz, ed, td, vc = g[f"ke{kind}_xyz"], g[f"ke{kind}_ed"], g[f"ke{kind}_td"], g[f"ke{kind}_valc"]
        dk = df = 0.0
        n = stops = 0
        for e in range(xyz.shape[0]):
            K, F, rc = O.element_ke(kind, xyz[e, 0], xyz[e, 1], xyz[e, 2], ed[e], td[e], vc[e])
            assert (rc != 0) == bool(g[f"ke{kind}_neg"][e])
            if rc:
                stops += 1
                continue
            n += 1
            dk, df = max(dk, maxdiff(K, g[f"ke{kind}_K"][e])), max(df, maxdiff(F, g[f"ke{kind}_F"][e]))
        print(f"| `{names[kind]}` | {n} | {stops} = {stops} | {dk:g} | {df:g} |")
    for kind, nm in ((2, "Tria"), (3, "Tetra")):
        xyz, ed, td, vc = g[f"ke{kind}_xyz"], g[f"ke{kind}_ed"], g[f"ke{kind}_td"], g[f"ke{kind}_valc"]
        dr = dm = 0.0
        n = 0
        for e in range(xyz.shape[0]):
            if g[f"ke{kind}_neg"][e]:
                continue
            n += 1
            dr = max(dr, maxdiff(O.residual_elasticity(kind, xyz[e, 0], xyz[e, 1], xyz[e, 2], ed[e], td[e], vc[e])[0], g[f"res{kind}_F"][e]))
            dm = max(dm, maxdiff(O.mass_matrix(kind, xyz[e, 0], xyz[e, 1], xyz[e, 2], ed[e])[0], g[f"mass{kind}_M"][e]))
        print(f"| `ResidualElasticityLinear{nm}` / `MassMatrixLinear{nm}` | {n} | - | {dr:g} (residual) | {dm:g} (mass) |")
    print("\n## The four `*parallelimpl1` PROGRAMs, end to end, on P simulated ranks\n")
    print("| case | ranks | N | nnz | integer arrays (maps, NodeDofArrayNew, ElemDofArray, assyForSoln, ranges) | pattern | max rel diff values | max rel diff RHS |")
    print("|---|---|---|---|---|---|---|---|")
    for name, p in T.CASE_IDS:
        gd = T._driver(name, p)
        m, kind = T._mesh(name, INPUT)
        o, conn_new, edof, rp, col, val, rhs = T._oracle_system(m, kind, p, gd["node_proc_id"] if p > 1 else None)
        ints = (np.array_equal(o["node_map_get_old"], gd["node_map_get_old"]) and np.array_equal(o["NodeDofArrayNew"].T, gd["NodeDofArrayNew"])
                and np.array_equal(edof.T, gd["ElemDofArray"]) and np.array_equal(X.free_slots(D.number(m, kind, p, gd["node_proc_id"] if p > 1 else None)), gd["assyForSoln"])
                and np.array_equal(np.c_[o["node_start"], o["node_end"], o["size_local"]], gd["part_info"][:, [0, 1, 4]]))
        patt = np.array_equal(rp, gd["rowptr"]) and np.array_equal(col, gd["col"])
        rows = np.repeat(np.arange(rp.size - 1), np.diff(rp))
        rowmax = np.zeros(rp.size - 1)
        np.maximum.at(rowmax, rows, np.abs(gd["val"]))
        dv = float((np.abs(val - gd["val"]) / np.where(rowmax[rows] > 0, rowmax[rows], 1.0)).max())
        dr = maxdiff(rhs, gd["rhs"]) / float(np.abs(gd["rhs"]).max())
        print(f"| {name} | {p} | {rp.size - 1} | {col.size} | {'identical' if ints else 'DIFFER'} | {'identical' if patt else 'DIFFERS'} | {dv:.2g} | {dr:.2g} |")
    print("\n(one rank: bit-identical; P ranks: the reference's own sums depend on the order PETSc's stash delivers off-rank rows, bar 1e-12)\n")
    print("## Explicit dynamics: `triaelasticityexplicit.F`, 40 steps, cook membrane\n")
    ge = np.load(os.path.join(G, "ref_explicit_cookmembranetria32.npz"))
    m, kind = T._mesh("cookmembranetria32", INPUT)
    num = D.number(m, kind)
    Mg, _ = O.explicit_lumped_mass(kind, num.conn_new, m.coords, X.DRIVER_ELEMDATA_TRIA)
    st = O.explicit_advance(kind, num.conn_new, m.coords, X.free_slots(num), X.DRIVER_ELEMDATA_TRIA, X.DRIVER_TIMEDATA, X.DRIVER_DT,
                            int(ge["steps"]), Mg)
    print("| quantity | max abs diff |")
    print("|---|---|")
    print(f"| lumped mass | {maxdiff(Mg, ge['globalM']):g} |")
    for k in ("disp", "dispPrev2", "velo", "acce"):
        print(f"| {k} after {int(ge['steps'])} steps | {maxdiff(st[k], ge[k]):g} |")
    print("\n## `TYPE PetscSolver` procedures (harness PROGRAM) and solver options\n")
    gs = np.load(os.path.join(G, "ref_solver_procedures.npz"))
    print(f"* `assembleMatrix` / `assembleVector` / `assembleMatrixAndVector`: not transposed, negative indices skipped (22 entries, 6 rows); "
          f"`factorise` STOPs at solverpetsc.F:{int(gs['stop_mode1_line'])}, `solve` at :{int(gs['stop_mode2_line'])} when called out of order.")
    gd = T._driver("tet10", 1)
    print(f"* every driver run: `KSPSetType({gd['ksp_type'][0]})`, `PCSetType({gd['pc_type'][0]})`; options {sorted(str(o) for o in set(gd['options']))}.")


if __name__ == "__main__":
    main()
