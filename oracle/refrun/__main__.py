"""python -m oracle.refrun <driver.F> <program arguments...> [--ranks P] [--shipped] [--cwd DIR] [--steps N]

Runs one of the reference's PROGRAMs from its own source (default /root/reference/src, or $PFEM_REFERENCE_SRC) on P
simulated MPI ranks and prints what it did: STOP message if any, the solver options the wrapper set, the size of the
system handed to KSPSolve, the first temp.dat records.  `--shipped` runs the text without the documented-intent
substitutions.  TEST INFRASTRUCTURE: the way to reproduce tests/golden/ref_driver_*.npz by hand, e.g.

    python -m oracle.refrun tetrapoissonparallelimpl1.F tet10-nodes.dat tet10-elems.dat tet10-DirichBC.dat --cwd /tmp/inputs
"""
import argparse
import sys

import numpy as np

from . import run_reference as R


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m oracle.refrun")
    ap.add_argument("driver")
    ap.add_argument("args", nargs="*")
    ap.add_argument("--ranks", type=int, default=1)
    ap.add_argument("--shipped", action="store_true")
    ap.add_argument("--cwd", default=".")
    ap.add_argument("--steps", type=int, default=0, help="explicit drivers: replace the hard-coded stepsMax = 50000")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args(argv)
    if not R.available():
        sys.exit(f"reference sources not found under {R.REF_SRC}")
    part = None
    if a.ranks > 1:
        import os
        nodes = sum(1 for line in open(os.path.join(a.cwd, a.args[0])) if line.strip())
        conn = np.array([[int(x) for x in line.split()[1:]] for line in open(os.path.join(a.cwd, a.args[1])) if line.strip()])
        npid = (np.arange(nodes) * a.ranks) // nodes          # contiguous blocks stand in for METIS (third-party)
        part = (npid[conn[:, 0] - 1], npid)
    patches = [("stepsMax = 50000", f"stepsMax = {a.steps}", 1)] if a.steps else None
    res = R.run_driver(a.driver, a.args, a.ranks, partition=part, cwd=a.cwd, intent=not a.shipped, quiet=not a.verbose,
                       extra_patches=patches)
    print("stopped:", res.stopped)
    for rank, name, info in res.world.trace:
        if rank == 0 and name in ("KSPSetType", "PCSetType", "MatSetOption", "VecSetOption", "KSPSolve"):
            print(f"  {name}: {info}")
    if res.system is not None:
        rowptr, col, val, rhs = res.system
        print(f"system at KSPSolve: N = {rowptr.size - 1}, nnz = {col.size}, sum diag = "
              f"{val[col == np.repeat(np.arange(rowptr.size - 1), np.diff(rowptr))].sum()!r}, |b|_2 = {np.linalg.norm(rhs)!r}")
    for fname, recs in res.ranks[0].written.items():
        print(f"{fname}: {len(recs)} records; first: {recs[:2]}")
    for note in sorted(set(n for rt in res.ranks for n in rt.notes)):
        print("note:", note)
    return 0


if __name__ == "__main__":
    sys.exit(main())
