"""CPU emulation of the DEFAULT value-pass kernels (pfemfort_b200/csrc/assembly_rows.cuh: the streamed row gather
assemble_sell_kernel and its binary-search fallback assemble_kernel), for all four element kinds: the kernel source
is compiled for the host through tests/emu/cuda_shim.h, run CTA by CTA, and compared bit for bit with the oracle.

What this adds to the GPU parity tests: the product's measured kernels are regression-tested in the CPU suite that
runs every round, so a change to their index logic, staging, summation order or arithmetic is caught without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as O
from pfemfort_b200 import driver as D, mesh as M, solver as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
LIB = os.path.join(EMU, "_build", "libemu_rows.so")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    srcs = [os.path.join(EMU, f) for f in ("emu_rows.cpp", "cuda_shim.h", "emu_formats.hpp")] + \
           [os.path.join(ROOT, "pfemfort_b200", "csrc", f) for f in ("assembly_rows.cuh", "elements.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(s) for s in srcs):
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-DPFEM_EMULATE",
               "-Dpfem=pfem_emu", "-Wl,-Bsymbolic", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "pfemfort_b200", "csrc"),
               "-I", EMU, srcs[0], "-o", LIB]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-4000:]
    return C.CDLL(LIB)


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _load(name, input_dir):
    if name == "beam3Dtet6366":
        return M.read_mesh(os.path.join(input_dir, name), swap_34=True), S.ELASTICITY_TETRA
    kind = {"tria20x20": S.POISSON_TRIA, "tet10": S.POISSON_TETRA, "cookmembranetria32": S.ELASTICITY_TRIA}[name]
    return M.read_mesh(os.path.join(input_dir, name)), kind


def _run(emu, m, kind, num, rank=0, elemData=None, timeData=None, R=32, streamed=1, val=None, rhs=None):
    elemData = D.DEFAULT_ELEMDATA[kind] if elemData is None else elemData
    timeData = D.DEFAULT_TIMEDATA if timeData is None else timeData
    lo, hi = num.row_range(rank)
    grp, gcol = O.pattern(num.elemDof, num.size_global)
    rp = np.ascontiguousarray(grp[lo:hi + 1] - grp[lo], np.int32)
    col = np.ascontiguousarray(gcol[grp[lo]:grp[hi]], np.int32)
    conn0 = np.ascontiguousarray(num.conn_new - 1, np.int32)
    edof = np.ascontiguousarray(num.elemDof, np.int32)
    xyz_new = np.ascontiguousarray(m.coords[:, num.node_map_get_old - 1])
    load = 0 if val is None else 1
    val = np.zeros(max(col.size, 1)) if val is None else val.copy()
    rhs = np.zeros(max(hi - lo, 1)) if rhs is None else rhs.copy()
    flags = np.zeros(4, np.int32)
    ed = np.zeros(8)
    ed[:len(elemData)] = elemData
    td = np.zeros(8)
    td[:len(timeData)] = timeData
    rc = emu.emu_assemble_rows(kind, m.nElem, m.nNode, _ip(conn0), _ip(edof), _dp(xyz_new), _dp(num.solnApplied), lo, hi - lo,
                               _ip(rp), _ip(col), _dp(ed), _dp(td), R, streamed, load, _dp(val), _dp(rhs), _ip(flags))
    assert rc == 0, rc
    oval, orhs, nbad = O.assemble(kind, num.conn_new, m.coords, num.node_map_get_old, num.elemDof, num.solnApplied, elemData,
                                  timeData, grp, gcol, row_lo=lo, row_hi=hi)
    return val[:col.size], rhs[:hi - lo], oval[grp[lo]:grp[hi]], orhs[lo:hi], int(flags[0]), nbad


def _check(res):
    val, rhs, oval, orhs, neg, nbad = res
    assert nbad == 0 and neg == 0
    assert np.array_equal(val, oval), "values not bit-identical to the oracle"
    assert np.array_equal(rhs, orhs), "rhs not bit-identical to the oracle"


@pytest.mark.parametrize("name", ["tria20x20", "tet10", "cookmembranetria32", "beam3Dtet6366"])
@pytest.mark.parametrize("streamed,R", [(1, 32), (1, 128), (0, 128)])
def test_default_kernels_bit_exact_on_fixtures(emu, input_dir, name, streamed, R):
    m, kind = _load(name, input_dir)
    num = D.number(m, kind)
    _check(_run(emu, m, kind, num, R=R, streamed=streamed))


def test_non_unit_coefficients_and_accumulate(emu, input_dir):
    m, kind = _load("tet10", input_dir)
    num = D.number(m, kind)
    ed, td = [1.3, 0.7, 2.1], [0.0, 0.9, 0.0]
    for streamed in (1, 0):
        _check(_run(emu, m, kind, num, elemData=ed, timeData=td, R=128 if not streamed else 32, streamed=streamed))
    val, rhs, oval, orhs, _, _ = _run(emu, m, kind, num)
    v2, r2, _, _, _, _ = _run(emu, m, kind, num, val=val, rhs=rhs)
    grp, gcol = O.pattern(num.elemDof, num.size_global)
    o2, or2, _ = O.assemble(kind, num.conn_new, m.coords, num.node_map_get_old, num.elemDof, num.solnApplied,
                            D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA, grp, gcol, val=oval.copy(), rhs=orhs.copy())
    assert np.array_equal(v2, o2) and np.array_equal(r2, or2)


def test_partial_nonzero_dirichlet_elasticity(emu, input_dir):
    """Some dofs of a node fixed, others free, with non-zero applied values: lifting with several fixed dofs per element."""
    m, kind = _load("cookmembranetria32", input_dir)
    rng = np.random.default_rng(3)
    nodes = rng.choice(m.nNode, 60, replace=False) + 1
    m.dbc_node = np.concatenate([m.dbc_node, nodes.astype(np.int32)])
    m.dbc_dof = np.concatenate([m.dbc_dof, rng.integers(1, 3, nodes.size).astype(np.int32)])
    m.dbc_val = np.concatenate([m.dbc_val, rng.standard_normal(nodes.size)])
    num = D.number(m, kind)
    for streamed in (1, 0):
        _check(_run(emu, m, kind, num, R=128 if not streamed else 32, streamed=streamed))


@pytest.mark.parametrize("name", ["tet10", "beam3Dtet6366"])
def test_rank_row_blocks(emu, input_dir, name):
    m, kind = _load(name, input_dir)
    _, npart = D.partition(m, kind, 2)
    num = D.number(m, kind, 2, npart)
    for rank in range(2):
        _check(_run(emu, m, kind, num, rank=rank))


def test_negative_jacobian_is_flagged(emu, input_dir):
    m = M.read_mesh(os.path.join(input_dir, "beam3Dtet6366"))          # as shipped: every Jacobian negative
    num = D.number(m, S.ELASTICITY_TETRA)
    _, _, _, _, neg, nbad = _run(emu, m, S.ELASTICITY_TETRA, num)
    assert nbad > 0 and neg != 0
