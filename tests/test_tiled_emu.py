"""CPU emulation of the tiled (compute-once) value pass: the kernel source of pfemfort_b200/csrc/assembly_tiled.cuh and
the host tile builder tiles.hpp, compiled for the host through tests/emu/cuda_shim.h and run CTA by CTA (one OS thread
per CUDA thread, a barrier for __syncthreads), against the oracle.

This checks what can be checked without a GPU: the tile construction, the staging/slot/summation-order logic and the
no-FMA arithmetic (bit-exact).  Memory-model behaviour and performance are GPU matters (tests/test_gpu_tiled.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as O
from pfemfort_b200 import driver as D, mesh as M, solver as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
LIB = os.path.join(EMU, "_build", "libemu_tiled.so")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    srcs = [os.path.join(EMU, "emu_tiled.cpp"), os.path.join(EMU, "cuda_shim.h")] + \
           [os.path.join(ROOT, "pfemfort_b200", "csrc", f) for f in ("assembly_tiled.cuh", "tiles.hpp", "elements.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(s) for s in srcs):
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-fopenmp", "-DPFEM_EMULATE",
               # own namespace + symbolic binding: libpfemb200.so (RTLD_GLOBAL) exports the same template names (CUDA stubs)
               "-Dpfem=pfem_emu", "-Wl,-Bsymbolic", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "pfemfort_b200", "csrc"), "-I", EMU, srcs[0], "-o", LIB]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-4000:]
    return C.CDLL(LIB)


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


MODE = [1]      # 1: staged columns + row gather, 2: sorted scatter + run sums (set by the autouse fixture below)


@pytest.fixture(autouse=True, params=[1, 2], ids=["gather", "scatter"])
def _mode(request):
    MODE[0] = request.param
    yield


def _run(emu, m, kind, num, rank=0, elemData=None, timeData=None, tile_rows=96, smem=100 * 1024, threads=128, val=None,
         rhs=None):
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
    stats = np.zeros(8, np.int64)
    ed = np.zeros(8)
    ed[:len(elemData)] = elemData
    td = np.zeros(8)
    td[:len(timeData)] = timeData
    rc = emu.emu_assemble_tiled(kind, m.nElem, m.nNode, _ip(conn0), _ip(edof), _dp(xyz_new), _dp(num.solnApplied), lo, hi - lo,
                                _ip(rp), _ip(col), _dp(ed), _dp(td), tile_rows, smem, threads, load, _dp(val), _dp(rhs),
                                stats.ctypes.data_as(C.POINTER(C.c_longlong)), MODE[0])
    assert rc == 0, rc
    # the oracle on the same rows
    oval, orhs, nbad = O.assemble(kind, num.conn_new, m.coords, num.node_map_get_old, num.elemDof, num.solnApplied, elemData,
                                  timeData, grp, gcol, row_lo=lo, row_hi=hi)
    return val[:col.size], rhs[:hi - lo], oval[grp[lo]:grp[hi]], orhs[lo:hi], stats, nbad


def _check(res):
    val, rhs, oval, orhs, stats, nbad = res
    assert nbad == 0 and stats[4] == 0
    assert np.array_equal(val, oval), "values not bit-identical to the oracle"
    assert np.array_equal(rhs, orhs), "rhs not bit-identical to the oracle"
    return stats


@pytest.mark.parametrize("name,kind", [("tria20x20", S.POISSON_TRIA), ("tet10", S.POISSON_TETRA)])
@pytest.mark.parametrize("tile_rows,threads", [(32, 128), (96, 128), (128, 128), (96, 256), (256, 256), (192, 512)])
def test_fixtures_bit_exact(emu, input_dir, name, kind, tile_rows, threads):
    m = M.read_mesh(os.path.join(input_dir, name))
    num = D.number(m, kind)
    stats = _check(_run(emu, m, kind, num, tile_rows=tile_rows, threads=threads, smem=200 * 1024))
    assert stats[0] >= (num.size_global + tile_rows - 1) // tile_rows
    assert stats[1] >= stats[2] > 0                        # every touched element is visited at least once


def test_generated_meshes_and_small_smem_budget(emu):
    m = M.gen_tetra(-1, 1, 7, -1, 1, 6, -1, 1, 5)
    num = D.number(m, S.POISSON_TETRA)
    s1 = _check(_run(emu, m, S.POISSON_TETRA, num))
    s2 = _check(_run(emu, m, S.POISSON_TETRA, num, smem=40 * 1024))     # the shared-memory budget splits the tiles
    assert s2[0] > s1[0] and s2[3] <= 40 * 1024
    m = M.gen_tria_poisson(23)
    num = D.number(m, S.POISSON_TRIA)
    _check(_run(emu, m, S.POISSON_TRIA, num, tile_rows=64))


def test_non_unit_coefficients_and_interior_dirichlet(emu, input_dir):
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    rng = np.random.default_rng(5)
    # extra Dirichlet nodes in the interior with non-zero values: lifting with several fixed dofs per element
    extra = rng.choice(m.nNode, 150, replace=False) + 1
    extra = np.setdiff1d(extra, m.dbc_node)
    m.dbc_node = np.concatenate([m.dbc_node, extra.astype(np.int32)])
    m.dbc_dof = np.ones(m.dbc_node.size, np.int32)
    m.dbc_val = np.concatenate([m.dbc_val, rng.standard_normal(extra.size)])
    num = D.number(m, S.POISSON_TETRA)
    _check(_run(emu, m, S.POISSON_TETRA, num))
    _check(_run(emu, m, S.POISSON_TETRA, num, elemData=[1.3, 0.7, 2.1], timeData=[0.0, 0.9, 0.0]))
    m2 = M.read_mesh(os.path.join(input_dir, "tria20x20"))
    num2 = D.number(m2, S.POISSON_TRIA)
    _check(_run(emu, m2, S.POISSON_TRIA, num2, elemData=[1.3, 0.7], timeData=[0.0, 0.9, 0.0]))


def test_accumulate_on_existing_values(emu, input_dir):
    """A second value pass without setZero adds on top, entry by entry in the same order (MatSetValues ADD twice)."""
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    num = D.number(m, S.POISSON_TETRA)
    val, rhs, oval, orhs, _, _ = _run(emu, m, S.POISSON_TETRA, num)
    v2, r2, _, _, _, _ = _run(emu, m, S.POISSON_TETRA, num, val=val, rhs=rhs)
    grp, gcol = O.pattern(num.elemDof, num.size_global)
    o2, or2, _ = O.assemble(S.POISSON_TETRA, num.conn_new, m.coords, num.node_map_get_old, num.elemDof, num.solnApplied,
                            D.DEFAULT_ELEMDATA[S.POISSON_TETRA], D.DEFAULT_TIMEDATA, grp, gcol, val=oval.copy(), rhs=orhs.copy())
    assert np.array_equal(v2, o2) and np.array_equal(r2, or2)


@pytest.mark.parametrize("nparts", [2, 3])
def test_rank_row_blocks(emu, input_dir, nparts):
    """Multi-rank layout: renumbered nodes, a rank's row block, columns owned by other ranks, overlap elements."""
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    _, npart = D.partition(m, S.POISSON_TETRA, nparts)
    num = D.number(m, S.POISSON_TETRA, nparts, npart)
    for rank in range(nparts):
        _check(_run(emu, m, S.POISSON_TETRA, num, rank=rank, tile_rows=64))


def test_negative_jacobian_flag(emu, input_dir):
    m = M.read_mesh(os.path.join(input_dir, "tet10"))
    e = int(np.flatnonzero((D.number(m, S.POISSON_TETRA).elemDof >= 0).all(axis=0))[5])     # an interior element
    m.conn[[0, 1], e] = m.conn[[1, 0], e]
    num = D.number(m, S.POISSON_TETRA)
    val, rhs, oval, orhs, stats, nbad = _run(emu, m, S.POISSON_TETRA, num)
    assert nbad == 1 and stats[4] == 1


def test_rows_without_elements_and_tiny_meshes(emu):
    """Free nodes that no element references (empty rows) and meshes smaller than one tile slice."""
    m = M.gen_tetra(-1, 1, 3, -1, 1, 3, -1, 1, 3)
    extra = np.array([[5.0, 6.0], [5.0, 6.0], [5.0, 7.0]])             # two stray free nodes, not in any element
    m.coords = np.ascontiguousarray(np.concatenate([m.coords, extra], axis=1))
    num = D.number(m, S.POISSON_TETRA)
    val, rhs, oval, orhs, stats, nbad = _run(emu, m, S.POISSON_TETRA, num, tile_rows=32)
    assert nbad == 0 and np.array_equal(val, oval) and np.array_equal(rhs, orhs)
    assert num.size_global == 8 + 2 and rhs[-1] == 0.0 and rhs[-2] == 0.0
    m = M.gen_tria_poisson(2)                                           # one free node, 8 triangles
    num = D.number(m, S.POISSON_TRIA)
    assert num.size_global == 1
    _check(_run(emu, m, S.POISSON_TRIA, num, tile_rows=32))
