"""CPU regression test of the explicit-dynamics GPU kernels: pfemfort_b200/csrc/explicit.cuh (the same source the product
compiles for sm_100a) is compiled for the host through tests/emu/cuda_shim.h and its lumped mass / time-loop state is
compared bit for bit with the sequential oracle.  Test infrastructure only: the product never runs this way."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as O
from pfemfort_b200 import driver as D, explicit as X, mesh as M, solver as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
LIB = os.path.join(EMU, "_build", "libemu_explicit.so")
ED2 = [200.0, 0.3, 10.0, 1.0, 0.0]
ED3 = [200.0, 0.3, 10.0, 0.5, -0.25, 1.0]
TD = [0.0, 1.0, 0.0]


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    srcs = [os.path.join(EMU, "emu_explicit.cpp")]
    deps = srcs + [os.path.join(EMU, "cuda_shim.h"), os.path.join(ROOT, "pfemfort_b200", "csrc", "explicit.cuh"),
                   os.path.join(ROOT, "pfemfort_b200", "csrc", "elements.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
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


def _device_layout(num, m, kind):
    """What pfem_explicit_set_mesh builds on the GPU: int4 node ids per element, AoS coordinates, node -> (e*npe+i) lists in
    ascending element id (a stable sort by node)."""
    npe, ndof, ndim = S.KIND_DIMS[kind]
    conn = num.conn_new - 1
    nE, nN = conn.shape[1], m.nNode
    conn4 = np.zeros((nE, 4), np.int32)
    conn4[:, :npe] = conn.T
    if npe == 3:
        conn4[:, 3] = conn4[:, 2]
    stride = 4 if ndim == 3 else 2
    xyz = np.zeros((nN, stride))
    xyz[:, :ndim] = m.coords.T
    keys = conn.T.ravel()                                   # t = e*npe + i
    order = np.argsort(keys, kind="stable").astype(np.int32)
    inc_ptr = np.searchsorted(keys[order], np.arange(nN + 1)).astype(np.int32)
    return np.ascontiguousarray(conn4), np.ascontiguousarray(xyz), inc_ptr, order


@pytest.mark.parametrize("name,kind,swap,ed,dt", [("cookmembranetria32", S.ELASTICITY_TRIA, False, ED2, 2e-4),
                                                  ("beam3Dtet6366", S.ELASTICITY_TETRA, True, ED3, 1e-3)])
def test_explicit_kernels_bit_exact_on_cpu(emu, input_dir, name, kind, swap, ed, dt):
    m = M.read_mesh(os.path.join(input_dir, name), swap_34=swap)
    num = D.number(m, kind)
    npe, ndof, ndim = S.KIND_DIMS[kind]
    conn4, xyz, inc_ptr, inc = _device_layout(num, m, kind)
    nd = m.nNode * ndof
    prm = np.zeros(8)
    prm[:len(ed)] = ed
    Mg = np.zeros(nd)
    neg = np.zeros(1, np.int32)
    emu.emu_explicit_mass(kind, m.nNode, _ip(inc_ptr), _ip(inc), _ip(conn4), _dp(xyz), _dp(prm), _dp(Mg), _ip(neg))
    Mo, nbad = O.explicit_lumped_mass(kind, num.conn_new, m.coords, ed)
    assert nbad == 0 and neg[0] == 0 and np.array_equal(Mg, Mo)
    fs = X.free_slots(num)
    mask = np.zeros(nd, np.uint8)
    mask[fs - 1] = 1
    d = [np.zeros(nd) for _ in range(3)]
    velo, acce = np.zeros(nd), np.zeros(nd)
    cur = 1
    steps = 12
    for _ in range(steps):                                    # the host loop of pfem_explicit_advance
        p2, nx = (cur + 1) % 3, (cur + 2) % 3
        emu.emu_explicit_step(kind, m.nNode, _ip(inc_ptr), _ip(inc), _ip(conn4), _dp(xyz), _dp(prm), _dp(Mg),
                              mask.ctypes.data_as(C.POINTER(C.c_ubyte)), _dp(d[cur]), _dp(d[p2]), _dp(d[nx]), _dp(velo), _dp(acce),
                              C.c_double(dt), _ip(neg))
        cur = nx
    st = O.explicit_advance(kind, num.conn_new, m.coords, fs, ed, TD, dt, steps, Mo)
    assert neg[0] == 0 and np.abs(st["disp"]).max() > 0
    assert np.array_equal(d[cur], st["disp"]) and np.array_equal(d[(cur + 1) % 3], st["dispPrev2"])
    assert np.array_equal(velo, st["velo"]) and np.array_equal(acce, st["acce"])


def test_explicit_kernel_counts_inverted_elements(emu, input_dir):
    m = M.read_mesh(os.path.join(input_dir, "cookmembranetria32"))
    kind = S.ELASTICITY_TRIA
    num = D.number(m, kind)
    num.conn_new[[0, 1], 5] = num.conn_new[[1, 0], 5]        # one clockwise triangle
    conn4, xyz, inc_ptr, inc = _device_layout(num, m, kind)
    prm = np.zeros(8)
    prm[:5] = ED2
    Mg = np.zeros(m.nNode * 2)
    neg = np.zeros(1, np.int32)
    emu.emu_explicit_mass(kind, m.nNode, _ip(inc_ptr), _ip(inc), _ip(conn4), _dp(xyz), _dp(prm), _dp(Mg), _ip(neg))
    assert neg[0] == 1                                       # every bad element is counted once (at its local node 0)
