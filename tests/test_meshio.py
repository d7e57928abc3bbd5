"""Mesh input either side of the hot path: the reference's text tables parsed in one pass
by the library (csrc/host_meshio.cu) and the PFEMB1 binary container holding the same arrays.  Host code: no GPU needed."""
import ctypes as C
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

from pfemfort_b200 import mesh as M, solver as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = ["tria20x20", "tet10", "cookmembranetria32", "beam3Dtet6366"]
FIELDS = ("coords", "conn", "dbc_node", "dbc_dof", "dbc_val", "fbc_node", "fbc_dof", "fbc_val")


def _gunzip(prefix, part, dst_dir):
    src = f"{prefix}-{part}.dat.gz"
    if not os.path.exists(src):
        return None
    dst = os.path.join(dst_dir, f"{os.path.basename(prefix)}-{part}.dat")
    with gzip.open(src, "rb") as f, open(dst, "wb") as g:
        shutil.copyfileobj(f, g)
    return dst


@pytest.mark.parametrize("name", FIXTURES)
def test_binary_container_round_trip(input_dir, tmp_path, name):
    m = M.read_mesh(os.path.join(input_dir, name))
    path = os.path.join(str(tmp_path), name + ".pfemb")
    M.write_binary(m, path)
    r = M.read_binary(path)
    for k in FIELDS:
        assert np.array_equal(getattr(m, k), getattr(r, k)), k
    assert r.coords.dtype == np.float64 and r.conn.dtype == np.int32


@pytest.mark.parametrize("name", FIXTURES)
def test_c_reader_and_writer_agree_with_python(input_dir, tmp_path, name):
    lib = S.load_library()
    m = M.read_mesh(os.path.join(input_dir, name))
    # python writes, C reads
    path = os.path.join(str(tmp_path), name + ".pfemb")
    M.write_binary(m, path)
    sz = (C.c_longlong * 6)()
    assert lib.pfem_host_mesh_read_binary_header(path.encode(), sz) == 0
    assert list(sz) == [m.ndim, m.npElem, m.nNode, m.nElem, m.dbc_node.size, m.fbc_node.size]
    coords, conn = np.zeros_like(m.coords), np.zeros_like(m.conn)
    dn, dd, dv = np.zeros_like(m.dbc_node), np.zeros_like(m.dbc_dof), np.zeros_like(m.dbc_val)
    fn, fd, fv = (np.zeros(max(m.fbc_node.size, 1), np.int32), np.zeros(max(m.fbc_node.size, 1), np.int32),
                  np.zeros(max(m.fbc_node.size, 1)))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.pfem_host_mesh_read_binary(path.encode(), dp(coords), ip(conn), ip(dn), ip(dd), dp(dv), ip(fn), ip(fd), dp(fv)) == 0
    nf = m.fbc_node.size
    for a, b in ((coords, m.coords), (conn, m.conn), (dn, m.dbc_node), (dd, m.dbc_dof), (dv, m.dbc_val), (fn[:nf], m.fbc_node),
                 (fd[:nf], m.fbc_dof), (fv[:nf], m.fbc_val)):
        assert np.array_equal(a, b)
    # C writes, python reads: byte-identical files
    path2 = os.path.join(str(tmp_path), name + "_c.pfemb")
    assert lib.pfem_host_mesh_write_binary(path2.encode(), m.ndim, m.npElem, m.nNode, m.nElem, dp(np.ascontiguousarray(m.coords)),
                                           ip(np.ascontiguousarray(m.conn)), m.dbc_node.size, ip(m.dbc_node), ip(m.dbc_dof),
                                           dp(m.dbc_val), nf, ip(m.fbc_node) if nf else None, ip(m.fbc_dof) if nf else None,
                                           dp(m.fbc_val) if nf else None) == 0
    assert open(path, "rb").read() == open(path2, "rb").read()
    # a truncated or foreign file is refused
    with open(path2, "r+b") as f:
        f.truncate(os.path.getsize(path2) - 8)
    assert lib.pfem_host_mesh_read_binary(path2.encode(), dp(coords), ip(conn), ip(dn), ip(dd), dp(dv), ip(fn), ip(fd), dp(fv)) != 0
    with pytest.raises(ValueError):
        M.read_binary(path2)


@pytest.mark.parametrize("name", ["tet10", "beam3Dtet6366"])
def test_one_pass_text_parser_matches_numpy(input_dir, tmp_path, name):
    """pfem_host_read_table on the reference's text files == numpy.loadtxt (the values are read with strtod)."""
    prefix = os.path.join(input_dir, name)
    for part, ncols in (("nodes", 4), ("elems", 5), ("DirichBC", 3), ("ForceBC", 3)):
        txt = _gunzip(prefix, part, str(tmp_path))
        if txt is None:
            continue
        with open(txt, "a") as f:
            f.write("\n   \n")                                  # trailing blank lines are skipped, like the drivers' READ
        a = M._load_table(txt, ncols)
        b = np.loadtxt(txt, dtype=np.float64, ndmin=2)
        assert a.shape == b.shape and np.array_equal(a, b)


def test_text_parser_edge_cases(tmp_path):
    """Ragged rows are skipped without stealing numbers from the next line; exponents, signs, long mantissas, CRLF and
    extra columns are handled; values are the correctly rounded doubles (== Python's float())."""
    path = os.path.join(str(tmp_path), "t.dat")
    rows = ["1 0.1 -2.5e-3 7", "2 3.0", "", "3 1e300 +4 5 99 98", "4 0.12345678901234567890 -0.0 6\r", "   ", "5 .5 5. 1.5D2 x"]
    with open(path, "w") as f:
        f.write("\n".join(rows))                          # no trailing newline
    a = M._load_table(path, 4)
    want = [[1, 0.1, -2.5e-3, 7], [3, 1e300, 4, 5], [4, float("0.12345678901234567890"), -0.0, 6], [5, 0.5, 5.0, 150.0]]
    assert a.shape == (4, 4) and np.array_equal(a, np.array(want))        # 1.5D2: Fortran's double-precision exponent letter
    assert np.signbit(a[2, 2])
    a5 = M._load_table(path, 5)                           # only the row with five numbers qualifies; "x" is not a number
    assert a5.shape == (1, 5) and np.array_equal(a5[0], [3, 1e300, 4, 5, 99])
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.standard_normal(2000) * 10.0 ** rng.integers(-8, 8, 2000), rng.integers(-10**9, 10**9, 500)])
    for fmt in ("%.8f", "%.17g", "%.3e", "%g"):
        with open(path, "w") as f:
            for i, v in enumerate(vals):
                f.write(f"{i + 1} {fmt % v}\n")
        a = M._load_table(path, 2)
        assert np.array_equal(a[:, 1], np.array([float(fmt % v) for v in vals])), fmt


def test_cpp_driver_reads_text_and_binary_alike(input_dir, tmp_path):
    """The compiled driver's host side (reading, numbering) on the text files and on the container: same sizes, same
    DOF count; without a GPU it then stops where the device is needed (with one it writes temp.dat: tests/test_gpu_driver_cpp.py)."""
    drv = os.path.join(ROOT, "pfemfort_b200", "bin", "pfem_driver")
    assert os.path.exists(drv), "run __graft_entry__.build()"
    prefix = os.path.join(input_dir, "beam3Dtet6366")
    m = M.read_mesh(prefix, swap_34=True)
    pfemb = os.path.join(str(tmp_path), "beam.pfemb")
    M.write_binary(m, pfemb)
    outs = []
    for args in ([pfemb],):
        r = subprocess.run([drv, "tetraelasticity"] + args, cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
        outs.append(r.stdout)
        assert "nElem_global = 7776" in r.stdout and f"nNode_global = {m.nNode}" in r.stdout and "Total DOF = 5292" in r.stdout
        if S.device_count() == 0:
            assert r.returncode != 0 and "no CUDA device" in r.stderr
    files = [_gunzip(prefix, p, str(tmp_path)) for p in ("nodes", "elems", "DirichBC", "ForceBC")]
    dump = os.path.join(str(tmp_path), "from_text.pfemb")
    r = subprocess.run([drv, "tetraelasticity"] + files, cwd=str(tmp_path), capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, PFEM_WRITE_PFEMB=dump))
    assert "nElem_global = 7776" in r.stdout and "Total DOF = 5292" in r.stdout
    # what the driver parsed from the text files is, byte for byte, the container python writes from its own reading
    m_text = M.read_mesh(prefix)                                        # (the driver reads the files as shipped: no 3<->4 swap)
    ref = os.path.join(str(tmp_path), "ref.pfemb")
    M.write_binary(m_text, ref)
    assert open(dump, "rb").read() == open(ref, "rb").read()
    # a container of the wrong element type is refused before anything else happens
    r = subprocess.run([drv, "triapoisson", pfemb], cwd=str(tmp_path), capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "3D mesh" in r.stderr


def test_text_parser_is_correctly_rounded_on_awkward_decimals(tmp_path):
    """The drivers READ list-directed (correctly rounded by libgfortran); the one-pass parser must give the same doubles:
    40 000 tokens -- 8-decimal fixed, shortest round-trip reprs over 60 decades, 17-digit exponents, 20-digit fractions,
    integers -- against Python's correctly rounded float()."""
    import ctypes as C
    lib = S.load_library()
    lib.pfem_host_read_table.restype = C.c_longlong
    lib.pfem_host_read_table.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_double), C.c_longlong]
    rng = np.random.default_rng(1)
    toks = []
    for i in range(40000):
        k = i % 6
        if k == 0:
            toks.append(f"{rng.standard_normal() * 10.0 ** rng.integers(-5, 6):.8f}")
        elif k == 1:
            toks.append(repr(float(rng.standard_normal() * 10.0 ** rng.integers(-30, 30))))
        elif k == 2:
            toks.append(f"{rng.standard_normal():.17e}")
        elif k == 3:
            toks.append(str(rng.integers(-10 ** 9, 10 ** 9)))
        elif k == 4:
            toks.append(f"{rng.standard_normal() * 1e-5:.12f}")
        else:
            toks.append(f"{rng.random():.20f}")
    path = tmp_path / "t.dat"
    with open(path, "w") as f:
        for i in range(0, len(toks), 4):
            f.write(f"{i // 4 + 1} " + " ".join(toks[i:i + 4]) + "\n")
    n = lib.pfem_host_read_table(str(path).encode(), 5, None, 0)
    out = np.zeros((5, n))
    n = lib.pfem_host_read_table(str(path).encode(), 5, out.ctypes.data_as(C.POINTER(C.c_double)), n)
    assert n == len(toks) // 4
    assert np.array_equal(out[1:, :n].T.ravel(), np.array([float(t) for t in toks]))
