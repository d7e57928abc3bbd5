"""The C-ABI library loads on a CPU-only box and exports every symbol include/pfem_b200.h declares; compute entry
points fail loudly (PFEM_ERR_CUDA) without a GPU instead of falling back to the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from pfemfort_b200 import solver as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "pfem_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pfem_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = S.load_library()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pfem_b200.h but not exported"


def test_host_side_exports():
    lib = S.load_library()
    for n in ("pfem_host_partition_mesh", "pfem_host_number_dofs", "pfem_host_renumber_conn", "pfem_host_elem_dof_array",
              "pfem_host_select_elements", "pfem_host_gather_rows", "pfem_host_read_table", "pfem_host_mesh_read_binary_header",
              "pfem_host_mesh_read_binary", "pfem_host_mesh_write_binary"):
        assert hasattr(lib, n)


def test_no_cpu_fallback_without_a_gpu():
    if S.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(S.PfemError) as ei:
        S.SolverB200(0)
    assert ei.value.status == S.ERR_CUDA
    with pytest.raises(S.PfemError) as ei:
        S.element_ke(S.POISSON_TRIA, [0, 1, 0], [0, 0, 1], None, [1, 1], [0, 1, 0])
    assert ei.value.status == S.ERR_CUDA


def test_product_does_not_link_or_import_the_oracle():
    import subprocess
    out = subprocess.run(["ldd", S.LIBPATH], capture_output=True, text=True).stdout
    assert "liborc" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "pfemfort_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "pyoracle" not in src and "liborc" not in src and "orc_" not in src, f


def test_cpp_driver_fails_loudly_without_a_gpu(tmp_path):
    """The compiled C++ driver links against libpfemb200.so only and refuses to run without a device."""
    import gzip
    import shutil
    import subprocess
    if S.device_count() > 0:
        pytest.skip("a GPU is visible")
    drv = os.path.join(ROOT, "pfemfort_b200", "bin", "pfem_driver")
    assert os.path.exists(drv), "run __graft_entry__.build()"
    out = subprocess.run(["ldd", drv], capture_output=True, text=True).stdout
    assert "libpfemb200.so" in out and "liborc" not in out
    files = []
    for part in ("nodes", "elems", "DirichBC"):
        dst = os.path.join(str(tmp_path), f"tet10-{part}.dat")
        with gzip.open(os.path.join(ROOT, "tests", "golden", "input", f"tet10-{part}.dat.gz"), "rb") as f, open(dst, "wb") as g:
            shutil.copyfileobj(f, g)
        files.append(dst)
    r = subprocess.run([drv, "tetrapoisson"] + files, cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CUDA device" in r.stderr
    assert "Total DOF = 729" in r.stdout          # the host side (reading, numbering) ran before the GPU was needed
