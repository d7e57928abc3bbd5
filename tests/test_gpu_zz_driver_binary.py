"""The C++ driver on the PFEMB1 binary container (csrc/host_meshio.cu): same result as on the reference's text files.
(Sorted after the established GPU parity files: added after the last GPU session of round 1.)"""
import os
import subprocess

import pytest

from pfemfort_b200 import mesh as M
from test_gpu_driver_cpp import DRIVER, _unpack

pytestmark = pytest.mark.gpu


def test_cpp_driver_binary_container_input(gpu, input_dir, tmp_path):
    """pfem_driver <physics> mesh.pfemb: same temp.dat, byte for byte, as with the three text files."""
    files = _unpack("tet10", input_dir, str(tmp_path))
    env = dict(os.environ, PFEM_KSP_RTOL="1e-10")
    r = subprocess.run([DRIVER, "tetrapoisson"] + files, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    text_out = open(os.path.join(str(tmp_path), "temp.dat")).read()
    os.remove(os.path.join(str(tmp_path), "temp.dat"))
    pfemb = os.path.join(str(tmp_path), "tet10.pfemb")
    M.write_binary(M.read_mesh(os.path.join(input_dir, "tet10")), pfemb)
    r = subprocess.run([DRIVER, "tetrapoisson", pfemb], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert open(os.path.join(str(tmp_path), "temp.dat")).read() == text_out
