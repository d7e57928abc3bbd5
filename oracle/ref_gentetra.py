"""The reference's own mesh generator, compiled here (oracle/_ref/genTetranovtk from /root/reference/src/genTetranovtk.cpp
by `make -C oracle ref`).  TEST INFRASTRUCTURE ONLY: run by tests/ to check `mesh.gen_tetra` (host) and
`pfem_gpu_gen_tetra` (GPU) against the real thing.  The binary travels to the GPU box; the reference tree does not."""
from __future__ import annotations

import hashlib
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BINARY = os.path.join(HERE, "_ref", "genTetranovtk")
REF_SOURCE = os.path.join(os.environ.get("PFEM_REFERENCE", "/root/reference"), "src", "genTetranovtk.cpp")


def build() -> str | None:
    """compile when the reference tree is present; returns the binary path or None."""
    if os.path.exists(REF_SOURCE):
        subprocess.run(["make", "-C", HERE, "ref", "REF=" + os.path.dirname(os.path.dirname(REF_SOURCE))], check=True,
                       stdout=subprocess.DEVNULL)
    return BINARY if os.path.exists(BINARY) else None


def available() -> bool:
    return os.path.exists(BINARY)


def run(x0, x1, nEx, y0, y1, nEy, z0, z1, nEz, keep_text=False):
    """runs the reference generator; returns dict(coords [3,nNode], conn [4,nElem], dbc_node, dbc_dof, sha256 of the
    nodes and elems files).  The DirichBC VALUES of this VTK-free variant are not meaningful (it evaluates an
    uninitialised `coord`, genTetranovtk.cpp:399-404); the node list is."""
    args = [repr(float(v)) if isinstance(v, float) else str(v) for v in (x0, x1, nEx, y0, y1, nEy, z0, z1, nEz)]
    with tempfile.TemporaryDirectory() as d:
        subprocess.run([BINARY] + args, cwd=d, check=True, stdout=subprocess.DEVNULL)
        out = {}
        for key in ("nodes", "elems", "DirichBC"):
            with open(os.path.join(d, f"mesh-{key}.dat"), "rb") as f:
                raw = f.read()
            out["sha_" + key] = hashlib.sha256(raw).hexdigest()
            if keep_text:
                out["text_" + key] = raw
        nodes = np.loadtxt(os.path.join(d, "mesh-nodes.dat"), ndmin=2)
        elems = np.loadtxt(os.path.join(d, "mesh-elems.dat"), dtype=np.int64, ndmin=2)
        dbc = np.loadtxt(os.path.join(d, "mesh-DirichBC.dat"), ndmin=2)
    out["coords"] = np.ascontiguousarray(nodes[:, 1:].T)
    out["conn"] = np.ascontiguousarray(elems[:, 1:].T.astype(np.int32))
    out["dbc_node"] = dbc[:, 0].astype(np.int32)
    out["dbc_dof"] = dbc[:, 1].astype(np.int32)
    return out


def mesh_text(coords, conn):
    """the generator's text format (fixed, 8 decimals, tab separated) for a host / GPU generated mesh."""
    n = coords.shape[1]
    nodes = "".join(f"{i + 1}\t{coords[0, i]:.8f}\t{coords[1, i]:.8f}\t{coords[2, i]:.8f}\n" for i in range(n))
    elems = "".join(f"{e + 1}\t{conn[0, e]}\t{conn[1, e]}\t{conn[2, e]}\t{conn[3, e]}\n" for e in range(conn.shape[1]))
    return nodes.encode(), elems.encode()
