"""Build libpfemb200.so (sm_100a only) in-tree with nvcc.

The shared library is the product: CUDA kernels + the C ABI of include/pfem_b200.h.  It is built next to
this file so that it travels with the repository snapshot to the GPU box (a JIT cache would not).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpfemb200.so")
OBJDIR = os.path.join(HERE, "_build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
          "-Xptxas", "-v", "--expt-relaxed-constexpr"]
# element arithmetic must not be contracted into FMAs (bit-faithful to the reference's evaluation order)
NO_FMA = {"elements.cu", "assembly.cu", "assembly_tiled.cu", "explicit.cu"}
# host-side set-up loops (tile construction) use OpenMP
OPENMP = {"assembly_tiled.cu"}
SOURCES = ["api.cu", "elements.cu", "pattern.cu", "assembly.cu", "assembly_tiled.cu", "assembly_ctile.cu", "assembly_fast.cu", "cg.cu", "comm.cu", "explicit.cu", "gpu_setup.cu", "host_driver.cu",
           "host_meshio.cu"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libpfemb200.so cannot be built (there is no CPU fallback)")
    return nvcc


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "pfem_b200.h"),
                                                               os.path.join(HERE, "..", "drivers", "pfem_driver.cpp"),
                                                               os.path.abspath(__file__)]
    stamp = os.path.join(OBJDIR, "stamp")
    extra = os.environ.get("PFEM_EXTRA_NVCC", "").split()      # e.g. -DPFEM_PCG_TRACE (tools/pcg_trace.py); part of the stamp
    dig = _digest(deps) + "|" + " ".join(extra)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()
    objs = []
    log = []
    cmds = []
    for src in SOURCES:
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if src in NO_FMA:
            cmd.insert(1, "-fmad=false")
        if src in OPENMP:
            cmd[1:1] = ["-Xcompiler", "-fopenmp"]
        cmds.append((src, cmd))
        objs.append(obj)
    # translation units are independent: compile them side by side
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda sc: (sc[0], sc[1], subprocess.run(sc[1], capture_output=True, text=True)), cmds))
    for src, cmd, r in results:
        log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            sys.stderr.write(log[-1])
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-Xlinker", "--exclude-libs,ALL", "-ldl", "-Xcompiler", "-fopenmp",
           "-L/usr/local/cuda/targets/x86_64-linux/lib", "-lmetis_static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError("link failed")
    # the C++ counterpart of the Fortran driver PROGRAMs, linked against the shared library only
    os.makedirs(os.path.join(HERE, "bin"), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-o", os.path.join(HERE, "bin", "pfem_driver"),
           os.path.join(HERE, "..", "drivers", "pfem_driver.cpp"), "-L" + HERE, "-lpfemb200", "-Wl,-rpath," + HERE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError("driver link failed")
    with open(os.path.join(OBJDIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    with open(stamp, "w") as f:
        f.write(dig)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
