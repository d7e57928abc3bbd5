"""include/pfem_b200.f90 (the ISO_C_BINDING module a PFEMFort driver would USE) has never met a Fortran compiler: none
exists in this image.  What can be done here, and is: (1) the module is parsed and translated by oracle/refrun's Fortran
front end (free form, INTERFACE blocks, BIND(C), VALUE, OPTIONAL / PRESENT, TYPE with initialised components and bound
procedures); (2) every BIND(C) interface is cross-checked against the prototype of the same name in include/pfem_b200.h --
argument count, by-value vs by-address, C type -- which is exactly the class of mistake neither a Fortran compiler nor the
linker would catch; (3) the translated module is EXECUTED against the real libpfemb200.so: on a box without a GPU the
library answers PFEM_ERR_CUDA and the module's `check` STOPs, on the GPU box the reference's own PROGRAM runs through it
(tests/test_gpu_zzzz_reference_vectors.py)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle.refrun import fortran_to_py as F
from oracle.refrun.runtime import FortranStop, Ref, Runtime, _rt, set_clib
from pfemfort_b200 import solver as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODULE = os.path.join(ROOT, "include", "pfem_b200.f90")
HEADER = os.path.join(ROOT, "include", "pfem_b200.h")


@pytest.fixture(scope="module")
def parsed():
    with open(MODULE) as f:
        return F.parse_file(f.read(), free_form=True)


def c_prototypes():
    with open(HEADER) as f:
        text = re.sub(r"/\*.*?\*/", " ", f.read(), flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for m in re.finditer(r"\b(int|long long)\s+(pfem_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = []
        for a in [x.strip() for x in m.group(3).replace("\n", " ").split(",")]:
            if a in ("void", ""):
                continue
            by_address = "*" in a or "[" in a
            base = "double" if "double" in a else "longlong" if "long long" in a else "int" if re.search(r"\bint\b", a) \
                else "char" if "char" in a else "handle" if re.search(r"pfem_\w+_t", a) else "void" if "void" in a else a
            args.append((base, by_address, a.count("*")))
        protos[m.group(2)] = args
    return protos


def test_module_parses_and_mirrors_the_petscsolver_procedures(parsed):
    units, typedefs, params, interfaces = parsed
    assert set(typedefs) == {"b200solver"}
    procs = set(typedefs["b200solver"].procs)
    # TYPE PetscSolver, solverpetsc.F:94-103
    assert {"initialise", "setzero", "free", "printinfo", "assemblematrix", "assemblevector", "assemblematrixandvector",
            "factorise", "solve", "factoriseandsolve"} <= procs
    assert {"create", "assemble", "setoptionsfromfile"} <= procs
    assert {u.name for u in units} == procs | {"check"}
    assert typedefs["b200solver"].fields == {"h": "h", "ierr": "i"}
    assert params["pfem_pc_bjacobi_ilu0"].init == "2" and params["pfem_poisson_tetra"].init == "1"
    assert len(interfaces) >= 40


def test_every_interface_matches_its_c_prototype(parsed):
    units, typedefs, params, interfaces = parsed
    protos = c_prototypes()
    lib = S.load_library()
    for name, itf in interfaces.items():
        assert name in protos, f"{name}: no prototype in pfem_b200.h"
        assert hasattr(lib, name), f"{name}: not exported by libpfemb200.so"
        cargs = protos[name]
        assert len(cargs) == len(itf.args), f"{name}: {len(itf.args)} dummies, {len(cargs)} C parameters"
        for dn, (base, by_address, stars) in zip(itf.args, cargs):
            d = itf.syms[dn]
            where = f"{name}({dn})"
            if base == "handle":
                assert d.typ == "h", where
                assert d.value == (stars == 1), where            # T* by VALUE, T** by reference
                continue
            assert d.value == (not by_address), f"{where}: VALUE = {d.value}, C passes by {'address' if by_address else 'value'}"
            want = {"int": "i", "double": "d", "char": "c", "void": "c"}[base]
            assert d.typ == want, f"{where}: Fortran kind {d.typ}, C type {base}"
            if by_address and base in ("char", "void"):
                assert d.dims is not None, where


def test_the_c_constants_agree_with_the_header(parsed):
    units, typedefs, params, interfaces = parsed
    with open(HEADER) as f:
        h = f.read()
    for name, sym in params.items():
        m = re.search(r"\b%s\s*=\s*(-?\d+)" % name.upper(), h) or re.search(r"#define\s+%s\s+(-?\d+)" % name.upper(), h)
        assert m, f"{name.upper()} is not in pfem_b200.h"
        assert int(m.group(1)) == int(sym.init), name


def test_translated_module_calls_the_real_library(parsed):
    """On a box without a GPU: B200Solver%create marshals its arguments into pfem_solver_create of the real shared library,
    gets the library's error status back, and `check` STOPs like the module says.  With a GPU it succeeds and frees."""
    with open(MODULE) as f:
        code = F.translate({"pfem_b200.f90": f.read()})
    ns = {}
    exec(compile(code, "<pfem_b200.f90>", "exec"), ns)
    set_clib(S.load_library())
    _rt.bind(Runtime())
    solver = ns["_new_b200solver"]()
    assert solver.h is None
    if S.device_count() > 0:
        solver.create(Ref(0), Ref(0), Ref(1))
        assert solver.h
        solver.free()
        assert solver.h is None
    else:
        with pytest.raises(FortranStop, match="Aborting... in Module_SolverB200"):
            solver.create(Ref(0), Ref(0), Ref(1))
        assert solver.h is None
        assert _rt.stdout and "create" in str(_rt.stdout[-1])
        # an element routine through its interface: column-major K comes back by address
        x, y = np.array([0.0, 1.0, 0.0]), np.array([0.0, 0.0, 1.0])
        K, Fl = np.zeros((3, 3), order="F"), np.zeros(3)
        rc = ns["pfem_poisson_tria_ke"](x, y, np.ones(8), np.array([0.0, 1.0, 0.0]), np.zeros(3), np.zeros(3), K, Fl)
        assert rc != 0 and b"CUDA" in ctypes.cast(S.load_library().pfem_last_error(), ctypes.c_char_p).value.upper()
