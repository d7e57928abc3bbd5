"""The drop-in, executed: the reference's own PROGRAM text with the INTEGRATION.md diff applied, run by oracle/refrun against
a `SolverB200` backend.  TEST INFRASTRUCTURE ONLY.

`tetrapoissonparallelimpl1.F` is read from the reference tree, the edits of INTEGRATION.md ("The diff a maintainer applies")
are applied to the text in memory (EDITS below: module USE, the solver TYPE, `create`, the pattern loop -> set_mesh +
set_pattern, the value loop -> set_applied + assemble, VecScatter / VecGetArray -> get_solution), and the program is
executed.  Everything the diff does not touch -- argument handling, file reading, numbering, `initialise`, `setZero`,
`factoriseAndSolve`, the temp.dat loop, deallocation, `free` -- is the reference's own statements.

`Module_SolverB200` (include/pfem_b200.f90, an ISO_C_BINDING module no compiler here can build) is mirrored by `Bridge`:
same type-bound procedure names, same C entry points, arguments passed the way Fortran passes them (column-major arrays =
SoA).  The backend behind it is pluggable: tests/test_gpu_zzzz_reference_vectors.py plugs in the real ctypes binding of
libpfemb200.so (GPU); tests/test_refrun_dropin.py plugs in a stand-in built on the oracle to check the plumbing on the CPU.
"""
from __future__ import annotations

import numpy as np

from . import fortran_to_py as F
from . import mocks
from . import run_reference as R
from .runtime import Runtime, _rt

PFEM_POISSON_TRIA, PFEM_POISSON_TETRA, PFEM_ELASTICITY_TRIA, PFEM_ELASTICITY_TETRA = 0, 1, 2, 3

# (first line, last line, replacement) on /root/reference/src/tetrapoissonparallelimpl1.F, 1-based inclusive
EDITS = {
    'tetrapoissonparallelimpl1.F': [
        (26, 26, ["      USE Module_SolverB200"]),
        (107, 107, ["      TYPE(B200Solver) :: solverpetsc",
                    "      DOUBLE PRECISION, DIMENSION(:), ALLOCATABLE :: soln_b200",
                    "      INTEGER :: ierr"]),
        (175, 175, ["      call MPI_Comm_rank(PETSC_COMM_WORLD, this_mpi_proc, errpetsc);",
                    "      call solverpetsc%create(0, this_mpi_proc, n_mpi_procs)"]),
        # pattern pass: LoopElem of MatSetValues(INSERT_VALUES)
        (791, 802, ["      ierr = pfem_solver_set_mesh(solverpetsc%h, PFEM_POISSON_TETRA,",
                    "     1   nElem_global, elemNodeConn, nNode_global, coords,",
                    "     2   node_map_get_old)",
                    "      ierr = pfem_solver_set_pattern(solverpetsc%h, nElem_global,",
                    "     1   nsize, ElemDofArray)"]),
        # value pass: element loop, MatSetValues(ADD), lifting, VecSetValues(ADD)
        (828, 884, ["      ierr = pfem_solver_set_applied(solverpetsc%h, solnApplied,",
                    "     1   nNode_global*ndof)",
                    "      call solverpetsc%assemble(elemData, timeData)"]),
        # VecScatterCreateToAll ... VecGetArray
        (922, 932, ["      ALLOCATE(soln_b200(size_global))",
                    "      ierr = pfem_solver_get_solution(solverpetsc%h, soln_b200)"]),
        (938, 938, ["          fact = soln_b200(ii)"]),
        (968, 968, ["      DEALLOCATE(soln_b200)"]),
    ],
}
# the 3-D elasticity driver: the same edits, plus the ForceBC loop with VecSetValue -> pfem_solver_add_value
EDITS['tetraelasticityparallelimpl1.F'] = [
    (24, 24, ["      USE Module_SolverB200"]),
    (111, 111, ["      TYPE(B200Solver) :: solverpetsc",
                "      DOUBLE PRECISION, DIMENSION(:), ALLOCATABLE :: soln_b200",
                "      INTEGER :: ierr"]),
    (174, 174, ["      call MPI_Comm_rank(PETSC_COMM_WORLD, this_mpi_proc, errpetsc);",
                "      call solverpetsc%create(0, this_mpi_proc, n_mpi_procs)"]),
    (862, 874, ["      ierr = pfem_solver_set_mesh(solverpetsc%h, PFEM_ELASTICITY_TETRA,",
                "     1   nElem_global, elemNodeConn, nNode_global, coords,",
                "     2   node_map_get_old)",
                "      ierr = pfem_solver_set_pattern(solverpetsc%h, nElem_global,",
                "     1   nsize, ElemDofArray)"]),
    (906, 965, ["      ierr = pfem_solver_set_applied(solverpetsc%h, solnApplied,",
                "     1   nNode_global*ndof)",
                "      call solverpetsc%assemble(elemData, timeData)"]),
    (979, 980, ["        ierr = pfem_solver_add_value(solverpetsc%h, row, fact)"]),
    (1019, 1029, ["      ALLOCATE(soln_b200(ntotdofs_global))",
                  "      ierr = pfem_solver_get_solution(solverpetsc%h, soln_b200)"]),
    (1035, 1035, ["          fact = soln_b200(ii)"]),
    (1077, 1077, ["      DEALLOCATE(soln_b200)"]),
]

# what the edited lines must contain today (the reference tree is read-only; this guards the line numbers)
ANCHORS = {
    'tetrapoissonparallelimpl1.F': {26: 'USE Module_SolverPetsc', 107: 'TYPE(PetscSolver) :: solverpetsc',
                                    175: 'MPI_Comm_rank', 791: 'LoopElem: DO ee=1, nElem_global', 802: 'END DO LoopElem',
                                    828: 'DO ee=1, nElem_global', 884: 'END DO', 922: 'VecScatterCreateToAll',
                                    932: 'VecGetArray', 938: 'fact = xx_v(xx_i+ii)', 968: 'VecRestoreArray'},
    'tetraelasticityparallelimpl1.F': {24: 'USE Module_SolverPetsc', 111: 'TYPE(PetscSolver) :: solverpetsc',
                                       174: 'MPI_Comm_rank', 862: 'LoopElem: DO ee=1, nElem_global', 874: 'END DO LoopElem',
                                       906: 'DO ee=1, nElem_global', 965: 'END DO', 979: 'call VecSetValue(solverpetsc%rhsVec, row, fact',
                                       980: 'ADD_VALUES, errpetsc)', 1019: 'VecScatterCreateToAll', 1029: 'VecGetArray',
                                       1035: 'fact = xx_v(xx_i+ii)', 1077: 'VecRestoreArray'},
}


def patched_source(driver_file: str) -> str:
    with open(R.os.path.join(R.REF_SRC, driver_file)) as f:
        lines = f.read().split('\n')
    for no, text in ANCHORS[driver_file].items():
        assert text in lines[no - 1], (no, lines[no - 1])
    for first, last, new in sorted(EDITS[driver_file], reverse=True):
        lines[first - 1:last] = new
    return '\n'.join(lines)


class Bridge:
    """TYPE B200Solver of include/pfem_b200.f90, procedure for procedure, over a python backend that has the methods of
    pfemfort_b200.solver.SolverB200."""
    backend = None          # set by run(): a callable (device, rank, nranks) -> solver object
    created = None

    def __init__(self):
        self.h = None
        self.captured = {}

    # type-bound procedures (Fortran names are case-insensitive: the translator lower-cases them)
    def create(self, device, rank, nranks, id128=None):
        self.h = Bridge.backend(int(device.v), int(rank.v), int(nranks.v))
        Bridge.created.append(self)

    def initialise(self, size_local, size_global, diag_nnz, offdiag_nnz):
        self.h.initialise(int(size_local.v), int(size_global.v), np.ascontiguousarray(diag_nnz, np.int32),
                          np.ascontiguousarray(offdiag_nnz, np.int32))

    def setzero(self):
        self.h.setZero()

    def assemble(self, elemdata, timedata):
        # the module hands the C side the address of elemData(50) / timeData(50); the kernels read the leading entries only.
        # Entries the PROGRAM never sets are NaN in this run-time; in a compiled run they are the zero-filled static
        # storage of a main program, which is what crosses the boundary here.
        self.h.assemble([float(v) for v in np.nan_to_num(elemdata[:8], nan=0.0)],
                        [float(v) for v in np.nan_to_num(timedata[:8], nan=0.0)])

    def factoriseandsolve(self):
        self.h.factoriseAndSolve()
        rp, col, val = self.h.get_csr()
        self.captured = dict(rowptr=np.array(rp), col=np.array(col), val=np.array(val), rhs=np.array(self.h.get_rhs()),
                             info=dict(self.h.info()))

    def free(self):
        self.h.free()


def _soa_i32(a, rows):
    """Fortran (rows, k) column-major array = SoA [k][rows] on the wire."""
    return np.ascontiguousarray(np.asarray(a).T[:, :rows], dtype=np.int32)


def _bridge_namespace():
    def pfem_solver_set_mesh(h, kind, nelem, conn, nnode, coords, old):
        h.set_mesh(int(kind), _soa_i32(conn, int(nelem)), np.ascontiguousarray(np.asarray(coords).T[:, :int(nnode)]),
                   np.ascontiguousarray(old, np.int32))
        return 0

    def pfem_solver_set_pattern(h, nelem, nsize, edof):
        e = _soa_i32(edof, int(nelem))
        assert e.shape[0] == int(nsize)
        h.set_pattern(e)
        return 0

    def pfem_solver_set_applied(h, applied, n):
        h.set_applied(np.ascontiguousarray(applied, np.float64)[:int(n)])
        return 0

    def pfem_solver_get_solution(h, out):
        out[...] = h.get_solution()
        return 0

    return {'pfem_solver_set_mesh': pfem_solver_set_mesh, 'pfem_solver_set_pattern': pfem_solver_set_pattern,
            'pfem_solver_set_applied': pfem_solver_set_applied, 'pfem_solver_get_solution': pfem_solver_get_solution,
            'pfem_poisson_tria': PFEM_POISSON_TRIA, 'pfem_poisson_tetra': PFEM_POISSON_TETRA,
            'pfem_elasticity_tria': PFEM_ELASTICITY_TRIA, 'pfem_elasticity_tetra': PFEM_ELASTICITY_TETRA,
            '_new_b200solver': Bridge}


def program_path(driver_file):
    return R.os.path.join(R.OUT_DIR, 'dropin_' + R.os.path.splitext(driver_file)[0] + '.py')


def build_program(driver_file='tetrapoissonparallelimpl1.F'):
    """translate the edited PROGRAM (needs the reference tree) into oracle/_ref/, where it travels to the GPU box with the
    snapshot like the other oracle/_ref/ artefacts (git-ignored: it is derived from the reference's source)."""
    sources = R.read_sources(R.ELEMENT_FILES, intent=True)
    sources[driver_file] = patched_source(driver_file)
    code = F.translate(sources, {'vecgetarray': mocks.vecgetarray_rewrite})
    R.os.makedirs(R.OUT_DIR, exist_ok=True)
    with open(program_path(driver_file), 'w') as f:
        f.write(code)
    return program_path(driver_file)


def available(driver_file='tetrapoissonparallelimpl1.F'):
    return R.available() or R.os.path.exists(program_path(driver_file))


def run(driver_file, argv, backend, cwd='.'):
    """one rank: execute the edited PROGRAM against `backend`.  Returns (Bridge object with .captured, Runtime).  Where the
    reference tree is present the program is translated afresh; elsewhere the prebuilt oracle/_ref/ file is used."""
    path = build_program(driver_file) if R.available() else program_path(driver_file)
    with open(path) as f:
        code = f.read()
    world = mocks.World(1)
    rt = Runtime([driver_file] + list(argv), cwd, 0, world, True)
    _rt.bind(rt)
    Bridge.backend, Bridge.created = backend, []
    ns = dict(mocks.namespace())
    ns.update(_bridge_namespace())
    exec(compile(code, path, 'exec'), ns)
    prog = [k for k in ns if k.startswith('program_')]
    assert len(prog) == 1
    ns[prog[0]]()
    assert len(Bridge.created) == 1
    return Bridge.created[0], rt


# ---- the same, through the real Fortran module ------------------------------------------------------------------------

MODULE_FILE = R.os.path.join(R.os.path.dirname(R.os.path.dirname(R.os.path.dirname(R.os.path.abspath(__file__)))), 'include',
                             'pfem_b200.f90')


def module_program_path(driver_file):
    return R.os.path.join(R.OUT_DIR, 'dropin_f90_' + R.os.path.splitext(driver_file)[0] + '.py')


def build_module_program(driver_file='tetrapoissonparallelimpl1.F'):
    """the edited PROGRAM + include/pfem_b200.f90 itself (not the python Bridge), translated into oracle/_ref/."""
    sources = R.read_sources(R.ELEMENT_FILES, intent=True)
    with open(MODULE_FILE) as f:
        sources['pfem_b200.f90'] = f.read()
    sources[driver_file] = patched_source(driver_file)
    code = F.translate(sources, {'vecgetarray': mocks.vecgetarray_rewrite}, static_zero_programs=True)
    R.os.makedirs(R.OUT_DIR, exist_ok=True)
    with open(module_program_path(driver_file), 'w') as f:
        f.write(code)
    return module_program_path(driver_file)


def module_available(driver_file='tetrapoissonparallelimpl1.F'):
    return R.available() or R.os.path.exists(module_program_path(driver_file))


def run_through_module(driver_file, argv, clib, cwd='.', before_free=None):
    """execute the edited PROGRAM with `Module_SolverB200` = the translated include/pfem_b200.f90, its BIND(C) interfaces bound
    to `clib` (a ctypes.CDLL: the real libpfemb200.so, or the test double of tests/fake_abi).  `before_free(handle)` is
    called just before the PROGRAM's own `call solverpetsc%free()` releases the handle.  Returns the Runtime."""
    from .runtime import set_clib
    path = build_module_program(driver_file) if R.available() else module_program_path(driver_file)
    with open(path) as f:
        code = f.read()
    world = mocks.World(1)
    rt = Runtime([driver_file] + list(argv), cwd, 0, world, True)
    _rt.bind(rt)
    set_clib(clib)
    ns = dict(mocks.namespace())
    exec(compile(code, path, 'exec'), ns)
    if before_free is not None:
        real_free = ns['pfem_solver_free']

        def free_hook(h):
            before_free(h.v if hasattr(h, 'v') else h)
            return real_free(h)
        ns['pfem_solver_free'] = free_hook
    prog = [k for k in ns if k.startswith('program_')]
    assert len(prog) == 1
    ns[prog[0]]()
    return rt
