#!/usr/bin/env python
"""Golden vectors produced by RUNNING THE REFERENCE'S OWN SOURCE (needs /root/reference; run in the build container):

    python tests/golden/make_reference_vectors.py

The reference is Fortran + PETSc + MPI + METIS and no Fortran compiler exists in the image, so its sources are executed
through oracle/refrun (a Fortran-subset -> Python translator with Fortran's typing / rounding rules; PETSc, MPI, METIS
and the VTK writer, which are third-party / outside the path, come from oracle/refrun/mocks.py).  What is written:

  ref_elements.npz      the eight element routines of the path (StiffnessResidual{Poisson,Elasticity}Linear{Tria,Tetra},
                        ResidualElasticityLinear{Tria,Tetra}, MassMatrixLinear{Tria,Tetra}) on seeded random elements:
                        inputs, outputs, and which elements the routine STOPs on (negative Jacobian)
  ref_driver_<case>.npz the four `*parallelimpl1` PROGRAMs run end to end on the bundled inputs, on 1 and on P simulated
                        ranks: numbering arrays, the matrix / right-hand side PETSc was handed (CSR, after
                        MatAssemblyEnd), the solver options the wrapper set, the temp.dat records

Documented-intent substitutions (oracle/refrun/run_reference.py INTENT_PATCHES) are applied where the shipped text
cannot run; `shipped_*` entries record what the shipped text does instead (it STOPs).  tests/test_reference_vectors.py
compares the oracle with these files on the CPU, tests/test_gpu_zzzz_reference_vectors.py the CUDA path on the GPU, and
(when /root/reference is present) re-runs a subset of this script and requires identical output.
"""
import gzip
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.refrun import run_reference as R            # noqa: E402
from oracle.refrun.runtime import FortranStop           # noqa: E402

KIND_NAMES = ['poisson_tria', 'poisson_tetra', 'elasticity_tria', 'elasticity_tetra']
KE = ['stiffnessresidualpoissonlineartria', 'stiffnessresidualpoissonlineartetra',
      'stiffnessresidualelasticitylineartria', 'stiffnessresidualelasticitylineartetra']
RES = {2: 'residualelasticitylineartria', 3: 'residualelasticitylineartetra'}
MASS = {2: 'massmatrixlineartria', 3: 'massmatrixlineartetra'}
DIMS = {0: (3, 1, 2), 1: (4, 1, 3), 2: (3, 2, 2), 3: (4, 3, 3)}     # npe, ndof, ndim


def random_elements(kind, n, seed):
    """n elements: mostly positive Jacobian, some negative, over four coordinate scales / offsets."""
    npe, ndof, ndim = DIMS[kind]
    rng = np.random.default_rng(seed)
    xyz = np.zeros((n, 3, npe))
    for e in range(n):
        scale = [1.0, 1e-3, 37.0, 1.0][e % 4]
        offset = [0.0, 0.0, 0.0, 1e3][e % 4]
        p = rng.standard_normal((npe, ndim)) * scale + offset * rng.standard_normal(ndim)
        xyz[e, :ndim, :] = p.T
    ed = np.zeros((n, 50))
    ed[:, 0] = rng.uniform(0.5, 300.0, n)
    ed[:, 1] = rng.uniform(0.05, 0.45, n)
    ed[:, 2] = rng.uniform(0.5, 2.0, n)
    ed[:, 3:6] = rng.standard_normal((n, 3))
    td = np.zeros((n, 50))
    td[:, 1] = rng.uniform(0.5, 1.0, n)
    td[:, 2] = rng.uniform(0.0, 1.0, n)
    val = rng.standard_normal((n, npe * ndof))
    val[: n // 4] = 0.0                                  # the implicit drivers pass valC = 0
    return xyz, ed, td, val


def _call(fn, kind, xyz, *rest):
    npe, ndof, ndim = DIMS[kind]
    coords = [np.array(xyz[d], order='F') for d in range(ndim)]
    try:
        fn(*coords, *rest)
        return 0
    except FortranStop as ex:
        assert 'Negative Jacobian' in ex.msg, ex
        return 1


def make_elements(path, n=96):
    ns = R.element_routines(intent=True)
    out = {}
    for kind in range(4):
        npe, ndof, ndim = DIMS[kind]
        nsz = npe * ndof
        xyz, ed, td, val = random_elements(kind, n, 1000 + kind)
        K = np.zeros((n, nsz, nsz))
        Fv = np.zeros((n, nsz))
        neg = np.zeros(n, np.int32)
        for e in range(n):
            Kl = np.full((nsz, nsz), np.nan, order='F')
            Fl = np.full(nsz, np.nan)
            neg[e] = _call(ns[KE[kind]], kind, xyz[e], ed[e].copy(), td[e].copy(), val[e].copy(), np.zeros(nsz), Kl, Fl)
            if not neg[e]:
                K[e], Fv[e] = Kl, Fl                      # K[e][i, j] = Klocal(i+1, j+1)
        assert not np.isnan(K).any() and not np.isnan(Fv).any()
        out.update({f'ke{kind}_xyz': xyz, f'ke{kind}_ed': ed[:, :8], f'ke{kind}_td': td[:, :4], f'ke{kind}_valc': val,
                    f'ke{kind}_K': K, f'ke{kind}_F': Fv, f'ke{kind}_neg': neg})
        if kind in RES:
            Fr = np.zeros((n, nsz))
            Ml = np.zeros((n, nsz))
            for e in range(n):
                if neg[e]:
                    continue
                Fl = np.full(nsz, np.nan)
                assert _call(ns[RES[kind]], kind, xyz[e], ed[e].copy(), td[e].copy(), val[e].copy(), np.zeros(nsz), Fl) == 0
                Fr[e] = Fl
                Mv = np.full(nsz, np.nan)
                assert _call(ns[MASS[kind]], kind, xyz[e], ed[e].copy(), Mv) == 0
                Ml[e] = Mv
            assert not np.isnan(Fr).any() and not np.isnan(Ml).any()
            out.update({f'res{kind}_F': Fr, f'mass{kind}_M': Ml})
    # what the SHIPPED 3-D elasticity text does: computeBasisFunctions3D is called with ETYPE = 1 and STOPs
    shipped = R.element_routines(intent=False)
    xyz, ed, td, val = random_elements(3, 1, 7)
    stops = []
    for name, extra in ((KE[3], (ed[0], td[0], val[0], np.zeros(12), np.zeros((12, 12), order='F'), np.zeros(12))),
                        (RES[3], (ed[0], td[0], val[0], np.zeros(12), np.zeros(12))),
                        (MASS[3], (ed[0], np.zeros(12)))):
        try:
            shipped[name](*[np.array(xyz[0][d]) for d in range(3)], *extra)
            stops.append(0)
        except FortranStop as ex:
            assert 'computeBasisFunctions3D' in ex.msg
            stops.append(ex.line)
    out['shipped_elasticity3d_stop_lines'] = np.array(stops, np.int32)
    np.savez_compressed(path, **out)
    return out


# ---- drivers ---------------------------------------------------------------------------------------------------------

CASES = {
    # name: (driver, input prefix, kind, has ForceBC, swap local nodes 3 <-> 4, rank counts)
    'tria20x20': ('triapoissonparallelimpl1.F', 'tria20x20', 0, False, False, (1, 3)),
    'tet10': ('tetrapoissonparallelimpl1.F', 'tet10', 1, False, False, (1, 2, 4, 8)),
    'cookmembranetria32': ('triaelasticityparallelimpl1.F', 'cookmembranetria32', 2, True, False, (1, 2)),
    'beam3Dtet6366': ('tetraelasticityparallelimpl1.F', 'beam3Dtet6366', 3, True, True, (1, 2)),
}


def stage_inputs(workdir):
    src = os.path.join(HERE, 'input')
    for f in os.listdir(src):
        with gzip.open(os.path.join(src, f)) as g, open(os.path.join(workdir, f[:-3]), 'wb') as o:
            o.write(g.read())


def swapped_elems(workdir, prefix):
    """the bundled beam has negative Jacobians under the reference's own basis functions (every element): the
    documented-intent input exchanges local nodes 3 and 4."""
    out = os.path.join(workdir, prefix + '-swap34-elems.dat')
    with open(os.path.join(workdir, prefix + '-elems.dat')) as f, open(out, 'w') as o:
        for line in f:
            t = line.split()
            if t:
                o.write(f"{t[0]}\t{t[1]}\t{t[2]}\t{t[4]}\t{t[3]}\n")
    return os.path.basename(out)


def synthetic_partition(workdir, prefix, nparts, seed=5):
    """A legal METIS answer without METIS: nodes in `nparts` blocks of a seeded random order (so that the reference's
    renumbering really permutes), every element with the part of its first node."""
    nodes = sum(1 for l in open(os.path.join(workdir, prefix + '-nodes.dat')) if l.strip())
    conn = np.array([[int(x) for x in l.split()[1:]] for l in open(os.path.join(workdir, prefix + '-elems.dat')) if l.strip()])
    rng = np.random.default_rng(seed + nparts)
    # spatially coherent blocks with a ragged, shuffled boundary layer
    npid = (np.arange(nodes) * nparts) // nodes
    flip = rng.random(nodes) < 0.15
    npid[flip] = rng.integers(0, nparts, flip.sum())
    for p in range(nparts):
        assert (npid == p).any()
    epid = npid[conn[:, 0] - 1]
    return epid.astype(np.int64), npid.astype(np.int64)


def run_case(name, nranks, workdir):
    drv, prefix, kind, fbc, swap, _ = CASES[name]
    argv = [f'{prefix}-nodes.dat', f'{prefix}-elems.dat', f'{prefix}-DirichBC.dat'] + ([f'{prefix}-ForceBC.dat'] if fbc else [])
    out = {}
    if swap:
        shipped = R.run_driver(drv, argv, 1, cwd=workdir)
        assert shipped.stopped is not None and 'Negative Jacobian' in shipped.stopped.msg
        out['shipped_stop_line'] = np.array([shipped.stopped.line], np.int32)
        argv[1] = swapped_elems(workdir, prefix)
    part = synthetic_partition(workdir, prefix, nranks) if nranks > 1 else None
    res = R.run_driver(drv, argv, nranks, partition=part, cwd=workdir)
    assert res.stopped is None, res.stopped
    rowptr, col, val, rhs = res.system
    fa = res.ranks[0].final_arrays
    out.update(rowptr=rowptr.astype(np.int32), col=col.astype(np.int32), val=val, rhs=rhs,
               NodeDofArrayNew=fa['nodedofarraynew'].astype(np.int32), ElemDofArray=fa['elemdofarray'].astype(np.int32),
               node_map_get_old=fa['node_map_get_old'].astype(np.int32),
               node_map_get_new=fa['node_map_get_new'].astype(np.int32),
               assyForSoln=fa['assyforsoln'].astype(np.int32), solnApplied=fa['solnapplied'],
               elem_proc_id=fa.get('elem_proc_id', fa.get('elem_procid')).astype(np.int32),
               node_proc_id=fa.get('node_proc_id', fa.get('node_procid')).astype(np.int32))
    sl = 'size_local' if 'size_local' in res.ranks[0].final_locals else 'ntotdofs_local'
    info = np.array([[rt.final_locals[k] for k in ('node_start', 'node_end', 'row_start', 'row_end', sl)] for rt in res.ranks],
                    np.int64)
    out['part_info'] = info                      # per rank: node_start, node_end, row_start, row_end (1-based), size_local
    trace = [t for t in res.world.trace if t[0] == 0]
    out['ksp_type'] = np.array([t[2] for t in trace if t[1] == 'KSPSetType'])
    out['pc_type'] = np.array([t[2] for t in trace if t[1] == 'PCSetType'])
    out['options'] = np.array([f'{t[1]}:{t[2]}' for t in trace if t[1] in ('MatSetOption', 'VecSetOption')])
    out['call_order'] = np.array([t[1] for t in trace])
    rec = res.ranks[0].written.get('temp.dat', [])
    if rec and len(rec[0]) == 3:                 # Poisson drivers write (ii, old node, value)
        out['temp_dat_index'] = np.array([[r[0], r[1]] for r in rec], np.int32)
        out['temp_dat_value'] = np.array([r[2] for r in rec])
    else:                                        # elasticity drivers write the value only
        out['temp_dat_value'] = np.array([r[0] for r in rec])
    out['vtk_soln'] = np.asarray(res.world.vtk['soln'], np.float64)
    return out


def make_drivers(outdir, only=None):
    made = {}
    with tempfile.TemporaryDirectory() as wd:
        stage_inputs(wd)
        for name, spec in CASES.items():
            for p in spec[5]:
                tag = f'{name}_p{p}'
                if only and tag not in only:
                    continue
                data = run_case(name, p, wd)
                if outdir:
                    np.savez_compressed(os.path.join(outdir, f'ref_driver_{tag}.npz'), **data)
                made[tag] = data
    return made


# ---- explicit dynamics: the reference's central-difference PROGRAM --------------------------------------------------

EXPLICIT_STEPS = 40


def make_explicit(path, steps=EXPLICIT_STEPS):
    """triaelasticityexplicit.F end to end on the cook membrane, one rank.  The only change to the text is the hard-coded
    run length (`stepsMax = 50000` -> `steps`); material data, dt = 0.0002 (single-precision literal) and the load switch
    are the program's own.  Recorded: the lumped mass, the state after `steps` steps, every solnoutput.dat record."""
    with tempfile.TemporaryDirectory() as wd:
        stage_inputs(wd)
        name = 'cookmembranetria32'
        argv = [f'{name}-nodes.dat', f'{name}-elems.dat', f'{name}-DirichBC.dat', f'{name}-ForceBC.dat']
        res = R.run_driver('triaelasticityexplicit.F', argv, 1, cwd=wd,
                           extra_patches=[('stepsMax = 50000', f'stepsMax = {steps}', 1)])
    assert res.stopped is None, res.stopped
    fa, fl = res.ranks[0].final_arrays, res.ranks[0].final_locals
    out = dict(globalM=fa['globalm'], disp=fa['disp'], dispPrev=fa['dispprev'], dispPrev2=fa['dispprev2'], velo=fa['velo'],
               acce=fa['acce'], elemData=fa['elemdata'][:6], timeData=fa['timedata'][:3], dt=np.float64(fl['dt']),
               steps=np.int32(fl['stepscompleted']), timeNow=np.float64(fl['timenow']),
               solnoutput=np.array(res.ranks[0].written['solnoutput.dat'], np.float64),
               assyForSoln=fa['assyforsoln'].astype(np.int32))
    if path:
        np.savez_compressed(path, **out)
    return out


# ---- TYPE PetscSolver's procedures (solverpetsc.F), driven by a small harness PROGRAM written here -------------------------

SOLVER_HARNESS = """
      PROGRAM SolverProcedures
      USE Module_SolverPetsc
      IMPLICIT NONE
      TYPE(PetscSolver) :: s
      PetscErrorCode errpetsc
      INTEGER :: dn(6), on(6), e1(3), e2(3), e3(3), r2(3), c2(3), r3(3)
      INTEGER :: ii, jj, mode
      DOUBLE PRECISION :: Z(3,3), K1(3,3), K2(3,3), F1(3), F3(3)
      CHARACTER(len=32) :: arg
      CALL getarg(1, arg)
      READ(arg,*) mode
      call PetscInitialize("petsc_options.dat", errpetsc)
      dn = 6; on = 6
      call s%initialise(6, 6, dn, on)
      IF(mode == 1) THEN
        call s%factorise()
      END IF
      IF(mode == 2) THEN
        call s%solve()
      END IF
!     pattern pass, as the drivers do it: zero blocks, INSERT_VALUES
      e1(1) = 0; e1(2) = 1; e1(3) = 2
      e2(1) = 2; e2(2) = 3; e2(3) = 4
      e3(1) = 3; e3(2) = 4; e3(3) = 5
      Z = 0.0
      call MatSetValues(s%mtx, 3, e1, 3, e1, Z, INSERT_VALUES, errpetsc)
      call MatSetValues(s%mtx, 3, e2, 3, e2, Z, INSERT_VALUES, errpetsc)
      call MatSetValues(s%mtx, 3, e3, 3, e3, Z, INSERT_VALUES, errpetsc)
      call s%setZero()
      DO jj=1,3
        DO ii=1,3
          K1(ii,jj) = 10.0d0*ii + jj + 0.125d0
          K2(ii,jj) = -1.0d0*ii + 100.0d0*jj + 0.5d0
        END DO
        F1(jj) = 7.0d0 + jj
        F3(jj) = -3.0d0*jj
      END DO
      call s%assembleMatrixAndVector(e1, e1, K1, F1)
      r2(1) = 2; r2(2) = -1; r2(3) = 4
      c2(1) = 3; c2(2) = 4;  c2(3) = -1
      call s%assembleMatrix(r2, c2, K2)
      r3(1) = 5; r3(2) = -1; r3(3) = 3
      call s%assembleVector(r3, F3)
      call s%assembleMatrixAndVector(e3, e3, K2, F1)
      call VecSetValue(s%rhsVec, 5, 1.5d0, ADD_VALUES, errpetsc)
      call s%factoriseAndSolve()
      call s%free()
      END PROGRAM SolverProcedures
"""


def make_solver_procedures(path):
    """solverpetsc.F's own procedures (initialise, setZero, assembleMatrix / Vector / MatrixAndVector, factorise, solve,
    factoriseAndSolve) executed under the harness above: the Mat / Vec at KSPSolve, and where the state machine STOPs."""
    import warnings
    from oracle.refrun import fortran_to_py as F
    from oracle.refrun import mocks
    from oracle.refrun.runtime import Runtime, _rt
    sources = R.read_sources(['solverpetsc.F'])
    sources['harness.F'] = SOLVER_HARNESS
    code = compile(F.translate(sources), '<solver harness>', 'exec')
    out = {}
    for mode in (0, 1, 2):
        world = mocks.World(1)
        _rt.bind(Runtime(['harness', str(mode)], '.', 0, world, True))
        ns = dict(mocks.namespace())
        exec(code, ns)
        try:
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                ns['program_solverprocedures']()
            assert mode == 0
            rowptr, col, val, rhs = world.system
            out.update(rowptr=rowptr.astype(np.int32), col=col.astype(np.int32), val=val, rhs=rhs)
        except FortranStop as ex:
            assert mode in (1, 2)
            out[f'stop_mode{mode}_line'] = np.int32(ex.line)
            out[f'stop_mode{mode}_msg'] = np.array(ex.msg)
    if path:
        np.savez_compressed(path, **out)
    return out


# ---- the compiled mesh generator ----------------------------------------------------------------------------------------

GENTETRA_GRIDS = {
    # name: (x0, x1, nEx, y0, y1, nEy, z0, z1, nEz)
    'tet10_fixture': (-2, 2, 10, -1, 1, 10, -1, 1, 10),
    'defaults5': (-1.6, 1.6, 5, -1.6, 1.6, 5, -1.6, 1.6, 5),
    'one_cell': (0, 1, 1, 0, 1, 1, 0, 1, 1),
    'cube30': (-1, 1, 30, -1, 1, 30, -1, 1, 30),
    'beam': (-0.5, 0.5, 6, 0.0, 6.0, 36, -0.5, 0.5, 6),
    'ragged': (0.1, 0.7, 7, -3, 5, 2, 2, 2.5, 3),
}


def make_gentetra(path):
    """sha-256 of the files the COMPILED reference generator (oracle/_ref/genTetranovtk) writes, per grid."""
    import hashlib
    import json
    from oracle import ref_gentetra as G
    assert G.build(), "the reference generator did not build"
    out = {}
    for name, grid in GENTETRA_GRIDS.items():
        r = G.run(*grid)
        out[name] = dict(grid=list(grid), nNode=int(r['coords'].shape[1]), nElem=int(r['conn'].shape[1]),
                         sha_nodes=r['sha_nodes'], sha_elems=r['sha_elems'], nDBC=int(r['dbc_node'].size),
                         sha_dbc_nodes=hashlib.sha256(r['dbc_node'].astype('<i4').tobytes()).hexdigest())
    with open(path, 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    return out


if __name__ == '__main__':
    if not R.available():
        sys.exit("the reference tree is not here (PFEM_REFERENCE_SRC / /root/reference/src)")
    make_gentetra(os.path.join(HERE, 'ref_gentetra.json'))
    print('ref_gentetra.json written')
    make_elements(os.path.join(HERE, 'ref_elements.npz'))
    print('ref_elements.npz written')
    for tag in make_drivers(HERE):
        print('ref_driver_%s.npz written' % tag)
    make_solver_procedures(os.path.join(HERE, 'ref_solver_procedures.npz'))
    print('ref_solver_procedures.npz written')
    make_explicit(os.path.join(HERE, 'ref_explicit_cookmembranetria32.npz'))
    print('ref_explicit_cookmembranetria32.npz written')
