"""Self-tests of oracle/refrun, the machinery that executes the reference's Fortran (test infrastructure): each Fortran
semantic the golden vectors lean on is checked on a few lines of Fortran written here, independent of the reference."""
import numpy as np
import pytest

from oracle.refrun import fortran_to_py as F
from oracle.refrun import mocks
from oracle.refrun.runtime import INT_SENTINEL, FortranStop, Ref, Runtime, _rt


def _load(text, extra=None):
    ns = dict(mocks.namespace())
    ns.update(extra or {})
    exec(compile(F.translate({"t.F": text}), "<t.F>", "exec"), ns)
    return ns


def _fixed(body):
    """indent free text into fixed form (statements from column 7)."""
    return "\n".join(("      " + l) if l and not l.startswith(("#", "!", "     ")) else l for l in body.splitlines()) + "\n"


def test_single_precision_literals_and_promotion():
    ns = _load(_fixed("""
SUBROUTINE lit(a, b, c, d)
DOUBLE PRECISION :: a, b, c, d
a = 1.0/3.0
b = 1.0d0/3.0d0
c = 0.1
d = a*3
END SUBROUTINE lit
"""))
    r = [Ref(None) for _ in range(4)]
    ns["lit"](*r)
    assert r[0].v == np.float64(np.float32(1.0) / np.float32(3.0)) and r[0].v != 1.0 / 3.0
    assert r[1].v == 1.0 / 3.0
    assert r[2].v == np.float64(np.float32(0.1))
    assert type(r[3].v) is np.float64 and r[3].v == r[0].v * 3.0


def test_integer_division_truncates_and_mod_follows_the_dividend():
    ns = _load(_fixed("""
SUBROUTINE idiv(a, b, q, r, x)
INTEGER :: a, b, q, r
DOUBLE PRECISION :: x
q = a/b
r = MOD(a, b)
x = a/b
END SUBROUTINE idiv
"""))
    for a, b in ((7, 2), (-7, 2), (7, -2), (-7, -2), (6, 3)):
        q, r, x = Ref(0), Ref(0), Ref(None)
        ns["idiv"](Ref(a), Ref(b), q, r, x)
        assert q.v == int(a / b) and r.v == a - int(a / b) * b
        assert x.v == float(int(a / b))          # integer quotient first, then converted


def test_real_to_integer_assignment_truncates_and_1e9_is_exact():
    ns = _load(_fixed("""
SUBROUTINE conv(i, j, k)
INTEGER :: i, j, k
i = 1e9
j = -2.7
k = 2.7d0
END SUBROUTINE conv
"""))
    r = [Ref(0), Ref(0), Ref(0)]
    ns["conv"](*r)
    assert [x.v for x in r] == [1000000000, -2, 2] and all(type(x.v) is int for x in r)


def test_do_loop_bounds_final_value_exit_and_zero_trip():
    ns = _load(_fixed("""
SUBROUTINE loops(n, s, last, lastexit, lastzero)
INTEGER :: n, s, last, lastexit, lastzero, i
s = 0
DO i=1,n
  s = s + i
END DO
last = i
DO i=1,n
  IF(i == 3) EXIT
END DO
lastexit = i
DO i=5,n-100
  s = s + 1000
END DO
lastzero = i
END SUBROUTINE loops
"""))
    s, last, le, lz = Ref(0), Ref(0), Ref(0), Ref(0)
    ns["loops"](Ref(10), s, last, le, lz)
    assert (s.v, last.v, le.v, lz.v) == (55, 11, 3, 5)


def test_step_loops_select_case_and_named_constructs():
    ns = _load(_fixed("""
SUBROUTINE sel(k, r, cnt)
INTEGER :: k, r, cnt, i
SELECT CASE (k)
  CASE (0)
    r = 10
  CASE (1, 2)
    r = 20
  CASE DEFAULT
    r = -1
END SELECT
cnt = 0
Outer: DO i=10,1,-3
  cnt = cnt + 1
END DO Outer
END SUBROUTINE sel
"""))
    for k, want in ((0, 10), (1, 20), (2, 20), (7, -1)):
        r, c = Ref(0), Ref(0)
        ns["sel"](Ref(k), r, c)
        assert (r.v, c.v) == (want, 4)


def test_scalars_by_reference_array_elements_and_save():
    ns = _load(_fixed("""
SUBROUTINE inc(x)
DOUBLE PRECISION :: x
INTEGER :: calls=0
calls = calls + 1
x = x + calls
END SUBROUTINE inc

SUBROUTINE caller(a, s)
DOUBLE PRECISION :: a(3), s
call inc(a(2))
call inc(a(2))
call inc(s)
END SUBROUTINE caller
"""))
    a, s = np.array([1.0, 2.0, 3.0]), Ref(np.float64(0.0))
    ns["caller"](a, s)
    assert list(a) == [1.0, 5.0, 3.0]            # +1, +2: the initialised local is SAVEd between calls
    assert s.v == 3.0


def test_matmul_is_the_ascending_inner_sum_from_zero():
    ns = _load(_fixed("""
SUBROUTINE mm(a, b, c, v, w)
DOUBLE PRECISION :: a(3,3), b(3,3), c(3,3), v(3), w(3)
c = MATMUL(a, b)
w = MATMUL(TRANSPOSE(a), v)
END SUBROUTINE mm
"""))
    rng = np.random.default_rng(0)
    a = np.asfortranarray(rng.standard_normal((3, 3)) * 10.0 ** rng.integers(-8, 8, (3, 3)))
    b = np.asfortranarray(rng.standard_normal((3, 3)) * 10.0 ** rng.integers(-8, 8, (3, 3)))
    v = rng.standard_normal(3)
    c, w = np.zeros((3, 3), order="F"), np.zeros(3)
    ns["mm"](a, b, c, v, w)
    for i in range(3):
        acc = 0.0
        for l in range(3):
            acc = acc + a[l, i] * v[l]
        assert w[i] == acc
        for j in range(3):
            acc = 0.0
            for l in range(3):
                acc = acc + a[i, l] * b[l, j]
            assert c[i, j] == acc


def test_uninitialised_storage_is_visible_and_nonconforming_assignment_is_noted():
    ns = _load(_fixed("""
SUBROUTINE uninit(x, n)
DOUBLE PRECISION :: x, work(4)
INTEGER :: n, iw(6), src(2,3)
src = 7
iw = src(1,:)
x = work(2)
n = iw(5)
END SUBROUTINE uninit
"""))
    rt = Runtime()
    _rt.bind(rt)
    x, n = Ref(None), Ref(0)
    ns["uninit"](x, n)
    assert np.isnan(x.v) and n.v == INT_SENTINEL
    assert rt.notes and "non-conforming" in rt.notes[0]


def test_stop_power_and_logical_operators():
    ns = _load(_fixed("""
SUBROUTINE misc(x, y, flag)
DOUBLE PRECISION :: x, y
LOGICAL :: flag
y = x**3
IF(flag .NEQV. .TRUE.) THEN
  STOP "not set"
END IF
IF(x < 0.0 .OR. .NOT. flag) STOP "negative"
END SUBROUTINE misc
"""))
    x = np.float64(1.1)
    y = Ref(None)
    ns["misc"](Ref(x), y, Ref(True))
    assert y.v == (x * x) * x
    with pytest.raises(FortranStop, match="not set"):
        ns["misc"](Ref(x), Ref(None), Ref(False))
    with pytest.raises(FortranStop, match="negative"):
        ns["misc"](Ref(np.float64(-1.0)), Ref(None), Ref(True))


def test_fixed_form_continuations_comments_semicolons_and_cpp_lines():
    text = (
        "#include <petsc/finclude/petscsysdef.h>\n"
        "! a comment in column 1\n"
        "      SUBROUTINE cont(a,   ! trailing comment\n"
        "     1    b)\n"
        "      INTEGER :: a, b\n"
        "    !   b = 99\n"
        "      a = 1; b = a +\n"
        "     +    41   ! '+' in column 6 continues the statement\n"
        "      CHKERRQ(a)\n"
        "      END SUBROUTINE cont\n")
    ns = _load(text)
    a, b = Ref(0), Ref(0)
    ns["cont"](a, b)
    assert (a.v, b.v) == (1, 42)


def test_unsupported_statements_are_refused_not_skipped():
    with pytest.raises(F.Unsupported):
        _load(_fixed("""
SUBROUTINE bad(a)
INTEGER :: a
GOTO 10
END SUBROUTINE bad
"""))


def test_mock_petsc_matsetvalues_is_row_major_ignores_negatives_and_keeps_inserted_zeros():
    world = mocks.World(1)
    rt = Runtime(rank=0, world=world)
    _rt.bind(rt)
    m, err = Ref(None), Ref(0)
    mocks.matcreate(Ref("comm"), m, err)
    mocks.matsetsizes(m, Ref(3), Ref(3), Ref(3), Ref(3), err)
    idx = np.array([0, -1, 2])
    k = np.asfortranarray(np.arange(9.0).reshape(3, 3))          # k[i, j] = 3 i + j  (Fortran element (i+1, j+1))
    mocks.matsetvalues(m, Ref(3), idx, Ref(3), idx, np.zeros((3, 3), order="F"), Ref(mocks.INSERT_VALUES), err)
    mocks.matassemblyend(m, Ref(mocks.MAT_FINAL_ASSEMBLY), err)
    mocks.matsetvalues(m, Ref(3), idx, Ref(3), idx, k, Ref(mocks.ADD_VALUES), err)
    rowptr, col, val = m.v.csr()
    assert list(rowptr) == [0, 2, 2, 4] and list(col) == [0, 2, 0, 2]
    # PETSc reads v row-major from the Fortran array's memory: entry (row idx[i], col idx[j]) gets Klocal(j+1, i+1)
    assert list(val) == [k[0, 0], k[2, 0], k[0, 2], k[2, 2]]


def test_mock_mpi_collectives_on_three_ranks():
    import threading
    world = mocks.World(3)
    out = [None] * 3

    def main(r):
        _rt.bind(Runtime(rank=r, world=world))
        err = Ref(0)
        gathered = np.zeros(3, np.int64)
        mocks.mpi_allgather(Ref(10 + r), Ref(1), Ref("MPI_INT"), gathered, Ref(1), Ref("MPI_INT"), Ref("c"), err)
        total = Ref(0)
        mocks.mpi_allreduce(Ref(r + 1), total, Ref(1), Ref("MPI_INT"), Ref("MPI_SUM"), Ref("c"), err)
        buf = np.full(2, r, np.int64)
        mocks.mpi_bcast(buf, Ref(2), Ref("MPI_INT"), Ref(0), Ref("c"), err)
        out[r] = (list(gathered), total.v, list(buf))

    ts = [threading.Thread(target=main, args=(r,)) for r in range(3)]
    [t.start() for t in ts]
    [t.join(30) for t in ts]
    assert out == [([10, 11, 12], 6, [0, 0])] * 3


def test_free_form_interfaces_optional_arguments_and_c_binding(tmp_path):
    """the free-form / ISO_C_BINDING side of the front end (what include/pfem_b200.f90 needs): `&` continuations, INTERFACE
    blocks with BIND(C), VALUE vs by-address dummies, OPTIONAL + PRESENT, a by-address scalar result written back into a
    local and into a derived-type component."""
    import ctypes
    import subprocess
    from oracle.refrun.runtime import set_clib
    csrc = tmp_path / "t.c"
    csrc.write_text("""
        int twice_plus(int a, double *x, int *out, const int *v, int n) {
            int s = 0; for (int i = 0; i < n; i++) s += v[i];
            *out = 2 * a + s; *x = *x + 0.5; return 7; }
        int make_handle(void **h) { static int cell = 41; *h = &cell; return 0; }
        int read_handle(void *h) { return *(int *)h + 1; }
    """)
    so = str(tmp_path / "libt.so")
    subprocess.run(["gcc", "-shared", "-fPIC", "-o", so, str(csrc)], check=True)
    text = """
    ! free form: comments anywhere, & continuations
    MODULE m
      USE, INTRINSIC :: ISO_C_BINDING
      IMPLICIT NONE
      INTERFACE
        INTEGER(C_INT) FUNCTION twice_plus(a, x, out, v, n) BIND(C)
          IMPORT; INTEGER(C_INT), VALUE :: a, n
          REAL(C_DOUBLE) :: x
          INTEGER(C_INT) :: out, v(*)
        END FUNCTION
        INTEGER(C_INT) FUNCTION make_handle(h) BIND(C)
          IMPORT; TYPE(C_PTR) :: h
        END FUNCTION
        INTEGER(C_INT) FUNCTION read_handle(h) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
        END FUNCTION
      END INTERFACE
      TYPE Box
        TYPE(C_PTR) :: h = C_NULL_PTR
        INTEGER :: n = 3
      END TYPE Box
    CONTAINS
      SUBROUTINE drive(b, a, x, res, rc, opt)
        CLASS(Box) :: b
        INTEGER :: a, res, rc, &
                   got            ! continued declaration
        DOUBLE PRECISION :: x
        INTEGER, OPTIONAL :: opt
        INTEGER :: v(3)
        v(1) = 1; v(2) = 2; v(3) = 3
        rc = twice_plus(a, x, got, v, b%n)
        res = got
        IF (PRESENT(opt)) res = res + opt
        rc = rc + make_handle(b%h)
        res = res + 1000*read_handle(b%h)
      END SUBROUTINE drive
    END MODULE m
    """
    ns = {}
    exec(compile(F.translate({"m.f90": text}), "<m.f90>", "exec"), ns)
    set_clib(ctypes.CDLL(so))
    b = ns["_new_box"]()
    assert b.h is None and b.n == 3
    x, res, rc = Ref(np.float64(1.25)), Ref(0), Ref(0)
    ns["drive"](b, Ref(5), x, res, rc)
    assert (x.v, rc.v) == (1.75, 7) and res.v == (2 * 5 + 6) + 1000 * 42 and b.h
    res2 = Ref(0)
    ns["drive"](b, Ref(5), Ref(np.float64(0.0)), res2, Ref(0), Ref(100))
    assert res2.v == res.v + 100
