/*
 * pfem_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the implicit hot path of chennachaos/PFEMFort:
 *   element Ke/Fe  ->  global sparse assembly (+ Dirichlet lifting)  ->  Jacobi-CG.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.  The product library (libpfemb200.so)
 * never links, loads or calls it.
 *
 * PARITY: PINNED BY RUNNING THE REFERENCE'S OWN SOURCE (round 2).  The reference ships no
 * tests, golden vectors or timings and its Fortran + PETSc + MPI
 * programs cannot be compiled here (no gfortran / MPI / PETSc).  Instead they are EXECUTED
 * from /root/reference/src by oracle/refrun (a Fortran-subset -> Python translator with
 * Fortran's typing and rounding rules; PETSc / MPI / METIS / the VTK writer, which are
 * external to the reference, come from oracle/refrun/mocks.py), and what those runs
 * produce is committed as tests/golden/ref_*.npz by tests/golden/make_reference_vectors.py:
 *   - the eight element routines of the path on seeded random elements, incl. where
 *     the reference STOPs                                      -> orc_*_ke, orc_residual_*, orc_mass_*
 *   - the four *parallelimpl1 PROGRAMs end to end on the bundled inputs, 1..4 simulated
 *     ranks: numbering, ElemDofArray, the Mat / Vec handed to KSPSolve, solver options
 *                                                              -> orc_number_dofs, orc_pattern, orc_assemble, ...
 *   - triaelasticityexplicit.F end to end (40 steps)           -> orc_explicit_*
 * tests/test_reference_vectors.py holds every function here to those files BIT FOR BIT
 * (P > 1 ranks: integers bit for bit, values to 1e-12 -- the reference's own sums then
 * depend on PETSc's stash order).  The mesh generator, the one program of the reference
 * that compiles with g++ alone, is built as oracle/_ref/genTetranovtk (make ref).
 * NOT pinned by a reference run: the Krylov iterates (KSPSolve is PETSc's, a third-party
 * dependency absent from /root/reference, pinned only by the path string at
 * CMakeLists.txt:43).  Its MatSetValues / KSPSolve_CG / PCApply_Jacobi / PCBJACOBI+ILU(0)
 * semantics are restated from its published behaviour and anchored on the reference's
 * call sites, on the analytic answers baked into the reference's BC fixtures (tria20x20:
 * Laplace solution; tet10: u = x^2+y^2+z^2 with source -6), on the closed-form Ke of
 * triapoissonserialimpl1.F:580-594 and on the identities in tests/test_oracle.py.
 *
 * Arithmetic conventions reproduced:
 *   - gfortran without -fdefault-real-8: every real literal is SINGLE precision,
 *     so 1.0/3.0 and 1.0/6.0 are float divisions widened to double;
 *   - no FMA contraction (build with -ffp-contract=off), left-to-right evaluation;
 *   - MATMUL sums the inner index in ascending order, zeros included;
 *   - PETSc reads the column-major Klocal as row-major: entry (row i, col j)
 *     receives Klocal(j,i).
 *
 * Array conventions: all 2-D arrays are Fortran column-major ("SoA on the wire"):
 *   conn[i*nElem + e]   = elemNodeConn(e+1, i+1)  (1-based node ids)
 *   coords[c*nNode + n] = coords(n+1, c+1)
 *   edof[k*nElem + e]   = ElemDofArray(e+1, k+1)  (0-based dof ids, -1 = Dirichlet)
 *   K[i + n*j]          = Klocal(i+1, j+1)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Basis functions                                                            */
/* ------------------------------------------------------------------------- */

/* elementutilitiesbasisfuncs.F:39-51 (LagrangeBasisFunsTria, degree 1) */
static void lagrange_tria_p1(double xi1, double xi2, double N[3], double d1[3], double d2[3])
{
    double xi3 = 1.0 - xi1 - xi2;
    N[0] = xi3; N[1] = xi1; N[2] = xi2;
    d1[0] = -1.0; d1[1] = 1.0; d1[2] = 0.0;
    d2[0] = -1.0; d2[1] = 0.0; d2[2] = 1.0;
}

/* elementutilitiesbasisfuncs.F:165-234 (computeBasisFunctions2D, ETYPE=1, degree 1) */
static void basis2d_tria(const double param[2], const double x[3], const double y[3],
                         double N[3], double dNdx[3], double dNdy[3], double *Jac)
{
    double d1[3], d2[3], B11 = 0.0, B21 = 0.0, B12 = 0.0, B22 = 0.0;
    lagrange_tria_p1(param[0], param[1], N, d1, d2);
    for (int ii = 0; ii < 3; ii++) {                 /* :208-217 */
        double xx = x[ii], yy = y[ii];
        B11 = B11 + (xx * d1[ii]);
        B21 = B21 + (xx * d2[ii]);
        B12 = B12 + (yy * d1[ii]);
        B22 = B22 + (yy * d2[ii]);
    }
    double J = B11 * B22 - B12 * B21;                 /* :219 */
    double detinv = 1.0 / J;                          /* :221 */
    double Bi11 = B22 * detinv, Bi12 = -B12 * detinv; /* :223-226 */
    double Bi21 = -B21 * detinv, Bi22 = B11 * detinv;
    for (int ii = 0; ii < 3; ii++) {                 /* :229-232 */
        dNdx[ii] = d1[ii] * Bi11 + d2[ii] * Bi12;
        dNdy[ii] = d1[ii] * Bi21 + d2[ii] * Bi22;
    }
    *Jac = J;
}

/* elementutilitiesbasisfuncs.F:261-281 (LagrangeBasisFunsTet, degree 1): node 3 is the origin */
static void lagrange_tet_p1(double xi1, double xi2, double xi3, double N[4],
                            double d1[4], double d2[4], double d3[4])
{
    N[0] = xi1; N[1] = xi2; N[2] = 1.0 - xi1 - xi2 - xi3; N[3] = xi3;
    d1[0] = 1.0; d1[1] = 0.0; d1[3] = 0.0; d1[2] = -1.0;
    d2[0] = 0.0; d2[1] = 1.0; d2[3] = 0.0; d2[2] = -1.0;
    d3[0] = 0.0; d3[1] = 0.0; d3[3] = 1.0; d3[2] = -1.0;
}

/* elementutilitiesbasisfuncs.F:430-538 (computeBasisFunctions3D, ETYPE=4, degree 1) */
static void basis3d_tet(const double param[3], const double x[4], const double y[4], const double z[4],
                        double N[4], double dNdx[4], double dNdy[4], double dNdz[4], double *Jac)
{
    double d1[4], d2[4], d3[4], B[3][3], Bi[3][3];   /* B[r][c] = B(r+1,c+1) */
    lagrange_tet_p1(param[0], param[1], param[2], N, d1, d2, d3);
    memset(B, 0, sizeof B);
    for (int ii = 0; ii < 4; ii++) {                 /* :493-509 */
        double xx = x[ii], yy = y[ii], zz = z[ii];
        B[0][0] = B[0][0] + (xx * d1[ii]);
        B[1][0] = B[1][0] + (xx * d2[ii]);
        B[2][0] = B[2][0] + (xx * d3[ii]);
        B[0][1] = B[0][1] + (yy * d1[ii]);
        B[1][1] = B[1][1] + (yy * d2[ii]);
        B[2][1] = B[2][1] + (yy * d3[ii]);
        B[0][2] = B[0][2] + (zz * d1[ii]);
        B[1][2] = B[1][2] + (zz * d2[ii]);
        B[2][2] = B[2][2] + (zz * d3[ii]);
    }
    double J;                                         /* :512-514 */
    J = B[0][0] * (B[1][1] * B[2][2] - B[1][2] * B[2][1]);
    J = J + B[0][1] * (B[1][2] * B[2][0] - B[1][0] * B[2][2]);
    J = J + B[0][2] * (B[1][0] * B[2][1] - B[1][1] * B[2][0]);
    double detinv = 1.0 / J;                          /* :517 */
    Bi[0][0] = +detinv * (B[1][1] * B[2][2] - B[1][2] * B[2][1]);   /* :520-528 */
    Bi[1][0] = -detinv * (B[1][0] * B[2][2] - B[1][2] * B[2][0]);
    Bi[2][0] = +detinv * (B[1][0] * B[2][1] - B[1][1] * B[2][0]);
    Bi[0][1] = -detinv * (B[0][1] * B[2][2] - B[0][2] * B[2][1]);
    Bi[1][1] = +detinv * (B[0][0] * B[2][2] - B[0][2] * B[2][0]);
    Bi[2][1] = -detinv * (B[0][0] * B[2][1] - B[0][1] * B[2][0]);
    Bi[0][2] = +detinv * (B[0][1] * B[1][2] - B[0][2] * B[1][1]);
    Bi[1][2] = -detinv * (B[0][0] * B[1][2] - B[0][2] * B[1][0]);
    Bi[2][2] = +detinv * (B[0][0] * B[1][1] - B[0][1] * B[1][0]);
    for (int ii = 0; ii < 4; ii++) {                 /* :532-536 */
        dNdx[ii] = d1[ii] * Bi[0][0] + d2[ii] * Bi[0][1] + d3[ii] * Bi[0][2];
        dNdy[ii] = d1[ii] * Bi[1][0] + d2[ii] * Bi[1][1] + d3[ii] * Bi[1][2];
        dNdz[ii] = d1[ii] * Bi[2][0] + d2[ii] * Bi[2][1] + d3[ii] * Bi[2][2];
    }
    *Jac = J;
}

/* ------------------------------------------------------------------------- */
/* Element routines.  Return 0 = ok, 1 = negative Jacobian (the reference STOPs) */
/* ------------------------------------------------------------------------- */

/* elementutilitiespoisson.F:23-101 */
ORC_API int orc_poisson_tria_ke(const double x[3], const double y[3], const double *elemData,
                                const double *timeData, const double valC[3], const double valDotC[3],
                                double K[9], double F[3])
{
    (void)valDotC;
    double kx = elemData[0], ky = elemData[1];
    double af = timeData[1];
    double param[2] = {(double)(1.0f / 3.0f), (double)(1.0f / 3.0f)};   /* :56 single-precision literal */
    double gw = 0.5;
    double N[3], dNdx[3], dNdy[3], Jac;
    for (int i = 0; i < 9; i++) K[i] = 0.0;
    for (int i = 0; i < 3; i++) F[i] = 0.0;
    basis2d_tria(param, x, y, N, dNdx, dNdy, &Jac);
    if (Jac < 0.0) return 1;                          /* :71 */
    double dvol = gw * Jac;                           /* :75 */
    double du[2] = {0.0, 0.0};
    for (int ii = 0; ii < 3; ii++) {                 /* :77-81 */
        du[0] = du[0] + valC[ii] * dNdx[ii];
        du[1] = du[1] + valC[ii] * dNdy[ii];
    }
    double force = 0.0;                               /* :83 */
    for (int ii = 0; ii < 3; ii++) {                 /* :85-96 */
        double b1 = dNdx[ii] * dvol, b2 = dNdy[ii] * dvol, b4 = N[ii] * dvol;
        F[ii] = F[ii] + b4 * force - b1 * du[0] - b2 * du[1];
        for (int jj = 0; jj < 3; jj++)
            K[ii + 3 * jj] = K[ii + 3 * jj] + af * (b1 * (kx * dNdx[jj]) + b2 * (ky * dNdy[jj]));
    }
    return 0;
}

/* elementutilitiespoisson.F:107-193 */
ORC_API int orc_poisson_tetra_ke(const double x[4], const double y[4], const double z[4],
                                 const double *elemData, const double *timeData, const double valC[4],
                                 const double valDotC[4], double K[16], double F[4])
{
    (void)valDotC;
    double kx = elemData[0], ky = elemData[1], kz = elemData[2];
    double af = timeData[1];
    double param[3] = {0.25, 0.25, 0.25};
    double gw = (double)(1.0f / 6.0f);                /* :142 single-precision literal 0.1666666716337204 */
    double N[4], dNdx[4], dNdy[4], dNdz[4], Jac;
    for (int i = 0; i < 16; i++) K[i] = 0.0;
    for (int i = 0; i < 4; i++) F[i] = 0.0;
    basis3d_tet(param, x, y, z, N, dNdx, dNdy, dNdz, &Jac);
    if (Jac < 0.0) return 1;                          /* :157 */
    double dvol = gw * Jac;                           /* :161 */
    double du[3] = {0.0, 0.0, 0.0};
    for (int ii = 0; ii < 4; ii++) {                 /* :165-170 */
        du[0] = du[0] + valC[ii] * dNdx[ii];
        du[1] = du[1] + valC[ii] * dNdy[ii];
        du[2] = du[2] + valC[ii] * dNdz[ii];
    }
    double force = -6.0;                              /* :172 */
    for (int ii = 0; ii < 4; ii++) {                 /* :174-189 */
        double b1 = dNdx[ii] * dvol, b2 = dNdy[ii] * dvol, b3 = dNdz[ii] * dvol, b4 = N[ii] * dvol;
        F[ii] = F[ii] + b4 * force;
        F[ii] = F[ii] - b1 * du[0] - b2 * du[1] - b3 * du[2];
        for (int jj = 0; jj < 4; jj++)
            K[ii + 4 * jj] = K[ii + 4 * jj] +
                             af * (b1 * (kx * dNdx[jj]) + b2 * (ky * dNdy[jj]) + b3 * (kz * dNdz[jj]));
    }
    return 0;
}

/* Fortran MATMUL restated: C(m x n) = A(m x p) * B(p x n), column-major, inner index ascending. */
static void matmul_cm(int m, int p, int n, const double *A, const double *B, double *C)
{
    for (int j = 0; j < n; j++)
        for (int i = 0; i < m; i++) {
            double s = 0.0;
            for (int k = 0; k < p; k++) s = s + A[i + m * k] * B[k + p * j];
            C[i + m * j] = s;
        }
}

/* elementutilitieselasticity2D.F:23-153 */
ORC_API int orc_elasticity_tria_ke(const double x[3], const double y[3], const double *elemData,
                                   const double *timeData, const double *valC, const double *valDotC,
                                   double K[36], double F[6])
{
    (void)valC; (void)valDotC; (void)timeData;        /* grad/strain/stress (:98-119) never reach K or F */
    double E = elemData[0], nu = elemData[1], thick = elemData[2];
    double bforce[2] = {elemData[3], elemData[4]};
    double b1 = E / (1.0 - nu * nu);                  /* :59 */
    double D[9];                                      /* column-major 3x3, :62-64 (D33 = b1*(1-nu), sic) */
    D[0] = b1;      D[3] = b1 * nu; D[6] = 0.0;
    D[1] = b1 * nu; D[4] = b1;      D[7] = 0.0;
    D[2] = 0.0;     D[5] = 0.0;     D[8] = b1 * (1.0 - nu);
    double param[2] = {(double)(1.0f / 3.0f), (double)(1.0f / 3.0f)};   /* :74 */
    double gw = 0.5;
    double N[3], dNdx[3], dNdy[3], Jac;
    for (int i = 0; i < 36; i++) K[i] = 0.0;
    for (int i = 0; i < 6; i++) F[i] = 0.0;
    basis2d_tria(param, x, y, N, dNdx, dNdy, &Jac);
    if (Jac < 0.0) return 1;                          /* :90 */
    double dvol = gw * (Jac * thick);                 /* :94 */
    double Bm[18], BT[18], DB[18], KK[36];           /* Bmat 3x6, BmatTrans 6x3 */
    for (int i = 0; i < 18; i++) Bm[i] = 0.0;
    for (int ii = 0; ii < 3; ii++) {                 /* :127-133 */
        int TI = 2 * ii, TIp1 = TI + 1;
        Bm[0 + 3 * TI] = dNdx[ii]; Bm[0 + 3 * TIp1] = 0.0;
        Bm[1 + 3 * TI] = 0.0;      Bm[1 + 3 * TIp1] = dNdy[ii];
        Bm[2 + 3 * TI] = dNdy[ii]; Bm[2 + 3 * TIp1] = dNdx[ii];
    }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 6; c++) BT[c + 6 * r] = Bm[r + 3 * c];   /* :135 */
    matmul_cm(3, 3, 6, D, Bm, DB);                    /* :136 */
    matmul_cm(6, 3, 6, BT, DB, KK);                   /* :138 */
    for (int i = 0; i < 36; i++) K[i] = dvol * KK[i]; /* :139 */
    for (int ii = 0; ii < 3; ii++) {                 /* :142-150 */
        int TI = 2 * ii, TIp1 = TI + 1;
        double b4 = dvol * N[ii];
        F[TI] = F[TI] + b4 * bforce[0];
        F[TIp1] = F[TIp1] + b4 * bforce[1];
    }
    return 0;
}

/* elementutilitieselasticity3D.F:248-393, documented intent: ETYPE=4, nGP=1 */
ORC_API int orc_elasticity_tetra_ke(const double x[4], const double y[4], const double z[4],
                                    const double *elemData, const double *timeData, const double *valC,
                                    const double *valDotC, double K[144], double F[12])
{
    (void)valC; (void)valDotC; (void)timeData;
    double E = elemData[0], nu = elemData[1];
    double bforce[3] = {elemData[3], elemData[4], elemData[5]};
    double b1 = E / ((1.0 + nu) * (1.0 - 2.0 * nu));  /* :284 */
    double b2 = (1.0 - 2.0 * nu) / 2.0;               /* :285 */
    double D[36];
    for (int i = 0; i < 36; i++) D[i] = 0.0;          /* :287-296 */
    D[0 + 6 * 0] = b1 * (1.0 - nu); D[0 + 6 * 1] = b1 * nu;         D[0 + 6 * 2] = b1 * nu;
    D[1 + 6 * 0] = b1 * nu;         D[1 + 6 * 1] = b1 * (1.0 - nu); D[1 + 6 * 2] = b1 * nu;
    D[2 + 6 * 0] = b1 * nu;         D[2 + 6 * 1] = b1 * nu;         D[2 + 6 * 2] = b1 * (1.0 - nu);
    D[3 + 6 * 3] = b1 * b2; D[4 + 6 * 4] = b1 * b2; D[5 + 6 * 5] = b1 * b2;
    double param[3] = {0.25, 0.25, 0.25};
    double gw = (double)(1.0f / 6.0f);                /* :305 */
    double N[4], dNdx[4], dNdy[4], dNdz[4], Jac;
    for (int i = 0; i < 144; i++) K[i] = 0.0;
    for (int i = 0; i < 12; i++) F[i] = 0.0;
    basis3d_tet(param, x, y, z, N, dNdx, dNdy, dNdz, &Jac);
    if (Jac < 0.0) return 1;                          /* :320 */
    double dvol = gw * Jac;                           /* :324 */
    double Bm[72], BT[72], DB[72], KK[144];          /* Bmat 6x12 */
    for (int i = 0; i < 72; i++) Bm[i] = 0.0;
    for (int ii = 0; ii < 4; ii++) {                 /* :360-371 */
        int TI = 3 * ii, T1 = TI + 1, T2 = TI + 2;
        Bm[0 + 6 * TI] = dNdx[ii]; Bm[0 + 6 * T1] = 0.0;      Bm[0 + 6 * T2] = 0.0;
        Bm[1 + 6 * TI] = 0.0;      Bm[1 + 6 * T1] = dNdy[ii]; Bm[1 + 6 * T2] = 0.0;
        Bm[2 + 6 * TI] = 0.0;      Bm[2 + 6 * T1] = 0.0;      Bm[2 + 6 * T2] = dNdz[ii];
        Bm[3 + 6 * TI] = dNdy[ii]; Bm[3 + 6 * T1] = dNdx[ii]; Bm[3 + 6 * T2] = 0.0;
        Bm[4 + 6 * TI] = 0.0;      Bm[4 + 6 * T1] = dNdz[ii]; Bm[4 + 6 * T2] = dNdy[ii];
        Bm[5 + 6 * TI] = dNdz[ii]; Bm[5 + 6 * T1] = 0.0;      Bm[5 + 6 * T2] = dNdx[ii];
    }
    for (int r = 0; r < 6; r++)
        for (int c = 0; c < 12; c++) BT[c + 12 * r] = Bm[r + 6 * c];  /* :373 */
    matmul_cm(6, 6, 12, D, Bm, DB);                   /* :374 */
    matmul_cm(12, 6, 12, BT, DB, KK);                 /* :376 */
    for (int i = 0; i < 144; i++) K[i] = dvol * KK[i];/* :377 */
    for (int ii = 0; ii < 4; ii++) {                 /* :380-390 */
        int TI = 3 * ii;
        double b4 = dvol * N[ii];
        F[TI] = F[TI] + b4 * bforce[0];
        F[TI + 1] = F[TI + 1] + b4 * bforce[1];
        F[TI + 2] = F[TI + 2] + b4 * bforce[2];
    }
    return 0;
}

/* Independent closed form for the P1 triangle, triapoissonserialimpl1.F:580-594: Ke = area * B B^T */
ORC_API void orc_poisson_tria_ke_closed_form(const double x[3], const double y[3], double K[9])
{
    double x1 = x[0], x2 = x[1], x3 = x[2], y1 = y[0], y2 = y[1], y3 = y[2];
    double area = 0.5 * (x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2));
    double Bx[3] = {(y2 - y3) / (2.0 * area), (y3 - y1) / (2.0 * area), (y1 - y2) / (2.0 * area)};
    double By[3] = {(x3 - x2) / (2.0 * area), (x1 - x3) / (2.0 * area), (x2 - x1) / (2.0 * area)};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) K[i + 3 * j] = area * (Bx[i] * Bx[j] + By[i] * By[j]);
}

/* ------------------------------------------------------------------------- */
/* Driver numbering (tetrapoissonparallelimpl1.F:357-367, 402-421, 500-677)   */
/* ------------------------------------------------------------------------- */

/*
 * In : nNode, ndof, DBC rows (1-based node, 1-based dof, value), nparts, node_proc_id (NULL if nparts==1).
 * Out: node_map_get_old/new (1-based, length nNode), NodeDofArrayNew (column-major nNode x ndof,
 *      1-based dof id, 0 = Dirichlet), solnApplied (length nNode*ndof, NEW numbering),
 *      node_start/node_end/row_start/row_end (1-based, per part), size_local (per part).
 * Returns size_global, or -1 on the reference's "Something wrong" STOPs.
 */
ORC_API int orc_number_dofs(int nNode, int ndof, int nDBC, const int *dbc_node, const int *dbc_dof,
                            const double *dbc_val, int nparts, const int *node_proc_id,
                            int *node_map_get_old, int *node_map_get_new, int *NodeDofArrayNew,
                            double *solnApplied, int *node_start, int *node_end, int *row_start,
                            int *row_end, int *size_local)
{
    int *NodeTypeOld = calloc((size_t)nNode * ndof, sizeof(int));
    int *NodeTypeNew = calloc((size_t)nNode * ndof, sizeof(int));
    int *dbc_node_new = malloc(sizeof(int) * (size_t)(nDBC > 0 ? nDBC : 1));
    for (size_t k = 0; k < (size_t)nNode * ndof; k++) { NodeDofArrayNew[k] = 0; solnApplied[k] = 0.0; }
    for (int ii = 0; ii < nDBC; ii++) {               /* :339-352 */
        int n1 = dbc_node[ii], n2 = dbc_dof[ii];
        NodeTypeOld[(n1 - 1) + (size_t)nNode * (n2 - 1)] = 1;
        solnApplied[(size_t)(n1 - 1) * ndof + (n2 - 1)] = dbc_val[ii];
    }
    int size_global = 0;                              /* :357-367 */
    for (int ii = 0; ii < nNode; ii++)
        for (int jj = 0; jj < ndof; jj++)
            if (NodeTypeOld[ii + (size_t)nNode * jj] == 0) size_global++;

    if (nparts == 1) {                                /* :402-421 */
        int ind = 1;
        for (int ii = 0; ii < nNode; ii++) {
            node_map_get_old[ii] = ii + 1;
            node_map_get_new[ii] = ii + 1;
            for (int jj = 0; jj < ndof; jj++)
                if (NodeTypeOld[ii + (size_t)nNode * jj] == 0) NodeDofArrayNew[ii + (size_t)nNode * jj] = ind++;
        }
        node_start[0] = 1; node_end[0] = nNode;
        row_start[0] = 1; row_end[0] = size_global; size_local[0] = size_global;
    } else {
        /* :500-564: ranks concatenate their ascending lists of owned OLD node ids */
        int kk = 0;
        for (int p = 0; p < nparts; p++) {
            node_start[p] = kk + 1;
            for (int ii = 0; ii < nNode; ii++)
                if (node_proc_id[ii] == p) node_map_get_old[kk++] = ii + 1;
            node_end[p] = kk;
        }
        if (kk != nNode) { free(NodeTypeOld); free(NodeTypeNew); free(dbc_node_new); return -1; }
        for (int ii = 0; ii < nNode; ii++) {         /* :588-595 */
            int n1 = node_map_get_old[ii];
            node_map_get_new[n1 - 1] = ii + 1;
            for (int jj = 0; jj < ndof; jj++)
                NodeTypeNew[ii + (size_t)nNode * jj] = NodeTypeOld[(n1 - 1) + (size_t)nNode * jj];
        }
        int ind = 1;                                  /* :601-616 */
        for (int ii = 0; ii < nNode; ii++)
            for (int jj = 0; jj < ndof; jj++)
                if (NodeTypeNew[ii + (size_t)nNode * jj] == 0) NodeDofArrayNew[ii + (size_t)nNode * jj] = ind++;
        if (ind - 1 != size_global) { free(NodeTypeOld); free(NodeTypeNew); free(dbc_node_new); return -1; }
        int total = 0;
        for (int p = 0; p < nparts; p++) {           /* :622-636 */
            int rs = 1000000000, re = -1000000000, sl = 0;
            for (int ii = node_start[p]; ii <= node_end[p]; ii++)
                for (int jj = 0; jj < ndof; jj++)
                    if (NodeTypeNew[(ii - 1) + (size_t)nNode * jj] == 0) {
                        int d = NodeDofArrayNew[(ii - 1) + (size_t)nNode * jj];
                        if (d < rs) rs = d;
                        if (d > re) re = d;
                        sl++;
                    }
            row_start[p] = rs; row_end[p] = re; size_local[p] = sl; total += sl;
        }
        if (total != size_global) { free(NodeTypeOld); free(NodeTypeNew); free(dbc_node_new); return -1; } /* :650-655 */
        /* :668-677: re-key Dirichlet data to NEW numbering (stale OLD-position entries are kept, as in the reference) */
        for (int ii = 0; ii < nDBC; ii++) {
            int n1 = node_map_get_new[dbc_node[ii] - 1];
            solnApplied[(size_t)(n1 - 1) * ndof + (dbc_dof[ii] - 1)] = dbc_val[ii];
        }
    }
    free(NodeTypeOld); free(NodeTypeNew); free(dbc_node_new);
    return size_global;
}

/* tetrapoissonparallelimpl1.F:698-713: ElemDofArray(e, ndof*(i-1)+j) = NodeDofArrayNew(conn(e,i), j) - 1 */
ORC_API void orc_elem_dof_array(int nElem, int npElem, int ndof, int nNode, const int *conn_new,
                                const int *NodeDofArrayNew, int *edof)
{
    for (int ee = 0; ee < nElem; ee++)
        for (int ii = 0; ii < npElem; ii++) {
            int n2 = conn_new[(size_t)ii * nElem + ee];
            for (int jj = 0; jj < ndof; jj++)
                edof[(size_t)(ndof * ii + jj) * nElem + ee] = NodeDofArrayNew[(n2 - 1) + (size_t)nNode * jj] - 1;
        }
}

/* ------------------------------------------------------------------------- */
/* Pattern pass: MatSetValues(INSERT zeros) + setZero                          */
/* (tetrapoissonparallelimpl1.F:791-802, solverpetsc.F:222-246)                */
/* Result per row: sorted unique columns; negative rows/cols dropped; zeros kept. */
/* ------------------------------------------------------------------------- */

static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* Two-call protocol: col == NULL -> fills rowptr[N+1] and returns nnz; else also fills col[nnz]. */
ORC_API long long orc_pattern(int nElem, int nsize, const int *edof, int N, int *rowptr, int *col)
{
    long long *cnt = calloc((size_t)N + 1, sizeof(long long));
    for (int ee = 0; ee < nElem; ee++) {
        int nfree = 0;
        for (int k = 0; k < nsize; k++) nfree += edof[(size_t)k * nElem + ee] >= 0;
        for (int k = 0; k < nsize; k++) {
            int r = edof[(size_t)k * nElem + ee];
            if (r >= 0) cnt[r + 1] += nfree;
        }
    }
    for (int r = 0; r < N; r++) cnt[r + 1] += cnt[r];
    long long total = cnt[N];
    int *buf = malloc(sizeof(int) * (size_t)(total > 0 ? total : 1));
    long long *pos = malloc(sizeof(long long) * ((size_t)N + 1));
    memcpy(pos, cnt, sizeof(long long) * ((size_t)N + 1));
    for (int ee = 0; ee < nElem; ee++)
        for (int k = 0; k < nsize; k++) {
            int r = edof[(size_t)k * nElem + ee];
            if (r < 0) continue;
            for (int l = 0; l < nsize; l++) {
                int c = edof[(size_t)l * nElem + ee];
                if (c >= 0) buf[pos[r]++] = c;
            }
        }
    long long nnz = 0;
    rowptr[0] = 0;
#pragma omp parallel for schedule(dynamic, 4096)
    for (int r = 0; r < N; r++) {
        long long a = cnt[r], b = cnt[r + 1];
        qsort(buf + a, (size_t)(b - a), sizeof(int), cmp_int);
        long long w = a;
        for (long long k = a; k < b; k++)
            if (k == a || buf[k] != buf[k - 1]) buf[w++] = buf[k];
        pos[r] = w - a;                               /* unique count */
    }
    for (int r = 0; r < N; r++) { nnz += pos[r]; rowptr[r + 1] = (int)nnz; }
    if (col)
        for (int r = 0; r < N; r++) memcpy(col + rowptr[r], buf + cnt[r], sizeof(int) * (size_t)pos[r]);
    free(buf); free(pos); free(cnt);
    return nnz;
}

/* ------------------------------------------------------------------------- */
/* Value pass (tetrapoissonparallelimpl1.F:828-884 and the elasticity siblings) */
/* ------------------------------------------------------------------------- */

enum { ORC_POISSON_TRIA = 0, ORC_POISSON_TETRA = 1, ORC_ELASTICITY_TRIA = 2, ORC_ELASTICITY_TETRA = 3 };

static void kind_dims(int kind, int *npElem, int *ndof, int *ndim)
{
    switch (kind) {
    case ORC_POISSON_TRIA: *npElem = 3; *ndof = 1; *ndim = 2; break;
    case ORC_POISSON_TETRA: *npElem = 4; *ndof = 1; *ndim = 3; break;
    case ORC_ELASTICITY_TRIA: *npElem = 3; *ndof = 2; *ndim = 2; break;
    default: *npElem = 4; *ndof = 3; *ndim = 3; break;
    }
}

static int element_ke(int kind, const double *xn, const double *yn, const double *zn, const double *elemData,
                      const double *timeData, double *K, double *F)
{
    static const double zeros[12] = {0};
    switch (kind) {
    case ORC_POISSON_TRIA: return orc_poisson_tria_ke(xn, yn, elemData, timeData, zeros, zeros, K, F);
    case ORC_POISSON_TETRA: return orc_poisson_tetra_ke(xn, yn, zn, elemData, timeData, zeros, zeros, K, F);
    case ORC_ELASTICITY_TRIA: return orc_elasticity_tria_ke(xn, yn, elemData, timeData, zeros, zeros, K, F);
    default: return orc_elasticity_tetra_ke(xn, yn, zn, elemData, timeData, zeros, zeros, K, F);
    }
}

/* PETSc MatSetValues_SeqAIJ row search: the row's columns are sorted; missing => -1 */
static inline long long find_slot(const int *rowptr, const int *col, int r, int c)
{
    int lo = rowptr[r], hi = rowptr[r + 1] - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        if (col[mid] < c) lo = mid + 1;
        else if (col[mid] > c) hi = mid - 1;
        else return mid;
    }
    return -1;
}

/*
 * Sequential-order ADD assembly (np=1 semantics of the reference): elements in ascending id;
 * MatSetValues(ADD) with the column-major Klocal read row-major (entry (row i, col j) += Klocal(j,i));
 * lifting F_j -= Klocal(j,i)*g_i in ascending Dirichlet local index i; VecSetValues(ADD).
 * elem_mask (may be NULL): only elements with elem_mask[e] != 0 are processed (the
 * "elem_proc_id(ee) == this_mpi_proc" test).  row_lo/row_hi: only rows in [row_lo,row_hi) are kept
 * (0-based, emulates one rank's owned block when the caller passes overlap elements); val/rhs are
 * indexed globally.  threads > 1 switches to an OpenMP element loop with atomic adds (timing baseline
 * only: summation order is then unspecified, like PETSc's stash arrival order).
 * Returns the number of negative-Jacobian elements (the reference would STOP at the first).
 */
ORC_API int orc_assemble(int kind, int nElem, const int *conn, int nNode, const double *coords,
                         const int *node_map_get_old, const int *edof, const double *solnApplied,
                         const double *elemData, const double *timeData, const unsigned char *elem_mask,
                         int row_lo, int row_hi, const int *rowptr, const int *col, double *val, double *rhs,
                         int threads)
{
    int npElem, ndof, ndim;
    kind_dims(kind, &npElem, &ndof, &ndim);
    int nsize = npElem * ndof, nbad = 0;
#ifdef _OPENMP
    if (threads < 1) threads = 1;
#else
    threads = 1;
#endif
#pragma omp parallel for schedule(static) num_threads(threads) reduction(+ : nbad) if (threads > 1)
    for (int ee = 0; ee < nElem; ee++) {
        if (elem_mask && !elem_mask[ee]) continue;
        double xn[4], yn[4], zn[4] = {0, 0, 0, 0}, K[144], F[12];
        int dofs[12], nodes[4];
        for (int ii = 0; ii < npElem; ii++) {        /* :832-838: coords stay in OLD numbering */
            nodes[ii] = conn[(size_t)ii * nElem + ee];
            int n1 = node_map_get_old ? node_map_get_old[nodes[ii] - 1] : nodes[ii];
            xn[ii] = coords[(size_t)0 * nNode + (n1 - 1)];
            yn[ii] = coords[(size_t)1 * nNode + (n1 - 1)];
            if (ndim == 3) zn[ii] = coords[(size_t)2 * nNode + (n1 - 1)];
        }
        if (element_ke(kind, xn, yn, zn, elemData, timeData, K, F)) { nbad++; continue; }
        for (int k = 0; k < nsize; k++) dofs[k] = edof[(size_t)k * nElem + ee];
        /* :851 MatSetValues(ADD_VALUES): row-major read of the column-major block */
        for (int i = 0; i < nsize; i++) {
            int r = dofs[i];
            if (r < 0 || r < row_lo || r >= row_hi) continue;
            for (int j = 0; j < nsize; j++) {
                int c = dofs[j];
                if (c < 0) continue;
                long long s = find_slot(rowptr, col, r, c);
                if (s < 0) continue;                  /* cannot happen for a pattern built from the same edof */
                double v = K[j + nsize * i];          /* Klocal(j,i) */
                if (threads > 1) {
#pragma omp atomic
                    val[s] += v;
                } else
                    val[s] = val[s] + v;
            }
        }
        /* :859-870 (ndof=1) / tetraelasticityparallelimpl1.F:938-958 (ndof>1): lifting */
        for (int ii = 0; ii < nsize; ii++) {
            if (dofs[ii] != -1) continue;
            int node = nodes[ii / ndof], d = ii % ndof;
            double fact = solnApplied[(size_t)(node - 1) * ndof + d];
            for (int jj = 0; jj < nsize; jj++)
                if (dofs[jj] != -1) F[jj] = F[jj] - K[jj + nsize * ii] * fact;
        }
        /* :880 VecSetValues(ADD_VALUES), VEC_IGNORE_NEGATIVE_INDICES */
        for (int i = 0; i < nsize; i++) {
            int r = dofs[i];
            if (r < 0 || r < row_lo || r >= row_hi) continue;
            if (threads > 1) {
#pragma omp atomic
                rhs[r] += F[i];
            } else
                rhs[r] = rhs[r] + F[i];
        }
    }
    return nbad;
}

/* ForceBC add, tetraelasticityparallelimpl1.F:971-982 (reference row formula: node-based, ignores
 * eliminated DOFs; 0-based row compared with the 1-based row_start/row_end).  fix != 0 uses the
 * NodeDofArrayNew index instead (documented-intent switch). */
ORC_API void orc_add_force_bc(int nFBC, const int *fbc_node, const int *fbc_dof, const double *fbc_val,
                              int ndof, int nNode, const int *node_map_get_new, const int *NodeDofArrayNew,
                              int row_start, int row_end, int size_global, int fix, double *rhs)
{
    for (int ii = 0; ii < nFBC; ii++) {
        int n1 = node_map_get_new[fbc_node[ii] - 1], n2 = fbc_dof[ii];
        if (fix) {
            int d = NodeDofArrayNew[(n1 - 1) + (size_t)nNode * (n2 - 1)];
            if (d >= 1 && d >= row_start && d <= row_end) rhs[d - 1] = rhs[d - 1] + fbc_val[ii];
        } else {
            int row = (n1 - 1) * ndof + n2 - 1;
            if (row >= row_start && row <= row_end && row < size_global) rhs[row] = rhs[row] + fbc_val[ii];
        }
    }
}

/* ------------------------------------------------------------------------- */
/* KSPSolve: CG + Jacobi, PETSc 3.6 semantics (solverpetsc.F:431-490 -> KSPSolve_CG)             */
/* left preconditioning, KSP_NORM_PRECONDITIONED, zero initial guess, KSPConvergedDefault.        */
/* reason codes: 2 RTOL, 3 ATOL, -3 ITS, -4 DTOL, -8 INDEFINITE_PC, -9 NANORINF, -10 INDEFINITE_MAT */
/* ------------------------------------------------------------------------- */

static void spmv(int N, const int *rowptr, const int *col, const double *val, const double *x, double *y, int threads)
{
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
    for (int r = 0; r < N; r++) {
        double s = 0.0;
        for (int k = rowptr[r]; k < rowptr[r + 1]; k++) s += val[k] * x[col[k]];
        y[r] = s;
    }
}

static double dot(int N, const double *a, const double *b, int threads)
{
    double s = 0.0;
#pragma omp parallel for schedule(static) num_threads(threads) reduction(+ : s) if (threads > 1)
    for (int i = 0; i < N; i++) s += a[i] * b[i];
    return s;
}

static int converged_default(int it, double rnorm, double rtol, double abstol, double dtol, double *ttol, double *rnorm0)
{
    if (it == 0) { *ttol = fmax(rtol * rnorm, abstol); *rnorm0 = rnorm; }
    if (isnan(rnorm) || isinf(rnorm)) return -9;
    if (rnorm <= *ttol) return rnorm < abstol ? 3 : 2;
    if (rnorm >= dtol * (*rnorm0)) return -4;
    return 0;
}

/* fixed_its > 0: run exactly that many iterations without convergence tests (timing baseline only). */
ORC_API int orc_cg_jacobi(int N, const int *rowptr, const int *col, const double *val, const double *b,
                          double *x, double rtol, double abstol, double dtol, int max_it, int threads,
                          int fixed_its, int *its_out, int *reason_out, double *rnorm_out)
{
    double *r = malloc(sizeof(double) * (size_t)N), *z = malloc(sizeof(double) * (size_t)N);
    double *p = malloc(sizeof(double) * (size_t)N), *w = malloc(sizeof(double) * (size_t)N);
    double *dinv = malloc(sizeof(double) * (size_t)N);
    if (threads < 1) threads = 1;
    /* PCSetUp_Jacobi: reciprocal of the diagonal, zero (or missing) diagonal -> 1 */
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
    for (int i = 0; i < N; i++) {
        long long s = find_slot(rowptr, col, i, i);
        double d = s >= 0 ? val[s] : 0.0;
        dinv[i] = d == 0.0 ? 1.0 : 1.0 / d;
        x[i] = 0.0;                                   /* solverpetsc.F:459 VecZeroEntries(solnVec) */
        r[i] = b[i];
        z[i] = r[i] * dinv[i];
    }
    double dp = sqrt(dot(N, z, z, threads)), ttol = 0, rnorm0 = 0;
    int reason = fixed_its > 0 ? 0 : converged_default(0, dp, rtol, abstol, dtol, &ttol, &rnorm0);
    int its = 0;
    if (!reason) {
        double beta = dot(N, z, r, threads), betaold = 0.0, dpi = 0.0, dpiold;
        int i = 0;
        do {
            its = i + 1;
            if (beta == 0.0) { reason = 3; break; }
            if (i > 0 && beta * betaold < 0.0) { reason = -8; break; }
            if (i == 0) {
                memcpy(p, z, sizeof(double) * (size_t)N);
            } else {
                double bb = beta / betaold;
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
                for (int k = 0; k < N; k++) p[k] = z[k] + bb * p[k];
            }
            dpiold = dpi;
            spmv(N, rowptr, col, val, p, w, threads);
            dpi = dot(N, p, w, threads);
            betaold = beta;
            if (dpi == 0.0 || (i > 0 && dpi * dpiold <= 0.0)) { reason = -10; break; }
            double a = beta / dpi;
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
            for (int k = 0; k < N; k++) {
                x[k] = x[k] + a * p[k];
                r[k] = r[k] - a * w[k];
                z[k] = r[k] * dinv[k];
            }
            dp = sqrt(dot(N, z, z, threads));
            if (fixed_its > 0) { if (its >= fixed_its) { reason = 4; break; } }
            else {
                reason = converged_default(i + 1, dp, rtol, abstol, dtol, &ttol, &rnorm0);
                if (reason) break;
            }
            beta = dot(N, z, r, threads);
            i++;
        } while (i < max_it);
        if (!reason && i >= max_it) reason = -3;
    }
    *its_out = its; *reason_out = reason; *rnorm_out = dp;
    free(r); free(z); free(p); free(w); free(dinv);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* KSPSolve with the reference's DEFAULT preconditioner: CG + PCBJACOBI (solverpetsc.F:187,206),   */
/* i.e. one block per MPI rank = the rank's diagonal block [block_start[b], block_start[b+1]),     */
/* sub-KSP preonly, sub-PC ILU(0) in natural ordering (PETSc defaults; PETSc 3.6.4 is not in the   */
/* tree: algorithm restated from MatLUFactorNumeric_SeqAIJ / MatSolve_SeqAIJ).                     */
/*   factor (row i, IKJ): for k in L(i) ascending: m = a_ik * (1/u_kk); a_ik = m;                  */
/*                         for j in U(k), j > k, j in pattern(i): a_ij -= m * u_kj                 */
/*                         pivot stored inverted (zero pivot -> reason -11)                        */
/*   solve: forward  y_i = r_i - sum_{j in L(i)} l_ij y_j   (ascending j)                          */
/*          backward z_i = (y_i - sum_{j in U(i), j > i} u_ij z_j) * (1/u_ii)   (ascending j)      */
/* Entries outside the block are ignored by the preconditioner (block Jacobi).                     */
/* ------------------------------------------------------------------------- */

ORC_API int orc_ilu0_factor(int N, const int *rowptr, const int *col, const double *val, int nblocks,
                            const int *block_start, double *fval, double *invdiag)
{
    /* fval: factor values on the CSR slots (slots outside the diagonal blocks keep the matrix value, unused) */
    memcpy(fval, val, sizeof(double) * (size_t)rowptr[N]);
    for (int b = 0; b < nblocks; b++) {
        const int lo = block_start[b], hi = block_start[b + 1];
        for (int i = lo; i < hi; i++) {
            long long di = -1;
            for (int q = rowptr[i]; q < rowptr[i + 1]; q++) {
                const int k = col[q];
                if (k < lo) continue;
                if (k >= i) { if (k == i) di = q; break; }
                const double m = fval[q] * invdiag[k];
                fval[q] = m;
                /* merge U(k) (cols > k of row k, inside the block) into row i */
                int qi = q + 1;
                for (int qk = rowptr[k]; qk < rowptr[k + 1]; qk++) {
                    const int j = col[qk];
                    if (j <= k) continue;
                    if (j >= hi) break;
                    while (qi < rowptr[i + 1] && col[qi] < j) qi++;
                    if (qi < rowptr[i + 1] && col[qi] == j) fval[qi] = fval[qi] - m * fval[qk];
                }
            }
            if (di < 0 || fval[di] == 0.0) return -11;      /* missing or zero pivot */
            invdiag[i] = 1.0 / fval[di];
        }
    }
    return 0;
}

ORC_API void orc_ilu0_solve(int N, const int *rowptr, const int *col, const double *fval, const double *invdiag,
                            int nblocks, const int *block_start, const double *r, double *z)
{
    for (int b = 0; b < nblocks; b++) {
        const int lo = block_start[b], hi = block_start[b + 1];
        for (int i = lo; i < hi; i++) {
            double sum = r[i];
            for (int q = rowptr[i]; q < rowptr[i + 1]; q++) {
                const int j = col[q];
                if (j < lo) continue;
                if (j >= i) break;
                sum = sum - fval[q] * z[j];
            }
            z[i] = sum;
        }
        for (int i = hi - 1; i >= lo; i--) {
            double sum = z[i];
            for (int q = rowptr[i]; q < rowptr[i + 1]; q++) {
                const int j = col[q];
                if (j <= i) continue;
                if (j >= hi) break;
                sum = sum - fval[q] * z[j];
            }
            z[i] = sum * invdiag[i];
        }
    }
}

ORC_API int orc_cg_bjacobi_ilu0(int N, const int *rowptr, const int *col, const double *val, const double *b,
                                double *x, int nblocks, const int *block_start, double rtol, double abstol, double dtol,
                                int max_it, int *its_out, int *reason_out, double *rnorm_out)
{
    double *r = malloc(sizeof(double) * (size_t)N), *z = malloc(sizeof(double) * (size_t)N);
    double *p = malloc(sizeof(double) * (size_t)N), *w = malloc(sizeof(double) * (size_t)N);
    double *invdiag = malloc(sizeof(double) * (size_t)N), *fval = malloc(sizeof(double) * (size_t)(rowptr[N] > 0 ? rowptr[N] : 1));
    int its = 0, reason = orc_ilu0_factor(N, rowptr, col, val, nblocks, block_start, fval, invdiag);
    double dp = 0.0, ttol = 0, rnorm0 = 0;
    for (int i = 0; i < N; i++) { x[i] = 0.0; r[i] = b[i]; }
    if (!reason) {
        orc_ilu0_solve(N, rowptr, col, fval, invdiag, nblocks, block_start, r, z);
        dp = sqrt(dot(N, z, z, 1));
        reason = converged_default(0, dp, rtol, abstol, dtol, &ttol, &rnorm0);
    }
    if (!reason) {
        double beta = dot(N, z, r, 1), betaold = 0.0, dpi = 0.0, dpiold;
        int i = 0;
        do {
            its = i + 1;
            if (beta == 0.0) { reason = 3; break; }
            if (i > 0 && beta * betaold < 0.0) { reason = -8; break; }
            if (i == 0) memcpy(p, z, sizeof(double) * (size_t)N);
            else {
                double bb = beta / betaold;
                for (int k = 0; k < N; k++) p[k] = z[k] + bb * p[k];
            }
            dpiold = dpi;
            spmv(N, rowptr, col, val, p, w, 1);
            dpi = dot(N, p, w, 1);
            betaold = beta;
            if (dpi == 0.0 || (i > 0 && dpi * dpiold <= 0.0)) { reason = -10; break; }
            double a = beta / dpi;
            for (int k = 0; k < N; k++) { x[k] = x[k] + a * p[k]; r[k] = r[k] - a * w[k]; }
            orc_ilu0_solve(N, rowptr, col, fval, invdiag, nblocks, block_start, r, z);
            dp = sqrt(dot(N, z, z, 1));
            reason = converged_default(i + 1, dp, rtol, abstol, dtol, &ttol, &rnorm0);
            if (reason) break;
            beta = dot(N, z, r, 1);
            i++;
        } while (i < max_it);
        if (!reason && i >= max_it) reason = -3;
    }
    *its_out = its; *reason_out = reason; *rnorm_out = dp;
    free(r); free(z); free(p); free(w); free(invdiag); free(fval);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Explicit dynamics: matrix-free element residual, lumped mass, central-difference   */
/* time loop of triaelasticityexplicit.F.  Single-precision literals (1.0/3.0, 1.0/6.0) reproduced.         */
/* 3-D routines under the documented intent of SURVEY.md 8c (ETYPE 4, one Gauss point).                     */
/* kind: 2 = tria (plane strain), 3 = tet.  Return 1 for a negative Jacobian (the reference STOPs).          */
/* ------------------------------------------------------------------------- */

/* elementutilitieselasticity2D.F:158-275 ResidualElasticityLinearTria */
ORC_API int orc_residual_elasticity_tria(const double *x, const double *y, const double *elemData, const double *timeData,
                                         const double *dispC, const double *veloC, double *Flocal)
{
    (void)timeData; (void)veloC;                       /* af, timefact and veloC are read but unused (:213-214) */
    double E = elemData[0], nu = elemData[1], dens = elemData[2], thick = 1.0;    /* :188-194 */
    double bforce[2] = {elemData[3], elemData[4]};
    double Dmat[3][3];
    double b1 = E / ((1.0 + nu) * (1.0 - 2.0 * nu));   /* plane strain, :203-206 */
    Dmat[0][0] = b1 * (1.0 - nu); Dmat[0][1] = b1 * nu;         Dmat[0][2] = 0.0;
    Dmat[1][0] = b1 * nu;         Dmat[1][1] = b1 * (1.0 - nu); Dmat[1][2] = 0.0;
    Dmat[2][0] = 0.0;             Dmat[2][1] = 0.0;             Dmat[2][2] = b1 * (1.0 - 2.0 * nu) * 0.5;
    double param[2] = {(double)(1.0f / 3.0f), (double)(1.0f / 3.0f)}, gw = 0.5;   /* :219 */
    double N[3], dNdx[3], dNdy[3], Jac;
    for (int i = 0; i < 6; i++) Flocal[i] = 0.0;
    basis2d_tria(param, x, y, N, dNdx, dNdy, &Jac);
    if (Jac < 0.0) return 1;                           /* :235-237 */
    double dvol = gw * (Jac * thick);                  /* :239 */
    double grad[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int ii = 0; ii < 3; ii++) {                   /* :244-254 */
        double c1 = dispC[2 * ii], c2 = dispC[2 * ii + 1];
        grad[0][0] = grad[0][0] + c1 * dNdx[ii];
        grad[0][1] = grad[0][1] + c1 * dNdy[ii];
        grad[1][0] = grad[1][0] + c2 * dNdx[ii];
        grad[1][1] = grad[1][1] + c2 * dNdy[ii];
    }
    double strain[3] = {grad[0][0], grad[1][1], 0.5 * (grad[0][1] + grad[1][0])}, stress[3];   /* :257-259 */
    for (int i = 0; i < 3; i++) {                      /* MATMUL(Dmat, strain) :261 */
        double sacc = 0.0;
        for (int j = 0; j < 3; j++) sacc = sacc + Dmat[i][j] * strain[j];
        stress[i] = sacc;
    }
    for (int ii = 0; ii < 3; ii++) {                   /* :264-274 */
        double c1 = dvol * dNdx[ii], c2 = dvol * dNdy[ii], c4 = (dens * dvol) * N[ii];
        Flocal[2 * ii]     = ((Flocal[2 * ii]     + c4 * bforce[0]) - c1 * stress[0]) - c2 * stress[2];
        Flocal[2 * ii + 1] = ((Flocal[2 * ii + 1] + c4 * bforce[1]) - c1 * stress[2]) - c2 * stress[1];
    }
    return 0;
}

/* elementutilitieselasticity2D.F:283-362 MassMatrixLinearTria: row sums of the consistent mass = lumped mass */
ORC_API int orc_mass_matrix_tria(const double *x, const double *y, const double *elemData, double *Mlocal)
{
    double dens = elemData[2];
    double param[2] = {(double)(1.0f / 3.0f), (double)(1.0f / 3.0f)}, gw = 0.5;
    double N[3], dNdx[3], dNdy[3], Jac, K[6][6];
    memset(K, 0, sizeof K);
    basis2d_tria(param, x, y, N, dNdx, dNdy, &Jac);
    if (Jac < 0.0) return 1;
    double dvol = gw * Jac;                            /* :325 */
    for (int ii = 0; ii < 3; ii++) {
        double b4 = (dens * dvol) * N[ii];             /* :333 */
        for (int jj = 0; jj < 3; jj++) {
            double fact = b4 * N[jj];
            K[2 * ii][2 * jj] = K[2 * ii][2 * jj] + fact;
            K[2 * ii + 1][2 * jj + 1] = K[2 * ii + 1][2 * jj + 1] + fact;
        }
    }
    for (int ii = 0; ii < 6; ii++) {                   /* :352-359 */
        double fact = 0.0;
        for (int jj = 0; jj < 6; jj++) fact = fact + K[ii][jj];
        Mlocal[ii] = fact;
    }
    return 0;
}

/* elementutilitieselasticity3D.F:575-723 ResidualElasticityLinearTetra (ETYPE 4 intent) */
ORC_API int orc_residual_elasticity_tet(const double *x, const double *y, const double *z, const double *elemData,
                                        const double *timeData, const double *valC, const double *valDotC, double *Flocal)
{
    (void)timeData; (void)valDotC;
    double E = elemData[0], nu = elemData[1];
    double bforce[3] = {elemData[3], elemData[4], elemData[5]};
    double b1 = E / ((1.0 + nu) * (1.0 - 2.0 * nu)), b2 = (1.0 - 2.0 * nu) / 2.0;      /* :617-618 */
    double Dmat[6][6];
    memset(Dmat, 0, sizeof Dmat);
    Dmat[0][0] = b1 * (1.0 - nu); Dmat[0][1] = b1 * nu;         Dmat[0][2] = b1 * nu;
    Dmat[1][0] = b1 * nu;         Dmat[1][1] = b1 * (1.0 - nu); Dmat[1][2] = b1 * nu;
    Dmat[2][0] = b1 * nu;         Dmat[2][1] = b1 * nu;         Dmat[2][2] = b1 * (1.0 - nu);
    Dmat[3][3] = b1 * b2; Dmat[4][4] = b1 * b2; Dmat[5][5] = b1 * b2;
    double param[3] = {0.25, 0.25, 0.25}, gw = (double)(1.0f / 6.0f);                  /* :637-638 */
    double N[4], dNdx[4], dNdy[4], dNdz[4], Jac;
    for (int i = 0; i < 12; i++) Flocal[i] = 0.0;
    basis3d_tet(param, x, y, z, N, dNdx, dNdy, dNdz, &Jac);
    if (Jac < 0.0) return 1;
    double dvol = gw * Jac;                            /* :657 */
    double grad[3][3];
    memset(grad, 0, sizeof grad);
    for (int ii = 0; ii < 4; ii++) {                   /* :661-679 */
        double c[3] = {valC[3 * ii], valC[3 * ii + 1], valC[3 * ii + 2]};
        for (int r = 0; r < 3; r++) {
            grad[r][0] = grad[r][0] + c[r] * dNdx[ii];
            grad[r][1] = grad[r][1] + c[r] * dNdy[ii];
            grad[r][2] = grad[r][2] + c[r] * dNdz[ii];
        }
    }
    double strain[6] = {grad[0][0], grad[1][1], grad[2][2], 0.5 * (grad[0][1] + grad[1][0]), 0.5 * (grad[1][2] + grad[2][1]),
                        0.5 * (grad[0][2] + grad[2][0])}, stress[6];                   /* :682-687 */
    for (int i = 0; i < 6; i++) {                      /* MATMUL(Dmat, strain) :689 */
        double sacc = 0.0;
        for (int j = 0; j < 6; j++) sacc = sacc + Dmat[i][j] * strain[j];
        stress[i] = sacc;
    }
    for (int ii = 0; ii < 4; ii++) {                   /* :705-721 */
        double c1 = dvol * dNdx[ii], c2 = dvol * dNdy[ii], c3 = dvol * dNdz[ii], c4 = dvol * N[ii];
        double *F = Flocal + 3 * ii;
        F[0] = F[0] + c4 * bforce[0];
        F[1] = F[1] + c4 * bforce[1];
        F[2] = F[2] + c4 * bforce[2];
        F[0] = F[0] - ((c1 * stress[0] + c2 * stress[3]) + c3 * stress[5]);
        F[1] = F[1] - ((c1 * stress[3] + c2 * stress[1]) + c3 * stress[4]);
        F[2] = F[2] - ((c1 * stress[5] + c2 * stress[4]) + c3 * stress[2]);
    }
    return 0;
}

/* elementutilitieselasticity3D.F:401-482 MassMatrixLinearTetra (one Gauss point intent) */
ORC_API int orc_mass_matrix_tet(const double *x, const double *y, const double *z, const double *elemData, double *Mlocal)
{
    double dens = elemData[2];
    double param[3] = {0.25, 0.25, 0.25}, gw = (double)(1.0f / 6.0f);
    double N[4], dNdx[4], dNdy[4], dNdz[4], Jac, K[12][12];
    memset(K, 0, sizeof K);
    basis3d_tet(param, x, y, z, N, dNdx, dNdy, dNdz, &Jac);
    if (Jac < 0.0) return 1;
    double dvol = gw * (Jac * dens);                   /* :446 */
    for (int ii = 0; ii < 4; ii++) {
        double b4 = dvol * N[ii];
        for (int jj = 0; jj < 4; jj++) {
            double fact = b4 * N[jj];
            for (int d = 0; d < 3; d++) K[3 * ii + d][3 * jj + d] = K[3 * ii + d][3 * jj + d] + fact;
        }
    }
    for (int ii = 0; ii < 12; ii++) {
        double fact = 0.0;
        for (int jj = 0; jj < 12; jj++) fact = fact + K[ii][jj];
        Mlocal[ii] = fact;
    }
    return 0;
}

static int explicit_gather(int kind, int e, int nElem, const int *conn, int nNode, const double *coords, double *xn, double *yn,
                           double *zn, int *nodes)
{
    int npe = kind == 2 ? 3 : 4;
    for (int ii = 0; ii < npe; ii++) {
        int n1 = conn[(size_t)ii * nElem + e] - 1;      /* np = 1: node_map_get_old is the identity */
        nodes[ii] = n1;
        xn[ii] = coords[n1]; yn[ii] = coords[(size_t)nNode + n1];
        if (kind == 3) zn[ii] = coords[2 * (size_t)nNode + n1];
    }
    return npe;
}

/* triaelasticityexplicit.F:881-921: globalM(node dof) += Mlocal, element by element */
ORC_API int orc_explicit_lumped_mass(int kind, int nElem, const int *conn, int nNode, const double *coords,
                                     const double *elemData, double *globalM)
{
    int ndof = kind == 2 ? 2 : 3, nbad = 0;
    for (size_t i = 0; i < (size_t)nNode * ndof; i++) globalM[i] = 0.0;
    for (int e = 0; e < nElem; e++) {
        double xn[4], yn[4], zn[4], Ml[12];
        int nodes[4];
        int npe = explicit_gather(kind, e, nElem, conn, nNode, coords, xn, yn, zn, nodes);
        nbad += kind == 2 ? orc_mass_matrix_tria(xn, yn, elemData, Ml) : orc_mass_matrix_tet(xn, yn, zn, elemData, Ml);
        for (int ii = 0; ii < npe; ii++)
            for (int d = 0; d < ndof; d++) {
                size_t g = (size_t)nodes[ii] * ndof + d;
                globalM[g] = globalM[g] + Ml[ii * ndof + d];
            }
    }
    return nbad;
}

/* triaelasticityexplicit.F:972-1118: nsteps central-difference steps with constant elemData.  disp == dispPrev on entry  */
/* (the loop's own invariant after its first pass); free_slots = assyForSoln (1-based node-dof slots of the free dofs).  */
ORC_API int orc_explicit_advance(int kind, int nElem, const int *conn, int nNode, const double *coords, int size_global,
                                 const int *free_slots, const double *elemData, const double *timeData, double dt, int nsteps,
                                 const double *globalM, double *disp, double *dispPrev, double *dispPrev2, double *velo,
                                 double *acce)
{
    int ndof = kind == 2 ? 2 : 3, nbad = 0;
    size_t nd = (size_t)nNode * ndof;
    double *rhs = malloc(sizeof(double) * nd);
    double DTT = dt * dt, IDTT = 1.0 / DTT;            /* :961-962 */
    for (int step = 0; step < nsteps; step++) {
        for (size_t i = 0; i < nd; i++) rhs[i] = 0.0;  /* :994 */
        for (int e = 0; e < nElem; e++) {              /* :996-1057 */
            double xn[4], yn[4], zn[4], de[12], ve[12], Fl[12];
            int nodes[4];
            int npe = explicit_gather(kind, e, nElem, conn, nNode, coords, xn, yn, zn, nodes);
            for (int ii = 0; ii < npe; ii++)
                for (int d = 0; d < ndof; d++) {
                    de[ii * ndof + d] = disp[(size_t)nodes[ii] * ndof + d];
                    ve[ii * ndof + d] = velo[(size_t)nodes[ii] * ndof + d];
                }
            nbad += kind == 2 ? orc_residual_elasticity_tria(xn, yn, elemData, timeData, de, ve, Fl)
                              : orc_residual_elasticity_tet(xn, yn, zn, elemData, timeData, de, ve, Fl);
            for (int ii = 0; ii < npe; ii++)
                for (int d = 0; d < ndof; d++) {
                    size_t g = (size_t)nodes[ii] * ndof + d;
                    rhs[g] = rhs[g] + Fl[ii * ndof + d];
                }
        }
        for (int ii = 0; ii < size_global; ii++) {     /* :1072-1079: free dofs only; Dirichlet dofs stay at zero */
            size_t jj = (size_t)free_slots[ii] - 1;
            rhs[jj] = rhs[jj] + IDTT * globalM[jj] * (2.0 * dispPrev[jj] - dispPrev2[jj]);
            disp[jj] = (DTT * rhs[jj]) / globalM[jj];
        }
        for (size_t i = 0; i < nd; i++) {              /* :1084-1085, then :1118-1121 */
            velo[i] = (disp[i] - dispPrev2[i]) / (2.0 * dt);
            acce[i] = (disp[i] - 2.0 * dispPrev[i] + dispPrev2[i]) / DTT;
        }
        for (size_t i = 0; i < nd; i++) { dispPrev2[i] = dispPrev[i]; dispPrev[i] = disp[i]; }
    }
    free(rhs);
    return nbad;
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
