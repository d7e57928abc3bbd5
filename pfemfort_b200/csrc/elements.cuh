// elements.cuh -- P1 triangle / tetrahedron element arithmetic for sm_100a (FP64).
//
// Device-side restatement of the reference element utilities:
//   elementutilitiesbasisfuncs.F:39-51,165-234 (tria), :261-281,430-538 (tet)
//   elementutilitiespoisson.F:23-101 (tria), :107-193 (tet)
//   elementutilitieselasticity2D.F:23-153, elementutilitieselasticity3D.F:248-393
//
// Every translation unit that includes this file is compiled with -fmad=false: the expressions
// below follow the reference's evaluation order and must not be contracted
// into FMAs, so that Ke/Fe reproduce the gfortran (no-FMA x86-64) results bit for bit.
// Multiplications by the constant parametric gradients (0, +-1) are exact and are folded away.
#pragma once
#ifndef PFEM_EMULATE            // tests/emu compiles this header for the host through a small CUDA shim
#include <cuda_runtime.h>
#endif

// keep a value in its register: stops ptxas from rematerialising the product per use (device code only)
#if defined(__CUDA_ARCH__)
#define PFEM_KEEP2(a, b) asm volatile("" : "+d"(a), "+d"(b))
#define PFEM_KEEP3(a, b, c) asm volatile("" : "+d"(a), "+d"(b), "+d"(c))
#else
#define PFEM_KEEP2(a, b) ((void)0)
#define PFEM_KEEP3(a, b, c) ((void)0)
#endif

namespace pfem {

enum { POISSON_TRIA = 0, POISSON_TETRA = 1, ELASTICITY_TRIA = 2, ELASTICITY_TETRA = 3 };

template <int KIND> struct ElemTraits;
template <> struct ElemTraits<POISSON_TRIA>     { static constexpr int NPE = 3, NDOF = 1, NDIM = 2, NSTR = 0; };
template <> struct ElemTraits<POISSON_TETRA>    { static constexpr int NPE = 4, NDOF = 1, NDIM = 3, NSTR = 0; };
template <> struct ElemTraits<ELASTICITY_TRIA>  { static constexpr int NPE = 3, NDOF = 2, NDIM = 2, NSTR = 3; };
template <> struct ElemTraits<ELASTICITY_TETRA> { static constexpr int NPE = 4, NDOF = 3, NDIM = 3, NSTR = 6; };

// single-precision literals of the reference, widened (gfortran without -fdefault-real-8)
__device__ __forceinline__ double third_f() { return (double)(1.0f / 3.0f); }   // 0.3333333432674408
__device__ __forceinline__ double sixth_f() { return (double)(1.0f / 6.0f); }   // 0.1666666716337204

// Runtime-indexed read of a small register array without forcing it into local memory.
template <int N> __device__ __forceinline__ double pick(const double (&v)[N], int i)
{
    double r = v[0];
#pragma unroll
    for (int q = 1; q < N; q++) r = (i == q) ? v[q] : r;
    return r;
}

// Geometry of one element: physical gradients of the NPE shape functions, their values at the single
// Gauss point, and the Jacobian determinant.
template <int NPE, int NDIM> struct Geom {
    double dN[NDIM][NPE];
    double N[NPE];
    double Jac;
};

// computeBasisFunctions2D, ETYPE=1, degree 1 at (1/3f, 1/3f)
__device__ __forceinline__ void tria_geom(const double x[3], const double y[3], Geom<3, 2> &g)
{
    const double xi = third_f();
    g.N[0] = 1.0 - xi - xi; g.N[1] = xi; g.N[2] = xi;       // basisfuncs.F:29,41-43
    const double B11 = x[1] - x[0], B21 = x[2] - x[0];      // :208-217
    const double B12 = y[1] - y[0], B22 = y[2] - y[0];
    const double Jac = B11 * B22 - B12 * B21;               // :219
    const double detinv = 1.0 / Jac;                        // :221
    const double Bi11 = B22 * detinv, Bi12 = -B12 * detinv; // :223-226
    const double Bi21 = -B21 * detinv, Bi22 = B11 * detinv;
    g.dN[0][0] = -(Bi11 + Bi12); g.dN[0][1] = Bi11; g.dN[0][2] = Bi12;   // :229-232
    g.dN[1][0] = -(Bi21 + Bi22); g.dN[1][1] = Bi21; g.dN[1][2] = Bi22;
    g.Jac = Jac;
}

// computeBasisFunctions3D, ETYPE=4, degree 1 at (1/4,1/4,1/4); local node 3 is the origin
__device__ __forceinline__ void tet_geom(const double x[4], const double y[4], const double z[4], Geom<4, 3> &g)
{
    g.N[0] = 0.25; g.N[1] = 0.25; g.N[2] = 1.0 - 0.25 - 0.25 - 0.25; g.N[3] = 0.25;
    // B(r,c): basisfuncs.F:493-509
    const double B00 = x[0] - x[2], B10 = x[1] - x[2], B20 = x[3] - x[2];
    const double B01 = y[0] - y[2], B11 = y[1] - y[2], B21 = y[3] - y[2];
    const double B02 = z[0] - z[2], B12 = z[1] - z[2], B22 = z[3] - z[2];
    double Jac = B00 * (B11 * B22 - B12 * B21);             // :512-514
    Jac = Jac + B01 * (B12 * B20 - B10 * B22);
    Jac = Jac + B02 * (B10 * B21 - B11 * B20);
    const double detinv = 1.0 / Jac;                        // :517
    const double ndet = -detinv;
    const double Bi00 = detinv * (B11 * B22 - B12 * B21);   // :520-528
    const double Bi10 = ndet * (B10 * B22 - B12 * B20);
    const double Bi20 = detinv * (B10 * B21 - B11 * B20);
    const double Bi01 = ndet * (B01 * B22 - B02 * B21);
    const double Bi11 = detinv * (B00 * B22 - B02 * B20);
    const double Bi21 = ndet * (B00 * B21 - B01 * B20);
    const double Bi02 = detinv * (B01 * B12 - B02 * B11);
    const double Bi12 = ndet * (B00 * B12 - B02 * B10);
    const double Bi22 = detinv * (B00 * B11 - B01 * B10);
    // :532-536
    g.dN[0][0] = Bi00; g.dN[0][1] = Bi01; g.dN[0][3] = Bi02; g.dN[0][2] = -((Bi00 + Bi01) + Bi02);
    g.dN[1][0] = Bi10; g.dN[1][1] = Bi11; g.dN[1][3] = Bi12; g.dN[1][2] = -((Bi10 + Bi11) + Bi12);
    g.dN[2][0] = Bi20; g.dN[2][1] = Bi21; g.dN[2][3] = Bi22; g.dN[2][2] = -((Bi20 + Bi21) + Bi22);
    g.Jac = Jac;
}

// Per-call material/time parameters, expanded once per thread.
template <int KIND> struct Params;

template <> struct Params<POISSON_TRIA> {
    double kx, ky, af, force, gw;
    __device__ __forceinline__ void init(const double *elemData, const double *timeData) {
        kx = elemData[0]; ky = elemData[1]; af = timeData[1]; force = 0.0; gw = 0.5;   // poisson.F:48,51,57,83
    }
};
template <> struct Params<POISSON_TETRA> {
    double kx, ky, kz, af, force, gw;
    __device__ __forceinline__ void init(const double *elemData, const double *timeData) {
        kx = elemData[0]; ky = elemData[1]; kz = elemData[2]; af = timeData[1];       // poisson.F:132,135
        force = -6.0; gw = sixth_f();                                                 // :172,:142
    }
};
template <> struct Params<ELASTICITY_TRIA> {
    double D[3][3], thick, bf[2], gw;
    __device__ __forceinline__ void init(const double *elemData, const double *) {
        const double E = elemData[0], nu = elemData[1];
        thick = elemData[2]; bf[0] = elemData[3]; bf[1] = elemData[4];
        const double b1 = E / (1.0 - nu * nu);                                         // elasticity2D.F:59
        D[0][0] = b1;      D[0][1] = b1 * nu; D[0][2] = 0.0;                            // :62-64
        D[1][0] = b1 * nu; D[1][1] = b1;      D[1][2] = 0.0;
        D[2][0] = 0.0;     D[2][1] = 0.0;     D[2][2] = b1 * (1.0 - nu);
        gw = 0.5;
    }
};
template <> struct Params<ELASTICITY_TETRA> {
    double D[6][6], bf[3], gw;
    __device__ __forceinline__ void init(const double *elemData, const double *) {
        const double E = elemData[0], nu = elemData[1];
        bf[0] = elemData[3]; bf[1] = elemData[4]; bf[2] = elemData[5];
        const double b1 = E / ((1.0 + nu) * (1.0 - 2.0 * nu));                         // elasticity3D.F:284
        const double b2 = (1.0 - 2.0 * nu) / 2.0;                                      // :285
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = 0; j < 6; j++) D[i][j] = 0.0;
        D[0][0] = b1 * (1.0 - nu); D[0][1] = b1 * nu;         D[0][2] = b1 * nu;        // :288-293
        D[1][0] = b1 * nu;         D[1][1] = b1 * (1.0 - nu); D[1][2] = b1 * nu;
        D[2][0] = b1 * nu;         D[2][1] = b1 * nu;         D[2][2] = b1 * (1.0 - nu);
        D[3][3] = b1 * b2; D[4][4] = b1 * b2; D[5][5] = b1 * b2;
        gw = sixth_f();                                                                 // :305
    }
};

// ------------------------------------------------------------------------------------------------
// Strain-displacement entries Bmat(s, a) for local dof a = ndof*i + d (elasticity2D.F:127-133,
// elasticity3D.F:360-371).  Written branch-free on (s, d) so that a runtime `a` costs selects only.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double bmat2(int s, int d, double dx, double dy)
{
    // rows: 0 = xx, 1 = yy, 2 = xy
    if (s == 0) return d == 0 ? dx : 0.0;
    if (s == 1) return d == 0 ? 0.0 : dy;
    return d == 0 ? dy : dx;
}
__device__ __forceinline__ double bmat3(int s, int d, double dx, double dy, double dz)
{
    // rows: 0 = xx, 1 = yy, 2 = zz, 3 = xy, 4 = yz, 5 = zx
    switch (s) {
    case 0: return d == 0 ? dx : 0.0;
    case 1: return d == 1 ? dy : 0.0;
    case 2: return d == 2 ? dz : 0.0;
    case 3: return d == 0 ? dy : (d == 1 ? dx : 0.0);
    case 4: return d == 1 ? dz : (d == 2 ? dy : 0.0);
    default: return d == 0 ? dz : (d == 2 ? dx : 0.0);
    }
}

// ------------------------------------------------------------------------------------------------
// Element operator: everything the two kernels need, per kind.
//   load_geom  : gradients + Jacobian from nodal coordinates
//   dvol       : quadrature weight * Jacobian (* thickness)
//   col_setup  : precompute what depends on the SECOND index b of Klocal(a,b)
//   K(a)       : Klocal(a, b) for the prepared b
//   F(a)       : Flocal(a) with valC = 0 (the batched drivers always pass zeros)
// ------------------------------------------------------------------------------------------------
template <int KIND> struct ElemOp;

template <> struct ElemOp<POISSON_TRIA> {
    using T = ElemTraits<POISSON_TRIA>;
    Geom<3, 2> g; double dvol; double px, py;
    __device__ __forceinline__ void load_geom(const double x[3], const double y[3], const double *) { tria_geom(x, y, g); }
    __device__ __forceinline__ void set_dvol(const Params<POISSON_TRIA> &p) { dvol = p.gw * g.Jac; }               // poisson.F:75
    __device__ __forceinline__ void col_setup(const Params<POISSON_TRIA> &p, int b) {
        px = p.kx * pick(g.dN[0], b); py = p.ky * pick(g.dN[1], b);
        PFEM_KEEP2(px, py);                        // keep the products: do not rematerialise them per entry
    }
    __device__ __forceinline__ double K(const Params<POISSON_TRIA> &p, int a) const {                              // :93-95
        const double b1 = pick(g.dN[0], a) * dvol, b2 = pick(g.dN[1], a) * dvol;
        return p.af * (b1 * px + b2 * py);
    }
    // kx = ky = af = 1.0 exactly (the drivers' constants): x*1.0 == x bit for bit, so the multiplications are skipped
    __device__ __forceinline__ void col_setup_unit(int b) { px = pick(g.dN[0], b); py = pick(g.dN[1], b); }
    __device__ __forceinline__ double K_unit(int a) const {
        const double b1 = pick(g.dN[0], a) * dvol, b2 = pick(g.dN[1], a) * dvol;
        return b1 * px + b2 * py;
    }
    // Flocal with valC = 0: the "- b*du" terms subtract exact zeros for finite gradients and are dropped
    __device__ __forceinline__ double F0(const Params<POISSON_TRIA> &p, int a) const { return (pick(g.N, a) * dvol) * p.force; }
    // Flocal(ii) = Flocal(ii) + b4*force - b1*du(1) - b2*du(2), poisson.F:90
    __device__ __forceinline__ double F(const Params<POISSON_TRIA> &p, int a, const double du[2]) const {
        const double b1 = pick(g.dN[0], a) * dvol, b2 = pick(g.dN[1], a) * dvol, b4 = pick(g.N, a) * dvol;
        return 0.0 + b4 * p.force - b1 * du[0] - b2 * du[1];
    }
};

template <> struct ElemOp<POISSON_TETRA> {
    using T = ElemTraits<POISSON_TETRA>;
    Geom<4, 3> g; double dvol; double px, py, pz;
    __device__ __forceinline__ void load_geom(const double x[4], const double y[4], const double z[4]) { tet_geom(x, y, z, g); }
    __device__ __forceinline__ void set_dvol(const Params<POISSON_TETRA> &p) { dvol = p.gw * g.Jac; }              // poisson.F:161
    __device__ __forceinline__ void col_setup(const Params<POISSON_TETRA> &p, int b) {
        px = p.kx * pick(g.dN[0], b); py = p.ky * pick(g.dN[1], b); pz = p.kz * pick(g.dN[2], b);
        PFEM_KEEP3(px, py, pz);
    }
    __device__ __forceinline__ double K(const Params<POISSON_TETRA> &p, int a) const {                             // :183-187
        const double b1 = pick(g.dN[0], a) * dvol, b2 = pick(g.dN[1], a) * dvol, b3 = pick(g.dN[2], a) * dvol;
        return p.af * (b1 * px + b2 * py + b3 * pz);
    }
    __device__ __forceinline__ void col_setup_unit(int b) { px = pick(g.dN[0], b); py = pick(g.dN[1], b); pz = pick(g.dN[2], b); }
    __device__ __forceinline__ double K_unit(int a) const {
        const double b1 = pick(g.dN[0], a) * dvol, b2 = pick(g.dN[1], a) * dvol, b3 = pick(g.dN[2], a) * dvol;
        return b1 * px + b2 * py + b3 * pz;
    }
    __device__ __forceinline__ double F0(const Params<POISSON_TETRA> &p, int a) const { return (pick(g.N, a) * dvol) * p.force; }
    // poisson.F:180-181
    __device__ __forceinline__ double F(const Params<POISSON_TETRA> &p, int a, const double du[3]) const {
        const double b1 = pick(g.dN[0], a) * dvol, b2 = pick(g.dN[1], a) * dvol, b3 = pick(g.dN[2], a) * dvol, b4 = pick(g.N, a) * dvol;
        double f = 0.0 + b4 * p.force;
        f = f - b1 * du[0] - b2 * du[1] - b3 * du[2];
        return f;
    }
};

template <> struct ElemOp<ELASTICITY_TRIA> {
    using T = ElemTraits<ELASTICITY_TRIA>;
    Geom<3, 2> g; double dvol; double DB[3];
    __device__ __forceinline__ void load_geom(const double x[3], const double y[3], const double *) { tria_geom(x, y, g); }
    __device__ __forceinline__ void set_dvol(const Params<ELASTICITY_TRIA> &p) { dvol = p.gw * (g.Jac * p.thick); } // elasticity2D.F:94
    // Column b of MATMUL(Dmat, Bmat) and entry (a,b) of MATMUL(BmatTrans, .) (:136-139).  The reference sums the inner
    // index in ascending order with the structural zeros of Bmat/Dmat included; a zero term is an exact +-0 and leaves the
    // running sum unchanged, so only the structurally non-zero terms are formed here, in the same ascending order: the
    // results are bit-identical (up to the sign of an exact zero).
    __device__ __forceinline__ void col_setup(const Params<ELASTICITY_TRIA> &p, int b) {
        const int j = b >> 1, d = b & 1;
        const double dx = pick(g.dN[0], j), dy = pick(g.dN[1], j);
        const double gd = d == 0 ? dx : dy;                       // the only non-zero of Bmat(0:1, b)
        DB[0] = (d == 0 ? p.D[0][0] : p.D[0][1]) * gd;
        DB[1] = (d == 0 ? p.D[1][0] : p.D[1][1]) * gd;
        DB[2] = p.D[2][2] * (d == 0 ? dy : dx);
    }
    // Klocal(a,b) = dvol * sum_s Bmat(s,a) * DB(s,b)  (:138-139)
    __device__ __forceinline__ double K(const Params<ELASTICITY_TRIA> &, int a) const {
        const int i = a >> 1, d = a & 1;
        const double dx = pick(g.dN[0], i), dy = pick(g.dN[1], i);
        // d = 0: rows xx (s=0) and xy (s=2); d = 1: rows yy (s=1) and xy (s=2)
        const double t1 = (d == 0 ? dx : dy) * (d == 0 ? DB[0] : DB[1]);
        const double t2 = (d == 0 ? dy : dx) * DB[2];
        return dvol * (t1 + t2);
    }
    __device__ __forceinline__ void col_setup_unit(int) {}
    __device__ __forceinline__ double K_unit(int) const { return 0.0; }
    __device__ __forceinline__ double F0(const Params<ELASTICITY_TRIA> &p, int a) const { return F(p, a, nullptr); }
    __device__ __forceinline__ double F(const Params<ELASTICITY_TRIA> &p, int a, const double *) const {            // :142-150
        const int i = a >> 1, d = a & 1;
        const double b4 = dvol * pick(g.N, i);
        return 0.0 + b4 * pick(p.bf, d);
    }
};

template <> struct ElemOp<ELASTICITY_TETRA> {
    using T = ElemTraits<ELASTICITY_TETRA>;
    Geom<4, 3> g; double dvol; double DB[6];
    __device__ __forceinline__ void load_geom(const double x[4], const double y[4], const double z[4]) { tet_geom(x, y, z, g); }
    __device__ __forceinline__ void set_dvol(const Params<ELASTICITY_TETRA> &p) { dvol = p.gw * g.Jac; }            // elasticity3D.F:324
    // see the 2-D operator: only the structurally non-zero terms of the two MATMULs, in the reference's ascending order
    __device__ __forceinline__ void col_setup(const Params<ELASTICITY_TETRA> &p, int b) {                            // :374
        const int j = b / 3, d = b - 3 * j;
        const double dx = pick(g.dN[0], j), dy = pick(g.dN[1], j), dz = pick(g.dN[2], j);
        const double gd = d == 0 ? dx : (d == 1 ? dy : dz);       // the only non-zero of Bmat(0:2, b)
#pragma unroll
        for (int s = 0; s < 3; s++) DB[s] = (d == 0 ? p.D[s][0] : (d == 1 ? p.D[s][1] : p.D[s][2])) * gd;
        DB[3] = p.D[3][3] * (d == 0 ? dy : (d == 1 ? dx : 0.0));  // xy
        DB[4] = p.D[4][4] * (d == 1 ? dz : (d == 2 ? dy : 0.0));  // yz
        DB[5] = p.D[5][5] * (d == 0 ? dz : (d == 2 ? dx : 0.0));  // zx
    }
    __device__ __forceinline__ double K(const Params<ELASTICITY_TETRA> &, int a) const {                             // :376-377
        const int i = a / 3, d = a - 3 * i;
        const double dx = pick(g.dN[0], i), dy = pick(g.dN[1], i), dz = pick(g.dN[2], i);
        // non-zero rows of Bmat(:, a), ascending: d=0: xx(0), xy(3), zx(5); d=1: yy(1), xy(3), yz(4); d=2: zz(2), yz(4), zx(5)
        const double t1 = (d == 0 ? dx : (d == 1 ? dy : dz)) * (d == 0 ? DB[0] : (d == 1 ? DB[1] : DB[2]));
        const double t2 = (d == 0 ? dy : (d == 1 ? dx : dy)) * (d == 2 ? DB[4] : DB[3]);
        const double t3 = (d == 2 ? dx : dz) * (d == 1 ? DB[4] : DB[5]);
        return dvol * ((t1 + t2) + t3);
    }
    __device__ __forceinline__ void col_setup_unit(int) {}
    __device__ __forceinline__ double K_unit(int) const { return 0.0; }
    __device__ __forceinline__ double F0(const Params<ELASTICITY_TETRA> &p, int a) const { return F(p, a, nullptr); }
    __device__ __forceinline__ double F(const Params<ELASTICITY_TETRA> &p, int a, const double *) const {            // :380-390
        const int i = a / 3, d = a - 3 * i;
        const double b4 = dvol * pick(g.N, i);
        return 0.0 + b4 * pick(p.bf, d);
    }
};

}  // namespace pfem
