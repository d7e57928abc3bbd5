// elements_fast.cuh -- the FMA ("fast") element operators of the value pass.
//
// Same interface as ElemOp<KIND> (elements.cuh), same mathematics (elementutilitiespoisson.F:23-193,
// elementutilitieselasticity2D.F:23-153, elementutilitieselasticity3D.F:248-393), but NOT the reference's evaluation order:
// translation units that include this header are compiled WITH FMA contraction, and the Poisson operators use the
// cofactor form of the P1 stiffness
//     Klocal(a,b) = af * dvol * sum_d k_d dN_d(a) dN_d(b) = (af * gw / Jac) * sum_d k_d c_d(a) c_d(b),
// c(a) = Jac * grad N_a (the cofactor vectors of the edge matrix B, basisfuncs.F:208-226,493-536), so the inverse Jacobian is
// applied once per element instead of once per gradient entry.  Results agree with the reference evaluation order to
// rounding (a few ulp of the largest term); the contract is 1e-12 relative (BASELINE.json north_star), tested at 1e-12
// against the no-FMA oracle.  The reference-order operators stay available (PFEM_ASM=rows / PFEM_ASM_ROWS).
#pragma once
#include "elements.cuh"

namespace pfem {

template <int KIND> struct FastOp : ElemOp<KIND> {};       // elasticity kinds: the reference-order operator, FMA-contracted

template <> struct FastOp<POISSON_TETRA> {
    struct { double Jac; } g;
    double c[3][4], s, f0, px, py, pz;
    __device__ __forceinline__ void load_geom(const double x[4], const double y[4], const double z[4])
    {
        // rows of B: local nodes 0, 1, 3 relative to node 2 (basisfuncs.F:493-509)
        const double ax = x[0] - x[2], ay = y[0] - y[2], az = z[0] - z[2];
        const double bx = x[1] - x[2], by = y[1] - y[2], bz = z[1] - z[2];
        const double dx = x[3] - x[2], dy = y[3] - y[2], dz = z[3] - z[2];
        c[0][0] = by * dz - bz * dy; c[1][0] = bz * dx - bx * dz; c[2][0] = bx * dy - by * dx;   // b x d
        c[0][1] = dy * az - dz * ay; c[1][1] = dz * ax - dx * az; c[2][1] = dx * ay - dy * ax;   // d x a
        c[0][3] = ay * bz - az * by; c[1][3] = az * bx - ax * bz; c[2][3] = ax * by - ay * bx;   // a x b
#pragma unroll
        for (int d = 0; d < 3; d++) c[d][2] = -((c[d][0] + c[d][1]) + c[d][3]);
        g.Jac = ax * c[0][0] + ay * c[1][0] + az * c[2][0];                                      // :512-514
    }
    __device__ __forceinline__ void set_dvol(const Params<POISSON_TETRA> &p)
    {
        s = (p.af * p.gw) / g.Jac;
        f0 = (0.25 * (p.gw * g.Jac)) * p.force;               // N_a = 1/4 at the Gauss point (poisson.F:172-181, valC = 0)
    }
    __device__ __forceinline__ void col_setup(const Params<POISSON_TETRA> &p, int b)
    {
        px = (s * p.kx) * pick(c[0], b); py = (s * p.ky) * pick(c[1], b); pz = (s * p.kz) * pick(c[2], b);
    }
    __device__ __forceinline__ void col_setup_unit(int b) { px = s * pick(c[0], b); py = s * pick(c[1], b); pz = s * pick(c[2], b); }
    __device__ __forceinline__ double K(const Params<POISSON_TETRA> &, int a) const { return K_unit(a); }
    __device__ __forceinline__ double K_unit(int a) const { return pick(c[0], a) * px + pick(c[1], a) * py + pick(c[2], a) * pz; }
    __device__ __forceinline__ double F0(const Params<POISSON_TETRA> &, int) const { return f0; }
};

template <> struct FastOp<POISSON_TRIA> {
    struct { double Jac; } g;
    double c[2][3], s, f0, px, py;
    __device__ __forceinline__ void load_geom(const double x[3], const double y[3], const double *)
    {
        const double ax = x[1] - x[0], ay = y[1] - y[0];      // basisfuncs.F:208-217
        const double bx = x[2] - x[0], by = y[2] - y[0];
        g.Jac = ax * by - ay * bx;
        c[0][1] = by;  c[1][1] = -bx;
        c[0][2] = -ay; c[1][2] = ax;
        c[0][0] = -(c[0][1] + c[0][2]); c[1][0] = -(c[1][1] + c[1][2]);
    }
    __device__ __forceinline__ void set_dvol(const Params<POISSON_TRIA> &p)
    {
        s = (p.af * p.gw) / g.Jac;
        f0 = (p.gw * g.Jac) * p.force;                         // times N_a (poisson.F:83-90, valC = 0)
    }
    __device__ __forceinline__ void col_setup(const Params<POISSON_TRIA> &p, int b) { px = (s * p.kx) * pick(c[0], b); py = (s * p.ky) * pick(c[1], b); }
    __device__ __forceinline__ void col_setup_unit(int b) { px = s * pick(c[0], b); py = s * pick(c[1], b); }
    __device__ __forceinline__ double K(const Params<POISSON_TRIA> &, int a) const { return K_unit(a); }
    __device__ __forceinline__ double K_unit(int a) const { return pick(c[0], a) * px + pick(c[1], a) * py; }
    __device__ __forceinline__ double F0(const Params<POISSON_TRIA> &, int a) const
    {
        const double xi = third_f();
        return (a == 0 ? 1.0 - xi - xi : xi) * f0;
    }
};

}  // namespace pfem
