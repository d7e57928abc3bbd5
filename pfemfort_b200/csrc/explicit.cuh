// explicit.cuh -- device code of the explicit-dynamics path (see explicit.cu for the reference map): per-node element
// residual / lumped-mass arithmetic in the reference's evaluation order, and the two node-gather kernels.
// Kept in a header so that tests/emu can compile the very same source for the host (tests/test_explicit_emu.py compares it
// bit for bit with the oracle in the CPU suite).  Every TU that includes it is compiled without FMA contraction.
#pragma once
#include "elements.cuh"

namespace pfem {

// ---- element arithmetic -------------------------------------------------------------------------------------------

struct ExplicitParams { double E, nu, dens, b[3]; };

// residual of local node `li` of a triangle: Flocal(2 li - 1 : 2 li)   (elasticity2D.F:158-275)
__device__ __forceinline__ void residual_node_tria(const double x[3], const double y[3], const double u[6], const ExplicitParams &p,
                                                   int li, double F[2], bool &neg)
{
    Geom<3, 2> g;
    tria_geom(x, y, g);
    neg = g.Jac < 0.0;
    const double b1 = p.E / ((1.0 + p.nu) * (1.0 - 2.0 * p.nu));         // plane strain, :203-206
    const double D11 = b1 * (1.0 - p.nu), D12 = b1 * p.nu, D33 = b1 * (1.0 - 2.0 * p.nu) * 0.5;
    const double dvol = 0.5 * (g.Jac * 1.0);                            // gwts * (Jac * thick), :239
    double g00 = 0.0, g01 = 0.0, g10 = 0.0, g11 = 0.0;
#pragma unroll
    for (int ii = 0; ii < 3; ii++) {                                    // :244-254
        const double c1 = u[2 * ii], c2 = u[2 * ii + 1];
        g00 = g00 + c1 * g.dN[0][ii];
        g01 = g01 + c1 * g.dN[1][ii];
        g10 = g10 + c2 * g.dN[0][ii];
        g11 = g11 + c2 * g.dN[1][ii];
    }
    const double e0 = g00, e1 = g11, e2 = 0.5 * (g01 + g10);            // :257-259
    // MATMUL(Dmat, strain): inner index ascending; the structurally zero terms add +0.0
    const double s0 = ((0.0 + D11 * e0) + D12 * e1) + 0.0 * e2;
    const double s1 = ((0.0 + D12 * e0) + D11 * e1) + 0.0 * e2;
    const double s2 = ((0.0 + 0.0 * e0) + 0.0 * e1) + D33 * e2;
    const double dnx = pick(g.dN[0], li), dny = pick(g.dN[1], li), Ni = pick(g.N, li);
    const double c1 = dvol * dnx, c2 = dvol * dny, c4 = (p.dens * dvol) * Ni;   // :268-270
    F[0] = ((0.0 + c4 * p.b[0]) - c1 * s0) - c2 * s2;                   // :272-273
    F[1] = ((0.0 + c4 * p.b[1]) - c1 * s2) - c2 * s1;
}

// lumped mass of local node li of a triangle (same value for both dofs): row sum of the consistent mass, :283-362
__device__ __forceinline__ double mass_node_tria(const double x[3], const double y[3], const ExplicitParams &p, int li, bool &neg)
{
    Geom<3, 2> g;
    tria_geom(x, y, g);
    neg = g.Jac < 0.0;
    const double dvol = 0.5 * g.Jac;                                    // :325
    const double b4 = (p.dens * dvol) * pick(g.N, li);                  // :333
    double fact = 0.0;                                                  // the zero columns in between add +0.0
    fact = fact + b4 * g.N[0];
    fact = fact + b4 * g.N[1];
    fact = fact + b4 * g.N[2];
    return fact;
}

// residual of local node li of a tetrahedron: Flocal(3 li - 2 : 3 li)   (elasticity3D.F:575-723)
__device__ __forceinline__ void residual_node_tet(const double x[4], const double y[4], const double z[4], const double u[12],
                                                  const ExplicitParams &p, int li, double F[3], bool &neg)
{
    Geom<4, 3> g;
    tet_geom(x, y, z, g);
    neg = g.Jac < 0.0;
    const double b1 = p.E / ((1.0 + p.nu) * (1.0 - 2.0 * p.nu)), b2 = (1.0 - 2.0 * p.nu) / 2.0;   // :617-618
    const double Dd = b1 * (1.0 - p.nu), Do = b1 * p.nu, Ds = b1 * b2;
    const double dvol = sixth_f() * g.Jac;                              // :657
    double gr[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
    for (int ii = 0; ii < 4; ii++) {                                    // :661-679
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const double c = u[3 * ii + r];
            gr[r][0] = gr[r][0] + c * g.dN[0][ii];
            gr[r][1] = gr[r][1] + c * g.dN[1][ii];
            gr[r][2] = gr[r][2] + c * g.dN[2][ii];
        }
    }
    const double e0 = gr[0][0], e1 = gr[1][1], e2 = gr[2][2];           // :682-687
    const double e3 = 0.5 * (gr[0][1] + gr[1][0]), e4 = 0.5 * (gr[1][2] + gr[2][1]), e5 = 0.5 * (gr[0][2] + gr[2][0]);
    // MATMUL(Dmat, strain), inner index ascending; zero entries of Dmat add +0.0 * e (exact for finite strains)
    const double s0 = ((0.0 + Dd * e0) + Do * e1) + Do * e2;
    const double s1 = ((0.0 + Do * e0) + Dd * e1) + Do * e2;
    const double s2 = ((0.0 + Do * e0) + Do * e1) + Dd * e2;
    const double s3 = 0.0 + Ds * e3, s4 = 0.0 + Ds * e4, s5 = 0.0 + Ds * e5;
    const double c1 = dvol * pick(g.dN[0], li), c2 = dvol * pick(g.dN[1], li), c3 = dvol * pick(g.dN[2], li);
    const double c4 = dvol * pick(g.N, li);                             // :709-712
    F[0] = (0.0 + c4 * p.b[0]) - ((c1 * s0 + c2 * s3) + c3 * s5);       // :714-720
    F[1] = (0.0 + c4 * p.b[1]) - ((c1 * s3 + c2 * s1) + c3 * s4);
    F[2] = (0.0 + c4 * p.b[2]) - ((c1 * s5 + c2 * s4) + c3 * s2);
}

__device__ __forceinline__ double mass_node_tet(const double x[4], const double y[4], const double z[4], const ExplicitParams &p,
                                                int li, bool &neg)
{
    Geom<4, 3> g;
    tet_geom(x, y, z, g);
    neg = g.Jac < 0.0;
    const double dvol = sixth_f() * (g.Jac * p.dens);                   // :446
    const double b4 = dvol * pick(g.N, li);
    double fact = 0.0;
    fact = fact + b4 * g.N[0];
    fact = fact + b4 * g.N[1];
    fact = fact + b4 * g.N[2];
    fact = fact + b4 * g.N[3];
    return fact;
}

__device__ __forceinline__ ExplicitParams load_params(const double *prm)
{
    ExplicitParams p;
    p.E = prm[0]; p.nu = prm[1]; p.dens = prm[2]; p.b[0] = prm[3]; p.b[1] = prm[4]; p.b[2] = prm[5];
    return p;
}

// ---- gathers --------------------------------------------------------------------------------------------------------

template <int KIND> struct ExTraits;
template <> struct ExTraits<ELASTICITY_TRIA> { static constexpr int NPE = 3, NDOF = 2, NDIM = 2; };
template <> struct ExTraits<ELASTICITY_TETRA> { static constexpr int NPE = 4, NDOF = 3, NDIM = 3; };

template <int KIND>
__device__ __forceinline__ void load_elem_coords(const int4 c, const double *__restrict__ xyz, double x[4], double y[4], double z[4])
{
    const int nd[4] = {c.x, c.y, c.z, c.w};
    if (ExTraits<KIND>::NDIM == 3) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const double4 v = reinterpret_cast<const double4 *>(xyz)[nd[i]];
            x[i] = v.x; y[i] = v.y; z[i] = v.z;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double2 v = reinterpret_cast<const double2 *>(xyz)[nd[i]];
            x[i] = v.x; y[i] = v.y; z[i] = 0.0;
        }
        x[3] = y[3] = z[3] = 0.0;
    }
}

// globalM(node dof) = sum over the node's elements (ascending id) of Mlocal   (triaelasticityexplicit.F:881-921)
template <int KIND>
__global__ void __launch_bounds__(128)
ex_mass_kernel(int nNode, const int *__restrict__ inc_ptr, const int *__restrict__ inc, const int *__restrict__ conn4,
               const double *__restrict__ xyz, const double *__restrict__ prm, double *__restrict__ M, int *__restrict__ negcount)
{
    constexpr int NPE = ExTraits<KIND>::NPE, NDOF = ExTraits<KIND>::NDOF;
    const ExplicitParams p = load_params(prm);
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < nNode; n += gridDim.x * blockDim.x) {
        double acc = 0.0;
        int nneg = 0;
        for (int m = inc_ptr[n]; m < inc_ptr[n + 1]; m++) {
            const int code = inc[m], e = code / NPE, li = code - e * NPE;
            const int4 c = reinterpret_cast<const int4 *>(conn4)[e];
            double x[4], y[4], z[4];
            load_elem_coords<KIND>(c, xyz, x, y, z);
            bool neg;
            const double ml = KIND == ELASTICITY_TRIA ? mass_node_tria(x, y, p, li, neg) : mass_node_tet(x, y, z, p, li, neg);
            acc = acc + ml;
            nneg += neg && li == 0;             // count every bad element once
        }
#pragma unroll
        for (int d = 0; d < NDOF; d++) M[(size_t)n * NDOF + d] = acc;
        if (nneg) atomicAdd(negcount, nneg);
    }
}

// one central-difference step for every node   (triaelasticityexplicit.F:994-1085 + the buffer rotation :1118-1121)
template <int KIND>
__global__ void __launch_bounds__(128)
ex_step_kernel(int nNode, const int *__restrict__ inc_ptr, const int *__restrict__ inc, const int *__restrict__ conn4,
               const double *__restrict__ xyz, const double *__restrict__ prm, const double *__restrict__ M,
               const unsigned char *__restrict__ free_mask, const double *__restrict__ d1, const double *__restrict__ d2,
               double *__restrict__ d0, double *__restrict__ velo, double *__restrict__ acce, double dt, int *__restrict__ negcount)
{
    constexpr int NPE = ExTraits<KIND>::NPE, NDOF = ExTraits<KIND>::NDOF;
    const ExplicitParams p = load_params(prm);
    const double DTT = dt * dt, IDTT = 1.0 / DTT;                        // :961-962
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < nNode; n += gridDim.x * blockDim.x) {
        double acc[NDOF];
#pragma unroll
        for (int d = 0; d < NDOF; d++) acc[d] = 0.0;
        int nneg = 0;
        for (int m = inc_ptr[n]; m < inc_ptr[n + 1]; m++) {
            const int code = inc[m], e = code / NPE, li = code - e * NPE;
            const int4 c = reinterpret_cast<const int4 *>(conn4)[e];
            const int nd[4] = {c.x, c.y, c.z, c.w};
            double x[4], y[4], z[4], u[NPE * NDOF], F[NDOF];
            load_elem_coords<KIND>(c, xyz, x, y, z);
#pragma unroll
            for (int i = 0; i < NPE; i++)
#pragma unroll
                for (int d = 0; d < NDOF; d++) u[i * NDOF + d] = d1[(size_t)nd[i] * NDOF + d];
            bool neg;
            if (KIND == ELASTICITY_TRIA) residual_node_tria(x, y, u, p, li, F, neg);
            else residual_node_tet(x, y, z, u, p, li, F, neg);
#pragma unroll
            for (int d = 0; d < NDOF; d++) acc[d] = acc[d] + F[d];       // rhsVec(k) = rhsVec(k) + Flocal(ii), element order
            nneg += neg && li == 0;
        }
#pragma unroll
        for (int d = 0; d < NDOF; d++) {
            const size_t jj = (size_t)n * NDOF + d;
            const double up = d1[jj], up2 = d2[jj];
            double un = up;                                              // Dirichlet dofs are never touched by the loop :1072
            if (free_mask[jj]) {
                const double mj = M[jj];
                const double rhs = acc[d] + IDTT * mj * (2.0 * up - up2);    // :1075
                un = (DTT * rhs) / mj;                                   // :1077
            }
            d0[jj] = un;
            velo[jj] = (un - up2) / (2.0 * dt);                          // :1084
            acce[jj] = (un - 2.0 * up + up2) / DTT;                      // :1085
        }
        if (nneg) atomicAdd(negcount, nneg);
    }
}

}  // namespace pfem
