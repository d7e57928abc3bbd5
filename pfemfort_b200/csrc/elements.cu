// elements.cu -- batched element Ke/Fe kernels (one thread per element, SoA in/out, coalesced).
// Compiled with -fmad=false (see elements.cuh).  Mirrors the call surface of
// StiffnessResidual{Poisson,Elasticity}Linear{Tria,Tetra} (elementutilitiespoisson.F:23-193,
// elementutilitieselasticity2D.F:23-153, elementutilitieselasticity3D.F:248-393).
#include "elements.cuh"
#include "internal.cuh"

namespace pfem {

template <int KIND>
__global__ void __launch_bounds__(128)
element_batch_kernel(int n, const double *__restrict__ x, const double *__restrict__ y,
                     const double *__restrict__ z, const double *__restrict__ elemData,
                     const double *__restrict__ timeData, const double *__restrict__ valC,
                     double *__restrict__ K, double *__restrict__ F, int *__restrict__ jac_neg)
{
    using T = ElemTraits<KIND>;
    constexpr int NPE = T::NPE, NDOF = T::NDOF, NDIM = T::NDIM, NSIZE = NPE * NDOF;
    Params<KIND> prm;
    prm.init(elemData, timeData);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        double xn[NPE], yn[NPE], zn[NPE];
#pragma unroll
        for (int i = 0; i < NPE; i++) {
            xn[i] = x[(size_t)i * n + e];
            yn[i] = y[(size_t)i * n + e];
            zn[i] = NDIM == 3 ? z[(size_t)i * n + e] : 0.0;
        }
        ElemOp<KIND> op;
        op.load_geom(xn, yn, zn);
        const bool neg = op.g.Jac < 0.0;
        if (jac_neg) jac_neg[e] = neg ? 1 : 0;
        if (neg) {   // the reference STOPs; leave Klocal = Flocal = 0 (their state at the STOP)
#pragma unroll 1
            for (int k = 0; k < NSIZE * NSIZE; k++) K[(size_t)k * n + e] = 0.0;
#pragma unroll 1
            for (int k = 0; k < NSIZE; k++) F[(size_t)k * n + e] = 0.0;
            continue;
        }
        op.set_dvol(prm);
        // du = sum valC * grad N (poisson.F:77-81,165-170); unused by the elasticity routines
        double du[3] = {0.0, 0.0, 0.0};
        if (NDOF == 1 && valC) {
#pragma unroll
            for (int i = 0; i < NPE; i++) {
                const double v = valC[(size_t)i * n + e];
#pragma unroll
                for (int c = 0; c < NDIM; c++) du[c] = du[c] + v * op.g.dN[c][i];
            }
        }
#pragma unroll
        for (int b = 0; b < NSIZE; b++) {
            op.col_setup(prm, b);
#pragma unroll
            for (int a = 0; a < NSIZE; a++) K[(size_t)(a + NSIZE * b) * n + e] = op.K(prm, a);
        }
#pragma unroll
        for (int a = 0; a < NSIZE; a++) F[(size_t)a * n + e] = op.F(prm, a, du);
    }
}

template <int KIND>
static int run_batch(int n, const double *x, const double *y, const double *z, const double *elemData,
                     const double *timeData, const double *valC, double *K, double *F, int *jac_neg)
{
    using T = ElemTraits<KIND>;
    constexpr int NPE = T::NPE, NSIZE = T::NPE * T::NDOF, NDIM = T::NDIM;
    DevBuf<double> dx, dy, dz, dK, dF, dvalC, dED, dTD;
    DevBuf<int> dneg;
    const size_t nn = (size_t)n;
    PFEM_TRY(dx.alloc(nn * NPE));
    PFEM_TRY(dy.alloc(nn * NPE));
    PFEM_TRY(dz.alloc(nn * NPE));
    PFEM_TRY(dK.alloc(nn * NSIZE * NSIZE));
    PFEM_TRY(dF.alloc(nn * NSIZE));
    PFEM_TRY(dED.alloc(8));
    PFEM_TRY(dTD.alloc(8));
    PFEM_TRY(dneg.alloc(nn));
    PFEM_CUDA(cudaMemcpy(dx.p, x, nn * NPE * sizeof(double), cudaMemcpyHostToDevice));
    PFEM_CUDA(cudaMemcpy(dy.p, y, nn * NPE * sizeof(double), cudaMemcpyHostToDevice));
    if (NDIM == 3) PFEM_CUDA(cudaMemcpy(dz.p, z, nn * NPE * sizeof(double), cudaMemcpyHostToDevice));
    // elemData: up to 6 entries are read (E, nu, thick, bx, by, bz); timeData: entries 2 and 3 (af, timefact)
    double ed[8] = {0}, td[8] = {0};
    const int ned = KIND == POISSON_TRIA ? 2 : KIND == POISSON_TETRA ? 3 : KIND == ELASTICITY_TRIA ? 5 : 6;
    for (int i = 0; i < ned; i++) ed[i] = elemData[i];
    td[1] = timeData[1];
    PFEM_CUDA(cudaMemcpy(dED.p, ed, sizeof ed, cudaMemcpyHostToDevice));
    PFEM_CUDA(cudaMemcpy(dTD.p, td, sizeof td, cudaMemcpyHostToDevice));
    if (valC && T::NDOF == 1) {
        PFEM_TRY(dvalC.alloc(nn * NSIZE));
        PFEM_CUDA(cudaMemcpy(dvalC.p, valC, nn * NSIZE * sizeof(double), cudaMemcpyHostToDevice));
    }
    const int threads = 128;
    int blocks = ceil_div(n, threads);
    if (blocks > 148 * 16) blocks = 148 * 16;
    element_batch_kernel<KIND><<<blocks, threads>>>(n, dx.p, dy.p, dz.p, dED.p, dTD.p,
                                                    (valC && T::NDOF == 1) ? dvalC.p : nullptr, dK.p, dF.p, dneg.p);
    PFEM_CUDA(cudaGetLastError());
    PFEM_CUDA(cudaMemcpy(K, dK.p, nn * NSIZE * NSIZE * sizeof(double), cudaMemcpyDeviceToHost));
    PFEM_CUDA(cudaMemcpy(F, dF.p, nn * NSIZE * sizeof(double), cudaMemcpyDeviceToHost));
    if (jac_neg) PFEM_CUDA(cudaMemcpy(jac_neg, dneg.p, nn * sizeof(int), cudaMemcpyDeviceToHost));
    return PFEM_OK;
}

int element_ke_batch(int kind, int n, const double *x, const double *y, const double *z, const double *elemData,
                     const double *timeData, const double *valC, double *K, double *F, int *jac_neg)
{
    if (n <= 0 || !x || !y || !elemData || !timeData || !K || !F) {
        set_error("pfem_element_ke_batch: bad argument");
        return PFEM_ERR_ARG;
    }
    switch (kind) {
    case POISSON_TRIA: return run_batch<POISSON_TRIA>(n, x, y, z, elemData, timeData, valC, K, F, jac_neg);
    case POISSON_TETRA: return z ? run_batch<POISSON_TETRA>(n, x, y, z, elemData, timeData, valC, K, F, jac_neg) : PFEM_ERR_ARG;
    case ELASTICITY_TRIA: return run_batch<ELASTICITY_TRIA>(n, x, y, z, elemData, timeData, valC, K, F, jac_neg);
    case ELASTICITY_TETRA: return z ? run_batch<ELASTICITY_TETRA>(n, x, y, z, elemData, timeData, valC, K, F, jac_neg) : PFEM_ERR_ARG;
    }
    set_error("pfem_element_ke_batch: unknown kind %d", kind);
    return PFEM_ERR_ARG;
}

}  // namespace pfem
