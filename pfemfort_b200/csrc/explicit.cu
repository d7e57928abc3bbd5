// explicit.cu -- matrix-free explicit dynamics: element residual, lumped mass and the
// central-difference time loop of the *elasticityexplicit drivers, on the GPU.
//
// Replaces, for P1 triangles (plane strain) and P1 tetrahedra:
//   ResidualElasticityLinearTria   elementutilitieselasticity2D.F:158-275
//   MassMatrixLinearTria           elementutilitieselasticity2D.F:283-362
//   ResidualElasticityLinearTetra  elementutilitieselasticity3D.F:575-723   (documented intent: ETYPE 4, one Gauss point)
//   MassMatrixLinearTetra          elementutilitieselasticity3D.F:401-482   (same)
//   the lumped-mass loop           triaelasticityexplicit.F:881-921
//   one time step                  triaelasticityexplicit.F:972-1121  (rhs = sum of element residuals; free dofs:
//                                  rhs += M/dt^2 (2 u_n - u_{n-1}), u_{n+1} = dt^2 rhs / M; velocity, acceleration)
//
// There is no matrix and no solver: the reference scatters Flocal into a plain array element by element.  Here one
// thread owns one NODE and gathers the contributions of its incident elements in ascending element id -- exactly the
// order in which the sequential loop adds them -- recomputing the element's stress for each incidence (the same
// deterministic, atomic-free gather as the implicit value pass), and the dof update, velocity and acceleration of the
// node are fused into the same kernel: ONE launch per time step, three rotating displacement buffers.
// Compiled with -fmad=false and written in the reference's evaluation order: results are bit-identical to the
// sequential CPU evaluation (tests/test_gpu_explicit.py).
#include <cub/cub.cuh>

#include "explicit.cuh"
#include "internal.cuh"

struct pfem_explicit {
    int device = 0, kind = -1, npe = 0, ndof = 0, ndim = 0, nElem = 0, nNode = 0, size_global = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    pfem::DevBuf<int> conn4, inc_ptr, inc, neg;
    pfem::DevBuf<double> xyz, M, d[3], velo, acce, prm;
    pfem::DevBuf<unsigned char> free_mask;
    int cur = 1;                         // d[cur] = disp (= dispPrev), d[(cur+1)%3] = dispPrev2, d[(cur+2)%3] = next
    bool have_mesh = false, have_mass = false, have_free = false;
    long long launches = 0, steps = 0;
    double t_advance = 0.0;
};

namespace pfem {

// ---- mesh set-up ----------------------------------------------------------------------------------------------------

__global__ void ex_pack_kernel(int nElem, int npe, const int *__restrict__ conn_soa, int *__restrict__ conn4, int *__restrict__ keys,
                               int *__restrict__ vals, int nNode, int *__restrict__ bad)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)nElem * npe; t += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(t / npe), i = (int)(t - (long long)e * npe);
        const int n = conn_soa[(size_t)i * nElem + e] - 1;
        if (n < 0 || n >= nNode) { atomicAdd(bad, 1); keys[t] = nNode; vals[t] = (int)t; continue; }
        conn4[(size_t)e * 4 + i] = n;
        if (npe == 3 && i == 2) conn4[(size_t)e * 4 + 3] = n;
        keys[t] = n;
        vals[t] = (int)t;                       // code = e * npe + i, ascending: a stable sort keeps ascending element id per node
    }
}

__global__ void ex_xyz_kernel(int nNode, int ndim, const double *__restrict__ coords_soa, double *__restrict__ xyz)
{
    const int stride = ndim == 3 ? 4 : 2;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < nNode; n += gridDim.x * blockDim.x)
        for (int c = 0; c < stride; c++) xyz[(size_t)n * stride + c] = c < ndim ? coords_soa[(size_t)c * nNode + n] : 0.0;
}

__global__ void ex_lower_bound_kernel(int nNode, long long total, const int *__restrict__ keys, int *__restrict__ ptr)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= nNode; r += gridDim.x * blockDim.x) {
        long long lo = 0, hi = total;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (keys[mid] < r) lo = mid + 1; else hi = mid;
        }
        ptr[r] = (int)lo;
    }
}

__global__ void ex_free_kernel(int size_global, long long nd, const int *__restrict__ slots, unsigned char *__restrict__ mask, int *__restrict__ bad)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < size_global; i += gridDim.x * blockDim.x) {
        const long long s = (long long)slots[i] - 1;
        if (s < 0 || s >= nd) atomicAdd(bad, 1); else mask[s] = 1;
    }
}

// single-element entry points (one thread)
__global__ void ex_single_kernel(int kind, int what, const double *__restrict__ in, double *__restrict__ out, int *__restrict__ negflag)
{
    // in: x[4] y[4] z[4] prm[6] u[12]
    const double *x = in, *y = in + 4, *z = in + 8;
    const ExplicitParams p = load_params(in + 12);
    const double *u = in + 18;
    const int npe = kind == ELASTICITY_TRIA ? 3 : 4, ndof = kind == ELASTICITY_TRIA ? 2 : 3;
    bool anyneg = false;
    for (int li = 0; li < npe; li++) {
        bool neg = false;
        if (what == 0) {
            double F[3] = {0.0, 0.0, 0.0};
            if (kind == ELASTICITY_TRIA) residual_node_tria(x, y, u, p, li, F, neg);
            else residual_node_tet(x, y, z, u, p, li, F, neg);
            for (int d = 0; d < ndof; d++) out[li * ndof + d] = neg ? 0.0 : F[d];
        } else {
            const double m = kind == ELASTICITY_TRIA ? mass_node_tria(x, y, p, li, neg) : mass_node_tet(x, y, z, p, li, neg);
            for (int d = 0; d < ndof; d++) out[li * ndof + d] = neg ? 0.0 : m;
        }
        anyneg |= neg;
    }
    *negflag = anyneg ? 1 : 0;
}

static int ex_single(int kind, int what, const double *x, const double *y, const double *z, const double *elemData, const double *u,
                     double *out)
{
    const int npe = kind == ELASTICITY_TRIA ? 3 : 4, ndof = kind == ELASTICITY_TRIA ? 2 : 3;
    if (!x || !y || (kind == ELASTICITY_TETRA && !z) || !elemData || !out || (what == 0 && !u)) { set_error("explicit element routine: NULL argument"); return PFEM_ERR_ARG; }
    double in[30] = {0};
    for (int i = 0; i < npe; i++) { in[i] = x[i]; in[4 + i] = y[i]; in[8 + i] = z ? z[i] : 0.0; }
    for (int i = 0; i < (kind == ELASTICITY_TRIA ? 5 : 6); i++) in[12 + i] = elemData[i];
    if (u) for (int i = 0; i < npe * ndof; i++) in[18 + i] = u[i];
    DevBuf<double> din, dout;
    DevBuf<int> dneg;
    PFEM_TRY(din.alloc(30)); PFEM_TRY(dout.alloc(12)); PFEM_TRY(dneg.alloc(1));
    PFEM_CUDA(cudaMemcpy(din.p, in, sizeof in, cudaMemcpyHostToDevice));
    ex_single_kernel<<<1, 1>>>(kind, what, din.p, dout.p, dneg.p);
    PFEM_CUDA(cudaGetLastError());
    int neg = 0;
    PFEM_CUDA(cudaMemcpy(out, dout.p, (size_t)npe * ndof * sizeof(double), cudaMemcpyDeviceToHost));
    PFEM_CUDA(cudaMemcpy(&neg, dneg.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (neg) { set_error("Negative Jacobian for the element in Elasticity"); return PFEM_ERR_NEG_JACOBIAN; }
    return PFEM_OK;
}

}  // namespace pfem

using namespace pfem;

#define EX_NEED(ex, name)                                                        \
    do {                                                                         \
        if (!(ex)) { set_error(name ": NULL handle"); return PFEM_ERR_ARG; }     \
        PFEM_CUDA(cudaSetDevice((ex)->device));                                  \
    } while (0)

extern "C" {

int pfem_residual_elasticity_linear_tria(const double *x, const double *y, const double *elemData, const double *timeData,
                                         const double *dispC, const double *veloC, double *Flocal)
{
    (void)timeData; (void)veloC;
    return ex_single(ELASTICITY_TRIA, 0, x, y, nullptr, elemData, dispC, Flocal);
}
int pfem_mass_matrix_linear_tria(const double *x, const double *y, const double *elemData, double *Mlocal)
{
    return ex_single(ELASTICITY_TRIA, 1, x, y, nullptr, elemData, nullptr, Mlocal);
}
int pfem_residual_elasticity_linear_tetra(const double *x, const double *y, const double *z, const double *elemData,
                                          const double *timeData, const double *valC, const double *valDotC, double *Flocal)
{
    (void)timeData; (void)valDotC;
    return ex_single(ELASTICITY_TETRA, 0, x, y, z, elemData, valC, Flocal);
}
int pfem_mass_matrix_linear_tetra(const double *x, const double *y, const double *z, const double *elemData, double *Mlocal)
{
    return ex_single(ELASTICITY_TETRA, 1, x, y, z, elemData, nullptr, Mlocal);
}

int pfem_explicit_create(pfem_explicit_t **out, int device)
{
    if (!out) { set_error("pfem_explicit_create: NULL output"); return PFEM_ERR_ARG; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        set_error("pfem_explicit_create: no CUDA device (there is no CPU fallback)");
        return PFEM_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("pfem_explicit_create: device %d out of range", device); return PFEM_ERR_ARG; }
    pfem_explicit *ex = new pfem_explicit;
    ex->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ex->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ex->ev0) != cudaSuccess || cudaEventCreate(&ex->ev1) != cudaSuccess) {
        set_error("pfem_explicit_create: %s", cudaGetErrorString(cudaGetLastError()));
        pfem_explicit_free(ex);
        return PFEM_ERR_CUDA;
    }
    *out = ex;
    return PFEM_OK;
}

int pfem_explicit_free(pfem_explicit_t *ex)
{
    if (!ex) return PFEM_OK;
    cudaSetDevice(ex->device);
    if (ex->ev0) cudaEventDestroy(ex->ev0);
    if (ex->ev1) cudaEventDestroy(ex->ev1);
    if (ex->stream) cudaStreamDestroy(ex->stream);
    delete ex;
    return PFEM_OK;
}

int pfem_explicit_set_mesh(pfem_explicit_t *ex, int kind, int nElem, const int *conn, int nNode, const double *coords)
{
    EX_NEED(ex, "pfem_explicit_set_mesh");
    if (kind != PFEM_ELASTICITY_TRIA && kind != PFEM_ELASTICITY_TETRA) { set_error("pfem_explicit_set_mesh: kind must be PFEM_ELASTICITY_TRIA or PFEM_ELASTICITY_TETRA"); return PFEM_ERR_ARG; }
    if (nElem <= 0 || nNode <= 0 || !conn || !coords) { set_error("pfem_explicit_set_mesh: bad argument"); return PFEM_ERR_ARG; }
    const int npe = kind == PFEM_ELASTICITY_TRIA ? 3 : 4, ndof = npe - 1, ndim = ndof;
    if ((long long)nElem * npe >= (1LL << 31)) { set_error("nElem*npElem exceeds 2^31"); return PFEM_ERR_SIZE; }
    cudaStream_t s = ex->stream;
    ex->kind = kind; ex->npe = npe; ex->ndof = ndof; ex->ndim = ndim; ex->nElem = nElem; ex->nNode = nNode;
    ex->have_mesh = ex->have_mass = ex->have_free = false;
    const long long total = (long long)nElem * npe;
    const size_t nd = (size_t)nNode * ndof;
    DevBuf<int> tconn, k_in, k_out, v_in, v_out, bad;
    DevBuf<double> tco;
    PFEM_TRY(tconn.alloc((size_t)total)); PFEM_TRY(k_in.alloc((size_t)total)); PFEM_TRY(k_out.alloc((size_t)total));
    PFEM_TRY(v_in.alloc((size_t)total)); PFEM_TRY(v_out.alloc((size_t)total)); PFEM_TRY(bad.alloc(1));
    PFEM_TRY(tco.alloc((size_t)nNode * ndim));
    PFEM_TRY(ex->conn4.alloc((size_t)nElem * 4));
    PFEM_TRY(ex->xyz.alloc((size_t)nNode * (ndim == 3 ? 4 : 2)));
    PFEM_TRY(ex->inc_ptr.alloc((size_t)nNode + 1));
    PFEM_TRY(ex->neg.alloc(1));
    PFEM_CUDA(cudaMemcpyAsync(tconn.p, conn, (size_t)total * sizeof(int), cudaMemcpyHostToDevice, s));
    PFEM_CUDA(cudaMemcpyAsync(tco.p, coords, (size_t)nNode * ndim * sizeof(double), cudaMemcpyHostToDevice, s));
    PFEM_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
    ex_pack_kernel<<<148 * 8, 256, 0, s>>>(nElem, npe, tconn.p, ex->conn4.p, k_in.p, v_in.p, nNode, bad.p);
    ex_xyz_kernel<<<148 * 4, 256, 0, s>>>(nNode, ndim, tco.p, ex->xyz.p);
    int bits = 1;
    while ((1LL << bits) <= nNode) bits++;
    size_t tmp_bytes = 0;
    PFEM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, total, 0, bits, s));
    DevBuf<char> tmp;
    PFEM_TRY(tmp.alloc(tmp_bytes));
    PFEM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, total, 0, bits, s));
    ex_lower_bound_kernel<<<148 * 4, 256, 0, s>>>(nNode, total, k_out.p, ex->inc_ptr.p);
    ex->launches += 4;
    int nbad = 0, ninc = 0;
    PFEM_CUDA(cudaMemcpyAsync(&nbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaMemcpyAsync(&ninc, ex->inc_ptr.p + nNode, sizeof(int), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    if (nbad) { set_error("pfem_explicit_set_mesh: %d connectivity entries outside 1..nNode", nbad); return PFEM_ERR_ARG; }
    PFEM_TRY(ex->inc.alloc((size_t)ninc));
    PFEM_CUDA(cudaMemcpyAsync(ex->inc.p, v_out.p, (size_t)ninc * sizeof(int), cudaMemcpyDeviceToDevice, s));
    PFEM_TRY(ex->M.alloc(nd)); PFEM_TRY(ex->velo.alloc(nd)); PFEM_TRY(ex->acce.alloc(nd)); PFEM_TRY(ex->free_mask.alloc(nd));
    PFEM_TRY(ex->prm.alloc(8));
    for (int k = 0; k < 3; k++) {
        PFEM_TRY(ex->d[k].alloc(nd));
        PFEM_CUDA(cudaMemsetAsync(ex->d[k].p, 0, nd * sizeof(double), s));       // disp = dispPrev = dispPrev2 = 0 (:952)
    }
    PFEM_CUDA(cudaMemsetAsync(ex->velo.p, 0, nd * sizeof(double), s));
    PFEM_CUDA(cudaMemsetAsync(ex->acce.p, 0, nd * sizeof(double), s));
    PFEM_CUDA(cudaMemsetAsync(ex->M.p, 0, nd * sizeof(double), s));
    PFEM_CUDA(cudaMemsetAsync(ex->free_mask.p, 0, nd, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    ex->cur = 1; ex->steps = 0;
    ex->have_mesh = true;
    return PFEM_OK;
}

int pfem_explicit_set_free_dofs(pfem_explicit_t *ex, int size_global, const int *assyForSoln)
{
    EX_NEED(ex, "pfem_explicit_set_free_dofs");
    if (!ex->have_mesh) { set_error("pfem_explicit_set_free_dofs: call pfem_explicit_set_mesh first"); return PFEM_ERR_STATE; }
    const long long nd = (long long)ex->nNode * ex->ndof;
    if (size_global < 0 || size_global > nd || (size_global > 0 && !assyForSoln)) { set_error("pfem_explicit_set_free_dofs: bad argument"); return PFEM_ERR_ARG; }
    cudaStream_t s = ex->stream;
    DevBuf<int> slots, bad;
    PFEM_TRY(slots.alloc((size_t)size_global + 1)); PFEM_TRY(bad.alloc(1));
    PFEM_CUDA(cudaMemcpyAsync(slots.p, assyForSoln, (size_t)size_global * sizeof(int), cudaMemcpyHostToDevice, s));
    PFEM_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
    PFEM_CUDA(cudaMemsetAsync(ex->free_mask.p, 0, (size_t)nd, s));
    ex_free_kernel<<<148 * 4, 256, 0, s>>>(size_global, nd, slots.p, ex->free_mask.p, bad.p);
    ex->launches++;
    int nbad = 0;
    PFEM_CUDA(cudaMemcpyAsync(&nbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    if (nbad) { set_error("pfem_explicit_set_free_dofs: %d slots outside 1..nNode*ndof", nbad); return PFEM_ERR_NUMBERING; }
    ex->size_global = size_global;
    ex->have_free = true;
    return PFEM_OK;
}

static int ex_upload_params(pfem_explicit *ex, const double *elemData)
{
    double prm[8] = {0};
    for (int i = 0; i < (ex->kind == PFEM_ELASTICITY_TRIA ? 5 : 6); i++) prm[i] = elemData[i];
    PFEM_CUDA(cudaMemcpyAsync(ex->prm.p, prm, sizeof prm, cudaMemcpyHostToDevice, ex->stream));
    PFEM_CUDA(cudaMemsetAsync(ex->neg.p, 0, sizeof(int), ex->stream));
    return PFEM_OK;
}

static int ex_check_neg(pfem_explicit *ex, const char *who)
{
    int neg = 0;
    PFEM_CUDA(cudaMemcpyAsync(&neg, ex->neg.p, sizeof(int), cudaMemcpyDeviceToHost, ex->stream));
    PFEM_CUDA(cudaStreamSynchronize(ex->stream));
    PFEM_CUDA(cudaGetLastError());
    if (neg) { set_error("%s: Negative Jacobian in %d element visit(s)", who, neg); return PFEM_ERR_NEG_JACOBIAN; }
    return PFEM_OK;
}

int pfem_explicit_lumped_mass(pfem_explicit_t *ex, const double *elemData)
{
    EX_NEED(ex, "pfem_explicit_lumped_mass");
    if (!ex->have_mesh || !elemData) { set_error("pfem_explicit_lumped_mass: no mesh / NULL elemData"); return PFEM_ERR_STATE; }
    PFEM_TRY(ex_upload_params(ex, elemData));
    const int grid = (ex->nNode + 127) / 128;
    if (ex->kind == PFEM_ELASTICITY_TRIA)
        ex_mass_kernel<ELASTICITY_TRIA><<<grid, 128, 0, ex->stream>>>(ex->nNode, ex->inc_ptr.p, ex->inc.p, ex->conn4.p, ex->xyz.p, ex->prm.p, ex->M.p, ex->neg.p);
    else
        ex_mass_kernel<ELASTICITY_TETRA><<<grid, 128, 0, ex->stream>>>(ex->nNode, ex->inc_ptr.p, ex->inc.p, ex->conn4.p, ex->xyz.p, ex->prm.p, ex->M.p, ex->neg.p);
    ex->launches++;
    PFEM_TRY(ex_check_neg(ex, "pfem_explicit_lumped_mass"));
    ex->have_mass = true;
    return PFEM_OK;
}

int pfem_explicit_advance(pfem_explicit_t *ex, int nsteps, double dt, const double *elemData, const double *timeData)
{
    (void)timeData;                        // af and timefact are read by the reference routines but never used
    EX_NEED(ex, "pfem_explicit_advance");
    if (!ex->have_mesh || !ex->have_mass || !ex->have_free) { set_error("pfem_explicit_advance: set_mesh, set_free_dofs and lumped_mass come first"); return PFEM_ERR_STATE; }
    if (nsteps < 0 || !(dt > 0.0) || !elemData) { set_error("pfem_explicit_advance: bad argument"); return PFEM_ERR_ARG; }
    PFEM_TRY(ex_upload_params(ex, elemData));
    cudaStream_t s = ex->stream;
    const int grid = (ex->nNode + 127) / 128;
    PFEM_CUDA(cudaEventRecord(ex->ev0, s));
    for (int k = 0; k < nsteps; k++) {
        const int c = ex->cur, p2 = (c + 1) % 3, nx = (c + 2) % 3;
        if (ex->kind == PFEM_ELASTICITY_TRIA)
            ex_step_kernel<ELASTICITY_TRIA><<<grid, 128, 0, s>>>(ex->nNode, ex->inc_ptr.p, ex->inc.p, ex->conn4.p, ex->xyz.p, ex->prm.p, ex->M.p,
                                                                  ex->free_mask.p, ex->d[c].p, ex->d[p2].p, ex->d[nx].p, ex->velo.p, ex->acce.p, dt, ex->neg.p);
        else
            ex_step_kernel<ELASTICITY_TETRA><<<grid, 128, 0, s>>>(ex->nNode, ex->inc_ptr.p, ex->inc.p, ex->conn4.p, ex->xyz.p, ex->prm.p, ex->M.p,
                                                                   ex->free_mask.p, ex->d[c].p, ex->d[p2].p, ex->d[nx].p, ex->velo.p, ex->acce.p, dt, ex->neg.p);
        // rotation :1118-1121: dispPrev2 = dispPrev (= old disp), dispPrev = disp (= new)
        ex->cur = nx;                      // new disp; its "prev2" slot (cur+1)%3 is the old disp: exactly dispPrev2 = dispPrev
        ex->launches++;
        ex->steps++;
    }
    PFEM_CUDA(cudaEventRecord(ex->ev1, s));
    PFEM_TRY(ex_check_neg(ex, "pfem_explicit_advance"));
    float ms = 0.f;
    PFEM_CUDA(cudaEventElapsedTime(&ms, ex->ev0, ex->ev1));
    ex->t_advance = ms * 1e-3;
    return PFEM_OK;
}

int pfem_explicit_get_state(pfem_explicit_t *ex, double *disp, double *dispPrev2, double *velo, double *acce, double *mass)
{
    EX_NEED(ex, "pfem_explicit_get_state");
    if (!ex->have_mesh) { set_error("pfem_explicit_get_state: no mesh"); return PFEM_ERR_STATE; }
    const size_t bytes = (size_t)ex->nNode * ex->ndof * sizeof(double);
    cudaStream_t s = ex->stream;
    if (disp) PFEM_CUDA(cudaMemcpyAsync(disp, ex->d[ex->cur].p, bytes, cudaMemcpyDeviceToHost, s));
    if (dispPrev2) PFEM_CUDA(cudaMemcpyAsync(dispPrev2, ex->d[(ex->cur + 1) % 3].p, bytes, cudaMemcpyDeviceToHost, s));
    if (velo) PFEM_CUDA(cudaMemcpyAsync(velo, ex->velo.p, bytes, cudaMemcpyDeviceToHost, s));
    if (acce) PFEM_CUDA(cudaMemcpyAsync(acce, ex->acce.p, bytes, cudaMemcpyDeviceToHost, s));
    if (mass) PFEM_CUDA(cudaMemcpyAsync(mass, ex->M.p, bytes, cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    return PFEM_OK;
}

int pfem_explicit_set_state(pfem_explicit_t *ex, const double *disp, const double *dispPrev2)
{
    EX_NEED(ex, "pfem_explicit_set_state");
    if (!ex->have_mesh || !disp || !dispPrev2) { set_error("pfem_explicit_set_state: bad argument"); return PFEM_ERR_ARG; }
    const size_t bytes = (size_t)ex->nNode * ex->ndof * sizeof(double);
    cudaStream_t s = ex->stream;
    PFEM_CUDA(cudaMemcpyAsync(ex->d[ex->cur].p, disp, bytes, cudaMemcpyHostToDevice, s));
    PFEM_CUDA(cudaMemcpyAsync(ex->d[(ex->cur + 1) % 3].p, dispPrev2, bytes, cudaMemcpyHostToDevice, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    return PFEM_OK;
}

int pfem_explicit_get_info(pfem_explicit_t *ex, long long *steps, long long *launches, double *t_advance)
{
    if (!ex) { set_error("pfem_explicit_get_info: NULL handle"); return PFEM_ERR_ARG; }
    if (steps) *steps = ex->steps;
    if (launches) *launches = ex->launches;
    if (t_advance) *t_advance = ex->t_advance;
    return PFEM_OK;
}

}  // extern "C"
