// assembly_fast.cu -- the streamed row-gather value pass (assembly_rows.cuh) instantiated with the FMA element operators
// (elements_fast.cuh).  This translation unit is compiled WITH FMA contraction (build.py: not in NO_FMA), unlike assembly.cu:
// same kernel, same sequential summation order per matrix entry (tetrapoissonparallelimpl1.F:828-884), faster arithmetic;
// the contract is 1e-12 relative against the reference evaluation order instead of bit-identity.
#include "assembly_rows.cuh"
#include "elements_fast.cuh"
#include "internal.cuh"

namespace pfem {

template <int KIND, int R>
static int launch_fast(pfem_solver *h, const AsmArgs &args)
{
    const int blocks = ceil_div(h->size_local, R);
    if (blocks == 0) return PFEM_OK;
    constexpr bool POISSON = KIND == POISSON_TRIA || KIND == POISSON_TETRA;
    const size_t smem = h->asm_smem;
    if (POISSON && args.unit) {
        PFEM_CUDA(cudaFuncSetAttribute(assemble_sell_kernel<KIND, R, POISSON, FastOp<KIND>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        assemble_sell_kernel<KIND, R, POISSON, FastOp<KIND>><<<blocks, R, smem, h->stream>>>(args);
    } else {
        PFEM_CUDA(cudaFuncSetAttribute(assemble_sell_kernel<KIND, R, false, FastOp<KIND>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        assemble_sell_kernel<KIND, R, false, FastOp<KIND>><<<blocks, R, smem, h->stream>>>(args);
    }
    h->launches++;
    PFEM_CUDA(cudaGetLastError());
    return PFEM_OK;
}

template <int KIND>
static int dispatch_fast(pfem_solver *h, const AsmArgs &args)
{
    switch (h->asm_rows_per_cta) {
    case 256: return launch_fast<KIND, 256>(h, args);
    case 128: return launch_fast<KIND, 128>(h, args);
    case 64: return launch_fast<KIND, 64>(h, args);
    case 32: return launch_fast<KIND, 32>(h, args);
    }
    set_error("assembly: no CTA shape fits shared memory");
    return PFEM_ERR_SIZE;
}

// requires h->asm_sell (rows no wider than 254 entries: the streamed kernel's slot bytes)
int dispatch_rows_fast(pfem_solver *h, const AsmArgs &args)
{
    switch (h->kind) {
    case PFEM_POISSON_TRIA: return dispatch_fast<POISSON_TRIA>(h, args);
    case PFEM_POISSON_TETRA: return dispatch_fast<POISSON_TETRA>(h, args);
    case PFEM_ELASTICITY_TRIA: return dispatch_fast<ELASTICITY_TRIA>(h, args);
    case PFEM_ELASTICITY_TETRA: return dispatch_fast<ELASTICITY_TETRA>(h, args);
    }
    set_error("assemble: mesh kind not set");
    return PFEM_ERR_STATE;
}

}  // namespace pfem
