// assembly_tiled.cu -- set-up and launch of the tiled (compute-once) value pass (kernels: assembly_tiled.cuh,
// tile construction: tiles.hpp).  Opt-in with PFEM_ASM=tiled (staged columns + row gather) or PFEM_ASM=tiled2
// (sorted scatter + run sums) for the one-dof-per-node kinds (Poisson tria/tet);
// everything else, and any mesh the tile builder declines, stays on the row-gather kernels of assembly.cu.
//
// Compiled with -fmad=false like assembly.cu: the element arithmetic must not be contracted into FMAs.
#include <cstdlib>

#include "assembly_tiled.cuh"
#include "internal.cuh"

namespace pfem {

template <typename T> static int fetch(std::vector<T> &dst, const T *src, size_t n, cudaStream_t s)
{
    dst.resize(n ? n : 1);
    if (n) PFEM_CUDA(cudaMemcpyAsync(dst.data(), src, n * sizeof(T), cudaMemcpyDeviceToHost, s));
    return PFEM_OK;
}

template <typename T> static int push(DevBuf<T> &dst, const std::vector<T> &src, cudaStream_t s)
{
    PFEM_TRY(dst.alloc(src.size()));
    if (!src.empty()) PFEM_CUDA(cudaMemcpyAsync(dst.p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return PFEM_OK;
}

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v && v[0] ? atoi(v) : dflt;
}

// Build the tiles of the current pattern (once per pattern pass).  The first version runs the construction on the
// host from copies of the pattern-pass arrays: it is set-up work (the reference's own pattern pass is host code).
// On return h->asm_tiled tells whether the tiled kernel can be used.
int build_tiles_device(pfem_solver *h, int mode)
{
    h->tiles_ready = true;
    h->asm_tiled = false;
    h->tile_mode = mode;
    if (h->ndof != 1 || !h->asm_sell || h->size_local == 0) return PFEM_OK;
    StageTimer tm("assembly: build tiles (host)");
    cudaStream_t s = h->stream;
    const int nloc = h->size_local, nslices = (nloc + 31) / 32;
    std::vector<int> erec, rowptr, rinc_ptr, rinc, ainc;
    std::vector<long long> ainc_off;
    std::vector<double> xyz;
    const int stride = h->ndim == 3 ? 4 : 2;
    PFEM_TRY(fetch(ainc_off, h->ainc_off.p, (size_t)nslices + 1, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    PFEM_TRY(fetch(erec, h->erec.p, (size_t)h->nElem * h->rec_ints, s));
    PFEM_TRY(fetch(rowptr, h->rowptr.p, (size_t)nloc + 1, s));
    PFEM_TRY(fetch(rinc_ptr, h->rinc_ptr.p, (size_t)nloc + 1, s));
    PFEM_TRY(fetch(rinc, h->rinc.p, (size_t)h->ninc, s));
    PFEM_TRY(fetch(ainc, h->ainc.p, (size_t)ainc_off[nslices] * h->ainc_words, s));
    PFEM_TRY(fetch(xyz, h->xyz.p, (size_t)h->nNode * stride, s));
    PFEM_CUDA(cudaStreamSynchronize(s));

    int max_smem_optin = 0;
    PFEM_CUDA(cudaDeviceGetAttribute(&max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    TileInput in;
    in.nloc = nloc; in.row_lo = h->row_lo; in.nElem = h->nElem; in.npe = h->npe; in.nsize = h->nsize;
    in.rec_ints = h->rec_ints; in.ndim = h->ndim; in.xyz_stride = stride;
    in.erec = erec.data(); in.xyz = xyz.data(); in.rowptr = rowptr.data(); in.rinc_ptr = rinc_ptr.data();
    in.rinc = rinc.data(); in.ainc_off = ainc_off.data(); in.ainc = ainc.data(); in.ainc_words = h->ainc_words;
    // tuning hooks: CTA size (128 | 256 | 512), rows per tile, shared memory per CTA.  Defaults: 256 threads, 96 rows,
    // 110 KB => two CTAs (16 warps) per SM, so that one CTA's gather phase overlaps the other's FP64 phase.
    in.cta_threads = env_int("PFEM_TILE_THREADS", 256);
    if (in.cta_threads != 128 && in.cta_threads != 512) in.cta_threads = 256;
    in.max_rows = env_int("PFEM_TILE_ROWS", in.cta_threads == 512 ? 192 : 96);
    in.max_rows = std::max(32, std::min(in.cta_threads, in.max_rows / 32 * 32));
    const int kb = env_int("PFEM_TILE_SMEM_KB", in.cta_threads == 512 ? 220 : (mode == 2 ? 112 : 110));
    in.mode = mode;
    in.smem_budget = std::min((size_t)kb * 1024, (size_t)max_smem_optin);
    TileSet ts;
    if (build_tiles(in, ts) != 0) return PFEM_OK;              // declined: keep the row-gather kernel
    PFEM_TRY(push(h->t_desc, ts.tdesc, s));
    PFEM_TRY(push(h->t_rows, ts.trows, s));
    PFEM_TRY(push(h->t_el, ts.tel, s));
    PFEM_TRY(push(h->t_slice_off, ts.tslice_off, s));
    PFEM_TRY(push(h->t_inc, ts.tinc, s));
    PFEM_TRY(push(h->t_crec, ts.crec, s));
    PFEM_TRY(push(h->t_cnt, ts.cnt, s));
    PFEM_TRY(push(h->t_ts2, ts.ts2, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    h->ntiles = ts.ntiles;
    h->tile_threads = in.cta_threads;
    h->tile_smem = ts.max_smem;
    h->tile_elem_visits = ts.elem_visits;
    h->tile_elems_touched = ts.elems_touched;
    h->asm_tiled = true;
    if (getenv("PFEM_TRACE"))
        fprintf(stderr, "[pfem] tiles: %d tiles of <= %d rows, %d threads, %zu B smem, element visits %lld / %lld touched = %.3f\n",
                ts.ntiles, in.max_rows, in.cta_threads, ts.max_smem, ts.elem_visits, ts.elems_touched,
                ts.elems_touched ? (double)ts.elem_visits / (double)ts.elems_touched : 0.0);
    return PFEM_OK;
}

template <int KIND, int THREADS, int MINB>
static int launch_tiled(pfem_solver *h, const TiledArgs &a, bool unit)
{
    const size_t smem = h->tile_smem;
    if (h->tile_mode == 2) {
        if (unit) {
            PFEM_CUDA(cudaFuncSetAttribute(assemble_tiled2_kernel<KIND, THREADS, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            assemble_tiled2_kernel<KIND, THREADS, MINB, true><<<h->ntiles, THREADS, smem, h->stream>>>(a);
        } else {
            PFEM_CUDA(cudaFuncSetAttribute(assemble_tiled2_kernel<KIND, THREADS, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            assemble_tiled2_kernel<KIND, THREADS, MINB, false><<<h->ntiles, THREADS, smem, h->stream>>>(a);
        }
    } else if (unit) {
        PFEM_CUDA(cudaFuncSetAttribute(assemble_tiled_kernel<KIND, THREADS, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        assemble_tiled_kernel<KIND, THREADS, MINB, true><<<h->ntiles, THREADS, smem, h->stream>>>(a);
    } else {
        PFEM_CUDA(cudaFuncSetAttribute(assemble_tiled_kernel<KIND, THREADS, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        assemble_tiled_kernel<KIND, THREADS, MINB, false><<<h->ntiles, THREADS, smem, h->stream>>>(a);
    }
    h->launches++;
    PFEM_CUDA(cudaGetLastError());
    return PFEM_OK;
}

// The value pass through the tiled kernel.  Preconditions checked by the caller (assemble_values): tiles built,
// Poisson kind.  elemData/timeData are device pointers (8 doubles each).
int assemble_values_tiled(pfem_solver *h, const double *dElemData, const double *dTimeData, bool unit)
{
    TiledArgs a;
    a.tdesc = h->t_desc.p;
    a.trows = reinterpret_cast<const int4 *>(h->t_rows.p);
    a.tel = reinterpret_cast<const int2 *>(h->t_el.p);
    a.tslice_off = h->t_slice_off.p;
    a.tinc = reinterpret_cast<const int2 *>(h->t_inc.p);
    a.crec = reinterpret_cast<const int2 *>(h->t_crec.p);
    a.cnt = reinterpret_cast<const uint4 *>(h->t_cnt.p);
    a.ts2 = reinterpret_cast<const int4 *>(h->t_ts2.p);
    a.conn4 = reinterpret_cast<const int4 *>(h->conn4.p);
    a.erec = h->erec.p; a.rec_ints = h->rec_ints;
    a.xyz = h->xyz.p; a.applied = h->applied.p; a.rowptr = h->rowptr.p;
    a.val = h->val.p; a.rhs = h->rhs.p;
    a.elemData = dElemData; a.timeData = dTimeData;
    a.neg_flag = h->neg_count.p;
    a.load_val = h->values_zero ? 0 : 1; a.load_rhs = h->rhs_zero ? 0 : 1;
    if (h->ntiles == 0) return PFEM_OK;
    // CTA shapes: 128 threads x 4 CTAs/SM (small tiles), 256 x 2 (default), 512 x 1 (large tiles): 16 warps per SM each
    if (h->kind == PFEM_POISSON_TETRA)
        return h->tile_threads == 512 ? launch_tiled<POISSON_TETRA, 512, 1>(h, a, unit)
             : h->tile_threads == 256 ? launch_tiled<POISSON_TETRA, 256, 2>(h, a, unit) : launch_tiled<POISSON_TETRA, 128, 4>(h, a, unit);
    if (h->kind == PFEM_POISSON_TRIA)
        return h->tile_threads == 512 ? launch_tiled<POISSON_TRIA, 512, 1>(h, a, unit)
             : h->tile_threads == 256 ? launch_tiled<POISSON_TRIA, 256, 2>(h, a, unit) : launch_tiled<POISSON_TRIA, 128, 4>(h, a, unit);
    set_error("tiled value pass: kind %d is not a one-dof-per-node kind", h->kind);
    return PFEM_ERR_STATE;
}

}  // namespace pfem
