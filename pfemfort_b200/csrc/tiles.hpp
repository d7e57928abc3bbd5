// tiles.hpp -- host-side construction of the row tiles of the tiled (compute-once) value pass.
//
// Plain C++17, no CUDA: assembly_tiled.cu runs it at set-up time on copies of the pattern-pass arrays, and
// tests/emu runs the very same code on the CPU.  It belongs to the pattern pass of the reference
// (tetrapoissonparallelimpl1.F:791-802 + solverpetsc.F:222-246): structure only, never values.
//
// A tile is a set of matrix rows that are close together in space (consecutive in the Morton order of their
// nodes) plus the sorted list of the elements that touch them.  One CTA computes every element of its list ONCE
// (the row-gather kernel recomputes an element once per incident row) and stages, in shared memory, the columns
// Klocal(:,k) and the lifted Flocal(k) of the local dofs k whose rows belong to the tile; the tile's rows then
// gather those staged columns in ascending element id, i.e. in the reference's sequential summation order.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace pfem {

constexpr int TILE_DESC_INTS = 8;
// tdesc[t] = { row_off, nrows_padded, el_off, nel, slice0, nnz, ncols, nrows }
enum { TD_ROW_OFF = 0, TD_NROWS_PAD = 1, TD_EL_OFF = 2, TD_NEL = 3, TD_SLICE0 = 4, TD_NNZ = 5, TD_NCOLS = 6, TD_NROWS = 7 };
// scatter mode (mode 2) re-uses the last two descriptor slots: first column record of the tile, contribution buffer size
enum { TD2_CREC_OFF = 6, TD2_CBUF = 7 };
constexpr int TILE_CB_STRIDE = 33;        // doubles between consecutive positions of one row: 32 lanes + 1 (bank rotation)

struct TileInput {
    int nloc = 0, row_lo = 0, nElem = 0, npe = 0, nsize = 0, rec_ints = 0, ndim = 0, xyz_stride = 0;
    const int *erec = nullptr;            // [nElem][rec_ints]: conn[npe] (0-based), dof[nsize] (global, -1 = Dirichlet)
    const double *xyz = nullptr;          // [nNode][xyz_stride]
    const int *rowptr = nullptr;          // [nloc+1]
    const int *rinc_ptr = nullptr;        // [nloc+1]
    const int *rinc = nullptr;            // codes e*nsize+k, ascending per row
    const long long *ainc_off = nullptr;  // [ceil(nloc/32)+1] entry offsets of the row-gather streams
    const int *ainc = nullptr;            // entries of ainc_words ints: code, slot bytes
    int ainc_words = 2;
    int max_rows = 96;                    // rows per tile (<= CTA threads), multiple of 32
    size_t smem_budget = 100 * 1024;      // bytes of dynamic shared memory per CTA
    int cta_threads = 128;
    int mode = 1;                         // 1: staged columns + row gather (assemble_tiled_kernel)
                                          // 2: sorted scatter + run sums (assemble_tiled2_kernel)
};

struct TileSet {
    int ntiles = 0;
    long long nslices = 0;
    std::vector<int> tdesc;               // [ntiles][TILE_DESC_INTS]
    std::vector<int> trows;               // int4 per padded tile row: { local row or -1, accumulator offset, rowptr[row], row length }
    std::vector<int> tel;                 // int2 per tile element: { e | dbc<<31, base | mask<<24 }
    std::vector<long long> tslice_off;    // [nslices+1] entry offsets into tinc
    std::vector<int> tinc;                // int2 per entry: { staged column or -1, slot bytes }
    // mode 2: one record per staged column { row position base | p0<<16 | p1<<24, p2 | p3<<8 | pF<<16 } (p = position of the
    // contribution inside the row's run-ordered list, 255 = Dirichlet column / unused), run lengths per (row, slot) as
    // bytes in 16-byte chunks (column-major per 32-row slice), and per tile slice { buffer offset, first chunk, chunks, width }
    std::vector<int> crec;
    std::vector<unsigned char> cnt;
    std::vector<int> ts2;
    size_t max_smem = 0;                  // largest per-tile shared-memory need
    long long elem_visits = 0;            // sum of nel over tiles (rho = elem_visits / #elements with a row here)
    long long elems_touched = 0;
};

// bytes of dynamic shared memory a tile needs: staged columns (4 doubles of Klocal + 1 of Flocal), accumulators, sinks
inline size_t tile_smem_bytes(long long ncols, long long nnz, int cta_threads)
{
    const size_t nlines = (size_t)(ncols + 3) / 4;              // four staged columns per 128-byte line (+ 4 Flocal)
    return nlines * 160 + (size_t)nnz * 8 + (size_t)cta_threads * 8 + 32;
}

// mode 2: contribution buffer (doubles) + accumulators
inline size_t tile2_smem_bytes(long long cbuf_doubles, long long nnz) { return (size_t)cbuf_doubles * 8 + (size_t)nnz * 8 + 32; }

inline uint64_t morton_spread3(uint64_t v)
{
    v &= 0x1fffffULL;
    v = (v | (v << 32)) & 0x1f00000000ffffULL;
    v = (v | (v << 16)) & 0x1f0000ff0000ffULL;
    v = (v | (v << 8)) & 0x100f00f00f00f00fULL;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ULL;
    v = (v | (v << 2)) & 0x1249249249249249ULL;
    return v;
}

// Returns 0 on success, 1 when some row cannot be tiled within the budget (the caller keeps the row-gather kernel).
inline int build_tiles(const TileInput &in, TileSet &out)
{
    const int nloc = in.nloc, nsize = in.nsize, npe = in.npe;
    if (in.nsize != in.npe || nsize > 4) return 1;                 // one dof per node (the Poisson kinds)
    if (in.max_rows % 32 || in.max_rows > in.cta_threads) return 1;
    out = TileSet();
    if (nloc == 0) { out.tslice_off.assign(1, 0); return 0; }

    // 1. the node of every owned row (ndof = 1: the dof of local node k is the row of node conn[k])
    std::vector<int> row_node(nloc, -1);
    for (int r = 0; r < nloc; r++) {
        if (in.rinc_ptr[r] == in.rinc_ptr[r + 1]) continue;
        const int code = in.rinc[in.rinc_ptr[r]];
        const int e = code / nsize, k = code - e * nsize;
        row_node[r] = in.erec[(size_t)e * in.rec_ints + k];
    }
    // 2. rows in Morton order of their node coordinates (21 bits per axis over the bounding box)
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    bool first = true;
    for (int r = 0; r < nloc; r++) {
        if (row_node[r] < 0) continue;
        const double *p = in.xyz + (size_t)row_node[r] * in.xyz_stride;
        for (int d = 0; d < in.ndim; d++) {
            if (first || p[d] < lo[d]) lo[d] = p[d];
            if (first || p[d] > hi[d]) hi[d] = p[d];
        }
        first = false;
    }
    double ext = 0.0;
    for (int d = 0; d < in.ndim; d++) ext = std::max(ext, hi[d] - lo[d]);
    const double scale = ext > 0.0 ? 1048575.0 / ext : 0.0;        // one scale for all axes: isotropic cells
    std::vector<std::pair<uint64_t, int>> order(nloc);
    for (int r = 0; r < nloc; r++) {
        uint64_t key = ~0ULL;                                      // rows without elements go last
        if (row_node[r] >= 0) {
            const double *p = in.xyz + (size_t)row_node[r] * in.xyz_stride;
            key = 0;
            for (int d = 0; d < in.ndim; d++) {
                const uint64_t q = (uint64_t)((p[d] - lo[d]) * scale);
                key |= morton_spread3(q) << d;
            }
        }
        order[r] = {key, r};
    }
    std::sort(order.begin(), order.end());

    // contributions per row (mode 2): one per free column of every incidence, plus one Flocal per incidence
    std::vector<int> crow;
    if (in.mode == 2) {
        crow.assign(nloc, 0);
        for (int r = 0; r < nloc; r++) {
            const long long ao = in.ainc_off[r >> 5] + (r & 31);
            int c = 0, mm = 0;
            for (int m = in.rinc_ptr[r]; m < in.rinc_ptr[r + 1]; m++, mm++) {
                const unsigned int sw = (unsigned int)in.ainc[(size_t)(ao + (long long)mm * 32) * in.ainc_words + 1];
                for (int j = 0; j < nsize; j++) c += ((sw >> (8 * j)) & 255u) != 255u;
                c++;
            }
            if (c > 254) return 1;                                 // positions are bytes, 255 is reserved
            crow[r] = c;
        }
    }
    // 3. greedy split into tiles under the row and shared-memory budgets
    std::vector<int> tile_first;                                   // index into `order` of each tile's first row
    if (in.mode == 2) {
        int rows = 0, cur_max = 0;
        long long done = 0, nnz = 0;                               // done: widths of the tile's completed 32-row slices
        for (int q = 0; q < nloc; q++) {
            const int r = order[q].second;
            const long long z = in.rowptr[r + 1] - in.rowptr[r];
            if (z > 254 || tile2_smem_bytes((long long)crow[r] * TILE_CB_STRIDE, z) > in.smem_budget) return 1;
            long long d2 = done;
            int m2 = cur_max;
            if (rows % 32 == 0) { d2 += m2; m2 = 0; }
            m2 = std::max(m2, crow[r]);
            if (rows == 0 || rows == in.max_rows || tile2_smem_bytes((d2 + m2) * TILE_CB_STRIDE, nnz + z) > in.smem_budget) {
                tile_first.push_back(q);
                rows = 0; nnz = 0; d2 = 0; m2 = crow[r];
            }
            rows++; nnz += z; done = d2; cur_max = m2;
        }
    } else {
        int rows = 0;
        long long cols = 0, nnz = 0;
        for (int q = 0; q < nloc; q++) {
            const int r = order[q].second;
            const long long c = in.rinc_ptr[r + 1] - in.rinc_ptr[r], z = in.rowptr[r + 1] - in.rowptr[r];
            if (z > 254 || tile_smem_bytes(c, z, in.cta_threads) > in.smem_budget) return 1;
            if (rows == 0 || rows == in.max_rows || tile_smem_bytes(cols + c, nnz + z, in.cta_threads) > in.smem_budget) {
                tile_first.push_back(q);
                rows = 0; cols = 0; nnz = 0;
            }
            rows++; cols += c; nnz += z;
        }
    }
    const int ntiles = (int)tile_first.size();
    tile_first.push_back(nloc);
    out.ntiles = ntiles;
    out.tdesc.assign((size_t)ntiles * TILE_DESC_INTS, 0);

    // 4. pass 1: per-tile sizes (element count, padded rows, slice widths)
    std::vector<int> tile_nel(ntiles), tile_pad(ntiles);
    std::vector<long long> row_off(ntiles + 1, 0), slice_cnt(ntiles + 1, 0);
    for (int t = 0; t < ntiles; t++) {
        const int n = tile_first[t + 1] - tile_first[t];
        tile_pad[t] = (n + 31) / 32 * 32;
        row_off[t + 1] = row_off[t] + tile_pad[t];
        slice_cnt[t + 1] = slice_cnt[t] + tile_pad[t] / 32;
    }
    out.nslices = slice_cnt[ntiles];
    if (row_off[ntiles] >= (1LL << 31)) return 1;
    std::vector<long long> slice_sz((size_t)out.nslices + 1, 0);
    std::vector<long long> tile_cols(ntiles + 1, 0), slice_chunks((size_t)out.nslices + 1, 0);
    std::vector<int> slice_width((size_t)out.nslices + 1, 0);
#if defined(_OPENMP)
#pragma omp parallel for schedule(dynamic, 64)
#endif
    for (int t = 0; t < ntiles; t++) {
        std::vector<int> els;
        for (int q = tile_first[t]; q < tile_first[t + 1]; q++) {
            const int r = order[q].second;
            for (int m = in.rinc_ptr[r]; m < in.rinc_ptr[r + 1]; m++) els.push_back(in.rinc[m] / nsize);
        }
        std::sort(els.begin(), els.end());
        tile_nel[t] = (int)(std::unique(els.begin(), els.end()) - els.begin());
        for (int s = 0; s < tile_pad[t] / 32; s++) {
            int w = 0;
            for (int l = 0; l < 32; l++) {
                const int q = tile_first[t] + s * 32 + l;
                if (q >= tile_first[t + 1]) break;
                const int r = order[q].second;
                w = std::max(w, in.rinc_ptr[r + 1] - in.rinc_ptr[r]);
                tile_cols[t + 1] += in.rinc_ptr[r + 1] - in.rinc_ptr[r];
                if (in.mode == 2) {
                    slice_width[slice_cnt[t] + s] = std::max(slice_width[slice_cnt[t] + s], crow[r]);
                    const long long ch = (in.rowptr[r + 1] - in.rowptr[r] + 1 + 15) / 16;
                    slice_chunks[slice_cnt[t] + s] = std::max(slice_chunks[slice_cnt[t] + s], ch);
                }
            }
            slice_sz[slice_cnt[t] + s] = in.mode == 2 ? 0 : (long long)w * 32;
        }
    }
    for (int t = 0; t < ntiles; t++) tile_cols[t + 1] += tile_cols[t];
    std::vector<long long> chunk_off((size_t)out.nslices + 1, 0);
    for (long long q = 0; q < out.nslices; q++) chunk_off[q + 1] = chunk_off[q] + slice_chunks[q] * 32;
    if (in.mode == 2) {
        if (tile_cols[ntiles] >= (1LL << 31) - 8 || chunk_off[out.nslices] >= (1LL << 31)) return 1;
        out.crec.assign((size_t)(tile_cols[ntiles] + 4) * 2, -1);            // +4: phase A prefetches four records per element
        out.cnt.assign((size_t)chunk_off[out.nslices] * 16, 0);
        out.ts2.assign((size_t)out.nslices * 4, 0);
    }
    std::vector<long long> el_off(ntiles + 1, 0);
    for (int t = 0; t < ntiles; t++) el_off[t + 1] = el_off[t] + tile_nel[t];
    if (el_off[ntiles] >= (1LL << 31)) return 1;
    out.tslice_off.assign((size_t)out.nslices + 1, 0);
    for (long long s = 0; s < out.nslices; s++) out.tslice_off[s + 1] = out.tslice_off[s] + slice_sz[s];
    if (out.tslice_off[out.nslices] >= (1LL << 30)) return 1;
    out.trows.assign((size_t)row_off[ntiles] * 4, 0);
    out.tel.assign((size_t)el_off[ntiles] * 2, 0);
    out.tinc.assign((size_t)out.tslice_off[out.nslices] * 2, -1);
    out.elem_visits = el_off[ntiles];

    // 5. pass 2: fill
    size_t max_smem = 0;
    int fail = 0;
#if defined(_OPENMP)
#pragma omp parallel for schedule(dynamic, 64) reduction(max : max_smem) reduction(| : fail)
#endif
    for (int t = 0; t < ntiles; t++) {
        const int q0 = tile_first[t], n = tile_first[t + 1] - q0;
        std::vector<int> els;
        for (int q = q0; q < q0 + n; q++) {
            const int r = order[q].second;
            for (int m = in.rinc_ptr[r]; m < in.rinc_ptr[r + 1]; m++) els.push_back(in.rinc[m] / nsize);
        }
        std::sort(els.begin(), els.end());
        els.erase(std::unique(els.begin(), els.end()), els.end());
        const int nel = (int)els.size();
        // which local dofs of each element have their row in this tile
        std::vector<unsigned char> mask(nel, 0);
        for (int q = q0; q < q0 + n; q++) {
            const int r = order[q].second;
            for (int m = in.rinc_ptr[r]; m < in.rinc_ptr[r + 1]; m++) {
                const int code = in.rinc[m], e = code / nsize, k = code - e * nsize;
                const int idx = (int)(std::lower_bound(els.begin(), els.end(), e) - els.begin());
                mask[idx] |= (unsigned char)(1u << k);
            }
        }
        std::vector<int> base(nel + 1, 0);
        for (int i = 0; i < nel; i++) base[i + 1] = base[i] + __builtin_popcount(mask[i]);
        if (base[nel] >= (1 << 24)) fail |= 1;
        int *te = out.tel.data() + (size_t)el_off[t] * 2;
        for (int i = 0; i < nel; i++) {
            const int e = els[i];
            const int *dof = in.erec + (size_t)e * in.rec_ints + npe;
            bool dbc = false;
            for (int j = 0; j < nsize; j++) dbc |= dof[j] < 0;
            te[2 * i] = e | (dbc ? (int)0x80000000u : 0);
            te[2 * i + 1] = base[i] | ((int)mask[i] << 24);
        }
        // mode 2: contribution-buffer offset of every 32-row slice of the tile
        std::vector<int> cb_slice(tile_pad[t] / 32 + 1, 0);
        if (in.mode == 2)
            for (int sidx = 0; sidx < tile_pad[t] / 32; sidx++) {
                const long long g = slice_cnt[t] + sidx;
                cb_slice[sidx + 1] = cb_slice[sidx] + slice_width[g] * TILE_CB_STRIDE;
                int *q4 = out.ts2.data() + (size_t)g * 4;
                q4[0] = cb_slice[sidx]; q4[1] = (int)chunk_off[g]; q4[2] = (int)slice_chunks[g]; q4[3] = slice_width[g];
            }
        // rows, accumulator offsets, incidence entries (SELL-32 inside the tile: slice = 32 tile rows, column-major)
        int *tr = out.trows.data() + (size_t)row_off[t] * 4;
        int acc = 0;
        for (int i = 0; i < n; i++) {
            const int r = order[q0 + i].second;
            tr[4 * i] = r;
            tr[4 * i + 1] = acc;
            tr[4 * i + 2] = in.rowptr[r];
            tr[4 * i + 3] = in.rowptr[r + 1] - in.rowptr[r];
            acc += in.rowptr[r + 1] - in.rowptr[r];
            const long long so = out.tslice_off[slice_cnt[t] + i / 32] + (i & 31);
            const long long ao = in.ainc_off[r >> 5] + (r & 31);
            int mm = 0;
            if (in.mode == 2) {
                // positions of the row's contributions: runs in slot order (ascending element inside a run), then Flocal
                const int len = in.rowptr[r + 1] - in.rowptr[r], ninc = in.rinc_ptr[r + 1] - in.rinc_ptr[r];
                int cntv[256] = {0}, start[256], fill[256] = {0};
                for (int m2 = 0; m2 < ninc; m2++) {
                    const unsigned int sw = (unsigned int)in.ainc[(size_t)(ao + (long long)m2 * 32) * in.ainc_words + 1];
                    for (int j = 0; j < nsize; j++) {
                        const unsigned int sl = (sw >> (8 * j)) & 255u;
                        if (sl != 255u) { if ((int)sl >= len) fail |= 1; else cntv[sl]++; }
                    }
                }
                int run = 0;
                for (int q2 = 0; q2 < len; q2++) { start[q2] = run; run += cntv[q2]; }
                const int startF = run;
                const long long sl_idx = slice_cnt[t] + i / 32;
                const int rb = cb_slice[i / 32] + (i & 31);
                if (rb >= 65536) fail |= 1;
                unsigned char *cb = out.cnt.data() + (size_t)(chunk_off[sl_idx] + (i & 31)) * 16;
                for (int q2 = 0; q2 <= len; q2++) {
                    const int v = q2 < len ? cntv[q2] : ninc;
                    if (v > 255) fail |= 1;
                    cb[(size_t)(q2 / 16) * 32 * 16 + (q2 % 16)] = (unsigned char)v;
                }
                for (int m2 = in.rinc_ptr[r]; m2 < in.rinc_ptr[r + 1]; m2++, mm++) {
                    const int code = in.rinc[m2], e = code / nsize, k = code - e * nsize;
                    const int idx = (int)(std::lower_bound(els.begin(), els.end(), e) - els.begin());
                    const int soff = base[idx] + __builtin_popcount(mask[idx] & ((1u << k) - 1u));
                    const int *src = in.ainc + (size_t)(ao + (long long)mm * 32) * in.ainc_words;
                    if (src[0] != code) fail |= 1;
                    const unsigned int sw = (unsigned int)src[1];
                    unsigned int pj[4] = {255u, 255u, 255u, 255u};
                    for (int j = 0; j < nsize; j++) {
                        const unsigned int sl = (sw >> (8 * j)) & 255u;
                        if (sl != 255u && (int)sl < len) pj[j] = (unsigned int)(start[sl] + fill[sl]++);
                    }
                    const unsigned int pF = (unsigned int)(startF + mm);
                    int *dst = out.crec.data() + (size_t)(tile_cols[t] + soff) * 2;
                    dst[0] = (int)((unsigned int)rb | (pj[0] << 16) | (pj[1] << 24));
                    dst[1] = (int)(pj[2] | (pj[3] << 8) | (pF << 16));
                }
                continue;
            }
            for (int m = in.rinc_ptr[r]; m < in.rinc_ptr[r + 1]; m++, mm++) {
                const int code = in.rinc[m], e = code / nsize, k = code - e * nsize;
                const int idx = (int)(std::lower_bound(els.begin(), els.end(), e) - els.begin());
                const int soff = base[idx] + __builtin_popcount(mask[idx] & ((1u << k) - 1u));
                const int *src = in.ainc + (size_t)(ao + (long long)mm * 32) * in.ainc_words;
                if (src[0] != code) fail |= 1;                     // the two streams must list the same incidences
                int *dst = out.tinc.data() + (size_t)(so + (long long)mm * 32) * 2;
                dst[0] = soff;
                dst[1] = src[1];
            }
        }
        for (int i = n; i < tile_pad[t]; i++) { tr[4 * i] = -1; tr[4 * i + 1] = acc; }
        int *td = out.tdesc.data() + (size_t)t * TILE_DESC_INTS;
        td[TD_ROW_OFF] = (int)row_off[t]; td[TD_NROWS_PAD] = tile_pad[t]; td[TD_EL_OFF] = (int)el_off[t]; td[TD_NEL] = nel;
        td[TD_SLICE0] = (int)slice_cnt[t]; td[TD_NNZ] = acc; td[TD_NCOLS] = base[nel]; td[TD_NROWS] = n;
        if (in.mode == 2) {
            td[TD2_CREC_OFF] = (int)tile_cols[t];
            td[TD2_CBUF] = cb_slice[tile_pad[t] / 32];
            if ((long long)base[nel] != tile_cols[t + 1] - tile_cols[t]) fail |= 1;
            max_smem = std::max(max_smem, tile2_smem_bytes(cb_slice[tile_pad[t] / 32], acc));
        } else
        max_smem = std::max(max_smem, tile_smem_bytes(base[nel], acc, in.cta_threads));
    }
    if (fail) return 1;
    out.max_smem = max_smem;
    // elements with at least one row here
    {
        long long touched = 0;
        std::vector<unsigned char> seen;
        // count via the incidence lists: an element is touched when it appears in some row's list; count first appearances
        // by its lowest owned local dof
        for (int r = 0; r < nloc; r++)
            for (int m = in.rinc_ptr[r]; m < in.rinc_ptr[r + 1]; m++) {
                const int code = in.rinc[m], e = code / nsize, k = code - e * nsize;
                const int *dof = in.erec + (size_t)e * in.rec_ints + npe;
                bool lowest = true;
                for (int j = 0; j < k; j++) lowest &= !(dof[j] >= in.row_lo && dof[j] < in.row_lo + nloc);
                touched += lowest;
            }
        out.elems_touched = touched;
    }
    return 0;
}

}  // namespace pfem
