// assembly_ctile.cu -- host side of the colour-scheduled tile value pass (kernel: assembly_ctile.cuh).
//
// Everything the kernel streams is built here ON THE GPU, once per pattern (it replaces the per-call work PETSc does inside
// MatSetValues: locating the row, searching the column, ordering the additions; tetrapoissonparallelimpl1.F:845-884):
//   1. row -> node map, bounding box, Morton key of every owned row's node, radix sort            => spatial row order
//   2. rows cut into tiles of TR consecutive Morton positions, rows of a tile re-sorted by row id  => contiguous write-out runs
//   3. every element emits one (tile, element) visit per distinct tile among its owned rows; stable radix sort by tile
//   4. greedy colouring of each tile's visits (one warp per tile, masks in shared memory): two visits that have the same
//      owned row at the same local position never share a round; rounds are filled evenly (rotating first choice)
//   5. per tile: halo nodes collected in a shared-memory hash set, sorted (deterministic numbering), node table
//      {x, y, z, applied value} written, visit records {4 x u16 local node, 4 slot words} written in round order
// All of it is deterministic (radix sorts, sorted halo lists, sequential greedy), so the value pass is run-to-run reproducible.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "assembly_ctile.cuh"
#include "internal.cuh"

namespace pfem {

namespace {

constexpr int HCAP = 4096;            // hash-set capacity per tile (halo nodes: at most HCAP/2 - 1); power of two
constexpr int ROUND_SLOTS = CT_MAX_ROUNDS + 1;

__global__ void ct_row_node_kernel(int nElem, int npe, int rec_ints, const int *__restrict__ erec, int row_lo, int row_hi,
                                   int *__restrict__ row_node)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)nElem * npe; t += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(t / npe), k = (int)(t - (long long)e * npe);
        const int *rec = erec + (size_t)e * rec_ints;
        const int d = rec[npe + k];
        if (d >= row_lo && d < row_hi) row_node[d - row_lo] = rec[k];      // every writer stores the same node id
    }
}

__device__ __forceinline__ unsigned long long ord_encode(double v)
{
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
static double ord_decode(unsigned long long u)
{
    const unsigned long long b = (u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFULL) : ~u;
    double v;
    memcpy(&v, &b, sizeof v);
    return v;
}

// mnmx[0..2] = min, mnmx[3..5] = max (order-preserving encoding) over the nodes of the owned rows
__global__ void ct_bbox_kernel(int nloc, const int *__restrict__ row_node, const double *__restrict__ xyz, int xstride, int ndim,
                               unsigned long long *__restrict__ mnmx)
{
    unsigned long long mn[3] = {~0ULL, ~0ULL, ~0ULL}, mx[3] = {0ULL, 0ULL, 0ULL};
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += gridDim.x * blockDim.x) {
        const int n = row_node[r];
        if (n < 0) continue;
        for (int d = 0; d < ndim; d++) {
            const unsigned long long u = ord_encode(xyz[(size_t)n * xstride + d]);
            mn[d] = u < mn[d] ? u : mn[d];
            mx[d] = u > mx[d] ? u : mx[d];
        }
    }
    for (int d = 0; d < ndim; d++) {
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn[d], o), b = __shfl_xor_sync(0xffffffffu, mx[d], o);
            mn[d] = a < mn[d] ? a : mn[d];
            mx[d] = b > mx[d] ? b : mx[d];
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(mnmx + d, mn[d]); atomicMax(mnmx + 3 + d, mx[d]); }
    }
}

__device__ __forceinline__ unsigned long long spread3(unsigned long long x)     // 21 bits -> every third bit
{
    x &= 0x1fffffULL;
    x = (x | (x << 32)) & 0x1f00000000ffffULL;
    x = (x | (x << 16)) & 0x1f0000ff0000ffULL;
    x = (x | (x << 8)) & 0x100f00f00f00f00fULL;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ULL;
    x = (x | (x << 2)) & 0x1249249249249249ULL;
    return x;
}
__device__ __forceinline__ unsigned long long spread2(unsigned long long x)     // 31 bits -> every second bit
{
    x &= 0x7fffffffULL;
    x = (x | (x << 16)) & 0x0000ffff0000ffffULL;
    x = (x | (x << 8)) & 0x00ff00ff00ff00ffULL;
    x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0fULL;
    x = (x | (x << 2)) & 0x3333333333333333ULL;
    x = (x | (x << 1)) & 0x5555555555555555ULL;
    return x;
}

struct Box { double lo[3], scale[3]; };

__global__ void ct_morton_kernel(int nloc, const int *__restrict__ row_node, const double *__restrict__ xyz, int xstride, int ndim,
                                 Box box, unsigned long long *__restrict__ key, int *__restrict__ rowid)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += gridDim.x * blockDim.x) {
        const int n = row_node[r];
        unsigned long long k = ~0ULL;                       // rows without a local element sort last
        if (n >= 0) {
            unsigned long long q[3] = {0, 0, 0};
            const double qmax = ndim == 3 ? 2097151.0 : 2147483647.0;
            for (int d = 0; d < ndim; d++) {
                double t = (xyz[(size_t)n * xstride + d] - box.lo[d]) * box.scale[d];
                t = t < 0.0 ? 0.0 : (t > qmax ? qmax : t);
                q[d] = (unsigned long long)t;
            }
            k = ndim == 3 ? (spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2)) : (spread2(q[0]) | (spread2(q[1]) << 1));
        }
        key[r] = k;
        rowid[r] = r;
    }
}

// second key: (tile of the Morton position, row id): rows of a tile in ascending row order
__global__ void ct_tilekey_kernel(int nloc, int TR, const int *__restrict__ rows_sorted, unsigned long long *__restrict__ key2)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += gridDim.x * blockDim.x)
        key2[i] = ((unsigned long long)(i / TR) << 32) | (unsigned int)rows_sorted[i];
}

__global__ void ct_rpos_kernel(int nloc, const unsigned long long *__restrict__ key2_sorted, int *__restrict__ trow_row, int *__restrict__ rpos)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += gridDim.x * blockDim.x) {
        const int row = (int)(key2_sorted[i] & 0xffffffffULL);
        trow_row[i] = row;
        rpos[row] = i;
    }
}

template <int NPE>
__device__ __forceinline__ int elem_tiles(const int *rec, int row_lo, int row_hi, const int *__restrict__ rpos, int TR, int (&tiles)[NPE])
{
    int n = 0;
#pragma unroll
    for (int k = 0; k < NPE; k++) {
        const int d = rec[NPE + k];
        if (d < row_lo || d >= row_hi) continue;
        const int t = rpos[d - row_lo] / TR;
        bool seen = false;
#pragma unroll
        for (int q = 0; q < NPE; q++) seen |= (q < n && tiles[q] == t);
        if (!seen) tiles[n++] = t;
    }
    return n;
}

template <int NPE>
__global__ void ct_visit_count_kernel(int nElem, int rec_ints, const int *__restrict__ erec, int row_lo, int row_hi,
                                      const int *__restrict__ rpos, int TR, int *__restrict__ cnt)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nElem; e += gridDim.x * blockDim.x) {
        int tiles[NPE];
        cnt[e] = elem_tiles<NPE>(erec + (size_t)e * rec_ints, row_lo, row_hi, rpos, TR, tiles);
    }
}

template <int NPE>
__global__ void ct_visit_fill_kernel(int nElem, int rec_ints, const int *__restrict__ erec, int row_lo, int row_hi,
                                     const int *__restrict__ rpos, int TR, const int *__restrict__ off, int *__restrict__ vt_tile,
                                     int *__restrict__ vt_elem)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nElem; e += gridDim.x * blockDim.x) {
        int tiles[NPE];
        const int n = elem_tiles<NPE>(erec + (size_t)e * rec_ints, row_lo, row_hi, rpos, TR, tiles);
        const int o = off[e];
        for (int q = 0; q < n; q++) { vt_tile[o + q] = tiles[q]; vt_elem[o + q] = e; }
    }
}

__global__ void ct_lower_bound_kernel(int ntiles, int total, const int *__restrict__ keys, int *__restrict__ ptr)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t <= ntiles; t += gridDim.x * blockDim.x) {
        int lo = 0, hi = total;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (keys[mid] < t) lo = mid + 1; else hi = mid;
        }
        ptr[t] = lo;
    }
}

__global__ void ct_max_rowlen_kernel(int nloc, const int *__restrict__ rowptr, int *__restrict__ out)
{
    int m = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += gridDim.x * blockDim.x) m = max(m, rowptr[r + 1] - rowptr[r]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// One warp per tile.  Lane 0 colours the tile's visits in order (sequential greedy: smallest free round at or after a
// rotating start, so rounds fill evenly); the other lanes fetch the next 32 visits' owned tile-local rows meanwhile.
// perm[v] = final position of visit v (round-major, original order inside a round); round offsets per tile.
// FULL = true : two visits that share ANY owned row get different rounds (the kernel commits a visit in one go);
// FULL = false: only visits with the same owned row at the same local position k are separated (column-by-column commits).
// B = threads of the value-pass CTA: with the full rule the first choices rotate over enough rounds for a round to fit one CTA pass.
template <int NPE, bool FULL>
__global__ void __launch_bounds__(32) ct_colour_kernel(int B, int TR, const int *__restrict__ tile_vbeg, const int *__restrict__ vt_elem,
                                                       int rec_ints, const int *__restrict__ erec, int row_lo, int row_hi,
                                                       const int *__restrict__ rpos, int nloc, unsigned char *__restrict__ vcol,
                                                       int *__restrict__ vpos, int *__restrict__ perm, int *__restrict__ round_off,
                                                       int *__restrict__ tile_nrounds, int *__restrict__ overflow)
{
    extern __shared__ unsigned long long ct_masks[];           // [NPE][TR]
    __shared__ unsigned short rl[32][NPE];
    __shared__ int cnt[CT_MAX_ROUNDS], offs[CT_MAX_ROUNDS + 1];
    __shared__ unsigned char cbuf[32];
    __shared__ int pbuf[32];
    const int tile = blockIdx.x, lane = threadIdx.x;
    const int vb = tile_vbeg[tile], ve = tile_vbeg[tile + 1];
    const int nrows = min(TR, nloc - tile * TR);
    for (int q = lane; q < (FULL ? 1 : NPE) * TR; q += 32) ct_masks[q] = 0ULL;
    for (int q = lane; q < CT_MAX_ROUNDS; q += 32) cnt[q] = 0;
    __syncwarp();
    int C0 = nrows > 0 ? (ve - vb + nrows - 1) / nrows : 1;          // ~ incidences per (row, position)
    if (FULL) C0 = max(((NPE - 1) * (ve - vb) + nrows - 1) / max(nrows, 1), (ve - vb + B - 1) / B) + 1;   // ~ incidences per row
    C0 = max(2, min(C0, 48));
    for (int chunk = vb; chunk < ve; chunk += 32) {
        const int v = chunk + lane;
        if (v < ve) {
            const int *rec = erec + (size_t)vt_elem[v] * rec_ints;
#pragma unroll
            for (int k = 0; k < NPE; k++) {
                const int d = rec[NPE + k];
                unsigned short x = 0xFFFFu;
                if (d >= row_lo && d < row_hi) {
                    const int p = rpos[d - row_lo];
                    if (p / TR == tile) x = (unsigned short)(p - tile * TR);
                }
                rl[lane][k] = x;
            }
        }
        __syncwarp();
        if (lane == 0) {
            const int n = min(32, ve - chunk);
            for (int q = 0; q < n; q++) {
                unsigned long long avail = ~0ULL;
#pragma unroll
                for (int k = 0; k < NPE; k++)
                    if (rl[q][k] != 0xFFFFu) avail &= ~ct_masks[(FULL ? 0 : k * TR) + rl[q][k]];
                int c = 0;
                if (avail == 0ULL) atomicOr(overflow, 1);
                else {
                    const int start = (chunk - vb + q) % C0;
                    const unsigned long long hi = (avail >> start) << start;
                    c = hi ? __ffsll((long long)hi) - 1 : __ffsll((long long)avail) - 1;
                }
#pragma unroll
                for (int k = 0; k < NPE; k++)
                    if (rl[q][k] != 0xFFFFu) ct_masks[(FULL ? 0 : k * TR) + rl[q][k]] |= 1ULL << c;
                cbuf[q] = (unsigned char)c;
                pbuf[q] = cnt[c]++;
            }
        }
        __syncwarp();
        if (v < ve) { vcol[v] = cbuf[lane]; vpos[v] = pbuf[lane]; }
        __syncwarp();
    }
    if (lane == 0) {
        // empty colours are squeezed out: offs[c] = start of colour c inside the tile, round_off lists the non-empty ones
        int o = 0, nr = 0;
        int *ro = round_off + (size_t)tile * ROUND_SLOTS;
        for (int c = 0; c < CT_MAX_ROUNDS; c++) {
            offs[c] = o;
            if (cnt[c] > 0) { ro[nr++] = vb + o; o += cnt[c]; }
        }
        ro[nr] = vb + o;
        tile_nrounds[tile] = nr;
    }
    __syncwarp();
    for (int v = vb + lane; v < ve; v += 32) perm[v] = vb + offs[vcol[v]] + vpos[v];
}

__device__ __forceinline__ unsigned hash_node(int n) { return ((unsigned)n * 2654435761u) >> (32 - 12); }   // HCAP = 2^12

// One CTA per tile.  pass 0: count the tile's halo nodes (nodes of its visits that are not rows of the tile).
// pass 1: number them (ascending node id), write the node table, the row descriptors and the visit records.
template <int NPE>
__global__ void __launch_bounds__(256) ct_records_kernel(int pass, int TR, int nloc, const int *__restrict__ tile_vbeg,
                                                         const int *__restrict__ vt_elem, int rec_ints, const int *__restrict__ erec,
                                                         int row_lo, int row_hi, const int *__restrict__ rpos,
                                                         const int *__restrict__ trow_row, const int *__restrict__ row_node,
                                                         const double *__restrict__ xyz, int xstride, int ndim,
                                                         const double *__restrict__ applied, const int *__restrict__ rowptr,
                                                         const int *__restrict__ col, const int *__restrict__ perm,
                                                         int *__restrict__ nhalo, const int *__restrict__ node0, double4 *__restrict__ tnode,
                                                         int4 *__restrict__ trow, uint2 *__restrict__ vnode, uint4 *__restrict__ vslot,
                                                         int *__restrict__ overflow)
{
    __shared__ int hkey[HCAP];
    __shared__ unsigned short hval[HCAP];
    __shared__ int sorted[HCAP / 2];
    __shared__ int s_count;
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int vb = tile_vbeg[tile], ve = tile_vbeg[tile + 1];
    const int nrows = min(TR, nloc - tile * TR);
    for (int q = tid; q < HCAP; q += blockDim.x) hkey[q] = -1;
    if (tid == 0) s_count = 0;
    __syncthreads();
    auto in_tile = [&](int d, int &rl) {
        if (d < row_lo || d >= row_hi) return false;
        const int p = rpos[d - row_lo];
        if (p / TR != tile) return false;
        rl = p - tile * TR;
        return true;
    };
    // ---- halo set ----
    for (int v = vb + tid; v < ve; v += blockDim.x) {
        const int *rec = erec + (size_t)vt_elem[v] * rec_ints;
#pragma unroll
        for (int k = 0; k < NPE; k++) {
            int rl;
            if (in_tile(rec[NPE + k], rl)) continue;
            const int n = rec[k];
            unsigned s = hash_node(n);
            while (true) {
                const int old = atomicCAS(&hkey[s], -1, n);
                if (old == -1) { atomicAdd(&s_count, 1); break; }
                if (old == n) break;
                s = (s + 1) & (HCAP - 1);
                if (s_count >= HCAP / 2) break;             // table too full: reported below
            }
        }
    }
    __syncthreads();
    const int nh = s_count;
    if (nh >= HCAP / 2 || nrows + nh > 65535) {
        if (tid == 0) { atomicOr(overflow, 2); if (pass == 0) nhalo[tile] = 0; }
        return;
    }
    if (pass == 0) {
        if (tid == 0) nhalo[tile] = nh;
        return;
    }
    // ---- number the halo nodes in ascending node id: compact, bitonic sort, write ranks back into the table ----
    __syncthreads();                     // every thread has read nh = s_count before thread 0 resets it (racecheck, r02)
    if (tid == 0) s_count = 0;
    int npad = 1;
    while (npad < nh) npad <<= 1;
    for (int q = tid; q < npad; q += blockDim.x) sorted[q] = 0x7fffffff;
    __syncthreads();
    for (int q = tid; q < HCAP; q += blockDim.x)
        if (hkey[q] >= 0) sorted[atomicAdd(&s_count, 1)] = hkey[q];
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npad; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const int a = sorted[i], b = sorted[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { sorted[i] = b; sorted[l] = a; }
                }
            }
            __syncthreads();
        }
    for (int p = tid; p < nh; p += blockDim.x) {
        const int n = sorted[p];
        unsigned s = hash_node(n);
        while (hkey[s] != n) s = (s + 1) & (HCAP - 1);
        hval[s] = (unsigned short)p;
    }
    __syncthreads();
    // ---- node table and row descriptors ----
    const int nb = node0[tile];
    const int row0 = tile * TR;
    for (int i = tid; i < nrows + nh; i += blockDim.x) {
        int n;
        double g = 0.0;
        if (i < nrows) {
            const int row = trow_row[row0 + i];
            n = row_node[row];
            trow[row0 + i] = make_int4(row, rowptr[row], rowptr[row + 1] - rowptr[row], 0);
        } else {
            n = sorted[i - nrows];
            g = applied[n];
        }
        double4 P = make_double4(0.0, 0.0, 0.0, g);
        if (n >= 0) {
            P.x = xyz[(size_t)n * xstride];
            P.y = xyz[(size_t)n * xstride + 1];
            if (ndim == 3) P.z = xyz[(size_t)n * xstride + 2];
        }
        tnode[nb + i] = P;
    }
    // ---- visit records, written at their round-ordered positions ----
    for (int v = vb + tid; v < ve; v += blockDim.x) {
        const int *rec = erec + (size_t)vt_elem[v] * rec_ints;
        unsigned nl[4] = {0u, 0u, 0u, 0u}, w[4] = {~0u, ~0u, ~0u, ~0u};
#pragma unroll
        for (int k = 0; k < NPE; k++) {
            int rl;
            if (in_tile(rec[NPE + k], rl)) {
                nl[k] = (unsigned)rl;
                const int row = rec[NPE + k] - row_lo;
                const int c0 = rowptr[row], len = rowptr[row + 1] - c0;
                unsigned word = 0u;
#pragma unroll
                for (int j = 0; j < NPE; j++) {
                    const int c = rec[NPE + j];
                    unsigned sl = 255u;
                    if (c >= 0) {
                        int lo = 0, hi = len;
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (col[c0 + mid] < c) lo = mid + 1; else hi = mid;
                        }
                        sl = (unsigned)lo;
                    }
                    word |= sl << (8 * j);
                }
                w[k] = word;
            } else {
                const int n = rec[k];
                unsigned s = hash_node(n);
                while (hkey[s] != n) s = (s + 1) & (HCAP - 1);
                nl[k] = (unsigned)nrows + hval[s];
            }
        }
        if (NPE == 3) nl[3] = nl[2];
        const int dst = perm[v];
        vnode[dst] = make_uint2(nl[0] | (nl[1] << 16), nl[2] | (nl[3] << 16));
        vslot[dst] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// PFEM_TRACE=1: synchronise and report the stage that just finished (stderr)
void trace_point(pfem_solver *h, const char *what)
{
    static double t_last = 0.0;
    const char *e = getenv("PFEM_TRACE");
    if (!e || e[0] != '1') return;
    const cudaError_t err = cudaStreamSynchronize(h->stream);
    const double now = StageTimer::now();
    fprintf(stderr, "[pfem trace]   %-26s %8.3f ms since the previous point (%s)\n", what, t_last > 0.0 ? 1e3 * (now - t_last) : 0.0,
            cudaGetErrorString(err));
    t_last = now;
}

template <typename K, typename V>
int sort_pairs(pfem_solver *h, const K *kin, K *kout, const V *vin, V *vout, int n, int end_bit)
{
    size_t bytes = 0;
    PFEM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, end_bit, h->stream));
    char *tmp = nullptr;
    PFEM_TRY(scratch_get<char>(h, 4, bytes, &tmp));
    PFEM_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, n, 0, end_bit, h->stream));
    h->launches++;
    return PFEM_OK;
}

int bits_for(long long n)
{
    int b = 1;
    while ((1LL << b) <= n) b++;
    return b;
}

template <int NPE>
int build_ctiles_kind(pfem_solver *h)
{
    cudaStream_t s = h->stream;
    const int G = h->sm_count * 8, nloc = h->size_local, nElem = h->nElem;
    const int xstride = h->ndim == 3 ? 4 : 2;
    h->ct_ready = false;
    if (nloc <= 0 || nElem <= 0) return PFEM_OK;
    int max_smem = 0;
    PFEM_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    // ---- stride of the accumulators = longest row ----
    DevBuf<int> dflag;
    PFEM_TRY(dflag.alloc(4));
    PFEM_CUDA(cudaMemsetAsync(dflag.p, 0, 4 * sizeof(int), s));
    ct_max_rowlen_kernel<<<G, 256, 0, s>>>(nloc, h->rowptr.p, dflag.p);
    int stride = 0;
    PFEM_CUDA(cudaMemcpyAsync(&stride, dflag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    if (stride <= 0 || stride > 254) return PFEM_OK;            // slot bytes cannot address the row: the row kernels handle it
    // ---- 1. spatial order of the rows ----
    int *row_node = nullptr, *rowid = nullptr, *rows_sorted = nullptr, *rpos = nullptr, *trow_row = nullptr;
    unsigned long long *key = nullptr, *key_sorted = nullptr;
    {
        char *base = nullptr;
        const size_t n8 = ((size_t)nloc + 64) * 8;
        PFEM_TRY(scratch_get<char>(h, 5, n8 * 2 + ((size_t)nloc + 64) * 4 * 5, &base));
        key = reinterpret_cast<unsigned long long *>(base);
        key_sorted = reinterpret_cast<unsigned long long *>(base + n8);
        int *ib = reinterpret_cast<int *>(base + 2 * n8);
        const size_t n4 = (size_t)nloc + 64;
        row_node = ib; rowid = ib + n4; rows_sorted = ib + 2 * n4; rpos = ib + 3 * n4; trow_row = ib + 4 * n4;
    }
    PFEM_CUDA(cudaMemsetAsync(row_node, 0xFF, (size_t)nloc * sizeof(int), s));
    ct_row_node_kernel<<<G, 256, 0, s>>>(nElem, NPE, h->rec_ints, h->erec.p, h->row_lo, h->row_hi, row_node);
    DevBuf<unsigned long long> mnmx;
    PFEM_TRY(mnmx.alloc(6));
    PFEM_CUDA(cudaMemsetAsync(mnmx.p, 0xFF, 3 * sizeof(unsigned long long), s));
    PFEM_CUDA(cudaMemsetAsync(mnmx.p + 3, 0, 3 * sizeof(unsigned long long), s));
    ct_bbox_kernel<<<G, 256, 0, s>>>(nloc, row_node, h->xyz.p, xstride, h->ndim, mnmx.p);
    unsigned long long hm[6];
    PFEM_CUDA(cudaMemcpyAsync(hm, mnmx.p, sizeof hm, cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    Box box;
    const double qmax = h->ndim == 3 ? 2097151.0 : 2147483647.0;
    for (int d = 0; d < 3; d++) {
        box.lo[d] = 0.0; box.scale[d] = 0.0;
        if (d < h->ndim && hm[3 + d] >= hm[d]) {
            const double lo = ord_decode(hm[d]), hi = ord_decode(hm[3 + d]);
            box.lo[d] = lo;
            box.scale[d] = hi > lo ? qmax / (hi - lo) : 0.0;
        }
    }
    ct_morton_kernel<<<G, 256, 0, s>>>(nloc, row_node, h->xyz.p, xstride, h->ndim, box, key, rowid);
    h->launches += 4;
    PFEM_TRY((sort_pairs<unsigned long long, int>(h, key, key_sorted, rowid, rows_sorted, nloc, 64)));
    trace_point(h, "ctile: morton order");

    // ---- CTA shape and commit rule of the value-pass kernel (fixed at build time: the rounds depend on them) ----
    const char *benv = getenv("PFEM_TILE_THREADS"), *renv = getenv("PFEM_TILE_RULE");
    int B = benv ? atoi(benv) : 384;
    if (B != 1024 && B != 768 && B != 512 && B != 384 && B != 256) B = 384;
    bool full = !(renv && !strcmp(renv, "position"));
    // ---- tile size: accumulators + RHS + row descriptors + node table (own + ~45 % halo) must fit in shared memory ----
    const char *env = getenv("PFEM_TILE_ROWS");
    int TR = env ? atoi(env) : 1024;
    const size_t budget = (size_t)max_smem - 2048;
    {
        const size_t per_row = (size_t)stride * 8 + 8 + 16 + 48;
        const int fit = (int)(budget / per_row);
        if (TR > fit) TR = fit;
        TR = std::max(32, (TR / 32) * 32);
    }
    for (int attempt = 0; attempt < 4; attempt++) {
        const int ntiles = (nloc + TR - 1) / TR;
        // ---- 2. rows of a tile in ascending row order ----
        ct_tilekey_kernel<<<G, 256, 0, s>>>(nloc, TR, rows_sorted, key);
        {
            size_t bytes = 0;
            const int eb = 32 + bits_for(ntiles);
            PFEM_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, key, key_sorted, nloc, 0, eb, s));
            char *tmp = nullptr;
            PFEM_TRY(scratch_get<char>(h, 4, bytes, &tmp));
            PFEM_CUDA(cub::DeviceRadixSort::SortKeys(tmp, bytes, key, key_sorted, nloc, 0, eb, s));
        }
        ct_rpos_kernel<<<G, 256, 0, s>>>(nloc, key_sorted, trow_row, rpos);
        h->launches += 3;
        // ---- 3. visits ----
        int *cnt = nullptr, *off = nullptr;
        PFEM_TRY(scratch_get<int>(h, 0, (size_t)nElem + 1, &cnt));
        PFEM_TRY(scratch_get<int>(h, 1, (size_t)nElem + 1, &off));
        PFEM_CUDA(cudaMemsetAsync(cnt + nElem, 0, sizeof(int), s));
        ct_visit_count_kernel<NPE><<<G, 256, 0, s>>>(nElem, h->rec_ints, h->erec.p, h->row_lo, h->row_hi, rpos, TR, cnt);
        {
            size_t bytes = 0;
            PFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt, off, nElem + 1, s));
            char *tmp = nullptr;
            PFEM_TRY(scratch_get<char>(h, 4, bytes, &tmp));
            PFEM_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, cnt, off, nElem + 1, s));
        }
        h->launches += 2;
        int V = 0;
        PFEM_CUDA(cudaMemcpyAsync(&V, off + nElem, sizeof(int), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaStreamSynchronize(s));
        if (V <= 0) return PFEM_OK;
        // cnt/off live in scratch 0/1: the visit arrays go to 2, 3, 6, 7 (scratch 4 = cub temp, 5 = row arrays)
        int *vt_tile = nullptr, *vt_elem = nullptr, *vt_tile_s = nullptr, *vt_elem_s = nullptr;
        PFEM_TRY(scratch_get<int>(h, 2, (size_t)V, &vt_tile));
        PFEM_TRY(scratch_get<int>(h, 3, (size_t)V, &vt_elem));
        PFEM_TRY(scratch_get<int>(h, 6, (size_t)V, &vt_tile_s));
        PFEM_TRY(scratch_get<int>(h, 7, (size_t)V, &vt_elem_s));
        ct_visit_fill_kernel<NPE><<<G, 256, 0, s>>>(nElem, h->rec_ints, h->erec.p, h->row_lo, h->row_hi, rpos, TR, off, vt_tile, vt_elem);
        h->launches++;
        PFEM_TRY((sort_pairs<int, int>(h, vt_tile, vt_tile_s, vt_elem, vt_elem_s, V, bits_for(ntiles))));
        trace_point(h, "ctile: visits sorted");
        // from here on: scratch 0 = tile_vbeg + per-tile ints, scratch 1 = perm, scratch 2 = vpos, scratch 3 = vcol
        int *tile_ints = nullptr, *perm = nullptr, *vpos = nullptr;
        unsigned char *vcol = nullptr;
        PFEM_TRY(scratch_get<int>(h, 0, (size_t)ntiles * 4 + 8, &tile_ints));
        PFEM_TRY(scratch_get<int>(h, 1, (size_t)V, &perm));
        PFEM_TRY(scratch_get<int>(h, 2, (size_t)V, &vpos));
        PFEM_TRY(scratch_get<unsigned char>(h, 3, (size_t)V, &vcol));
        int *tile_vbeg = tile_ints, *tile_nrounds = tile_ints + ntiles + 1, *nhalo = tile_nrounds + ntiles, *node0 = nhalo + ntiles;
        ct_lower_bound_kernel<<<G, 256, 0, s>>>(ntiles, V, vt_tile_s, tile_vbeg);
        // ---- 4. colouring ----
        PFEM_TRY(h->ct_round_off.alloc((size_t)ntiles * ROUND_SLOTS));
        PFEM_CUDA(cudaMemsetAsync(dflag.p + 1, 0, sizeof(int), s));
        if (full) {
            const size_t csm = (size_t)TR * sizeof(unsigned long long);
            PFEM_CUDA(cudaFuncSetAttribute(ct_colour_kernel<NPE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
            ct_colour_kernel<NPE, true><<<ntiles, 32, csm, s>>>(B, TR, tile_vbeg, vt_elem_s, h->rec_ints, h->erec.p, h->row_lo, h->row_hi, rpos,
                                                              nloc, vcol, vpos, perm, h->ct_round_off.p, tile_nrounds, dflag.p + 1);
        } else {
            const size_t csm = (size_t)NPE * TR * sizeof(unsigned long long);
            PFEM_CUDA(cudaFuncSetAttribute(ct_colour_kernel<NPE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
            ct_colour_kernel<NPE, false><<<ntiles, 32, csm, s>>>(B, TR, tile_vbeg, vt_elem_s, h->rec_ints, h->erec.p, h->row_lo, h->row_hi, rpos,
                                                               nloc, vcol, vpos, perm, h->ct_round_off.p, tile_nrounds, dflag.p + 1);
        }
        trace_point(h, "ctile: colouring");
        // ---- 5. halo counts, node offsets, records ----
        PFEM_TRY(h->ct_trow.alloc((size_t)ntiles * TR * 4));
        ct_records_kernel<NPE><<<ntiles, 256, 0, s>>>(0, TR, nloc, tile_vbeg, vt_elem_s, h->rec_ints, h->erec.p, h->row_lo, h->row_hi, rpos,
                                                     trow_row, row_node, h->xyz.p, xstride, h->ndim, h->applied.p, h->rowptr.p, h->col.p,
                                                     perm, nhalo, nullptr, nullptr, nullptr, nullptr, nullptr, dflag.p + 1);
        h->launches += 3;
        std::vector<int> hn(ntiles), hr(ntiles), hv(ntiles + 1);
        int flag = 0;
        PFEM_CUDA(cudaMemcpyAsync(hn.data(), nhalo, (size_t)ntiles * sizeof(int), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaMemcpyAsync(hr.data(), tile_nrounds, (size_t)ntiles * sizeof(int), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaMemcpyAsync(hv.data(), tile_vbeg, ((size_t)ntiles + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaMemcpyAsync(&flag, dflag.p + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaStreamSynchronize(s));
        trace_point(h, "ctile: halo count");
        if ((flag & 1) && full) {              // more than 64 rounds with the full rule (high-valence nodes): per-position rule
            full = false;
            attempt--;
            continue;
        }
        if (flag) return PFEM_OK;              // more than 64 rounds, or a halo beyond the hash set: the row kernels handle it
        std::vector<int> hnode0(ntiles), desc((size_t)ntiles * CT_DESC);
        long long nodes_total = 0;
        int node_cap = 0, max_rounds = 0;
        for (int t = 0; t < ntiles; t++) {
            const int nrows = std::min(TR, nloc - t * TR);
            hnode0[t] = (int)nodes_total;
            nodes_total += nrows + hn[t];
            node_cap = std::max(node_cap, nrows + hn[t]);
            max_rounds = std::max(max_rounds, hr[t]);
            int *d = &desc[(size_t)t * CT_DESC];
            d[CT_ROW0] = t * TR; d[CT_NROWS] = nrows; d[CT_NODE0] = hnode0[t]; d[CT_NNODES] = nrows + hn[t];
            d[CT_ROUND0] = t * ROUND_SLOTS; d[CT_NROUNDS] = hr[t]; d[CT_STRIDE] = stride; d[CT_NVISITS] = hv[t + 1] - hv[t];
        }
        if (nodes_total >= (1LL << 31)) return PFEM_OK;
        const size_t smem = (size_t)node_cap * 32 + (size_t)TR * 16 + (size_t)TR * stride * 8 + (size_t)TR * 8 + (size_t)B * 8 + 64;
        if (smem > budget) {                   // the halo was larger than planned: smaller tiles
            TR = std::max(32, ((TR * 3 / 4) / 32) * 32);
            continue;
        }
        PFEM_CUDA(cudaMemcpyAsync(node0, hnode0.data(), (size_t)ntiles * sizeof(int), cudaMemcpyHostToDevice, s));
        PFEM_TRY(h->ct_tdesc.alloc(desc.size()));
        PFEM_CUDA(cudaMemcpyAsync(h->ct_tdesc.p, desc.data(), desc.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        PFEM_TRY(h->ct_tnode.alloc((size_t)nodes_total * 4));
        PFEM_TRY(h->ct_vnode.alloc((size_t)V * 2));
        PFEM_TRY(h->ct_vslot.alloc((size_t)V * 4));
        ct_records_kernel<NPE><<<ntiles, 256, 0, s>>>(1, TR, nloc, tile_vbeg, vt_elem_s, h->rec_ints, h->erec.p, h->row_lo, h->row_hi, rpos,
                                                     trow_row, row_node, h->xyz.p, xstride, h->ndim, h->applied.p, h->rowptr.p, h->col.p,
                                                     perm, nhalo, node0, reinterpret_cast<double4 *>(h->ct_tnode.p),
                                                     reinterpret_cast<int4 *>(h->ct_trow.p), reinterpret_cast<uint2 *>(h->ct_vnode.p),
                                                     reinterpret_cast<uint4 *>(h->ct_vslot.p), dflag.p + 1);
        h->launches++;
        PFEM_CUDA(cudaGetLastError());
        PFEM_CUDA(cudaStreamSynchronize(s));
        h->ct_ntiles = ntiles; h->ct_TR = TR; h->ct_stride = stride; h->ct_node_cap = node_cap; h->ct_visits = V;
        h->ct_max_rounds = max_rounds; h->ct_smem = smem; h->ct_threads = B; h->ct_full = full;
        h->ct_ready = true;
        return PFEM_OK;
    }
    return PFEM_OK;
}

template <int KIND, int B, bool SUBSYNC>
int launch_ctile(pfem_solver *h, const CtileArgs &a)
{
    PFEM_CUDA(cudaFuncSetAttribute(assemble_ctile_kernel<KIND, B, SUBSYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ct_smem));
    int grid = std::min(h->ct_ntiles, h->sm_count);
    const char *genv = getenv("PFEM_TILE_GRID");              // test hook: fewer CTAs => several tiles per CTA on small meshes
    if (genv && atoi(genv) > 0) grid = std::min(grid, atoi(genv));
    assemble_ctile_kernel<KIND, B, SUBSYNC><<<grid, B, h->ct_smem, h->stream>>>(a);
    h->launches++;
    PFEM_CUDA(cudaGetLastError());
    return PFEM_OK;
}

template <int KIND, bool SUBSYNC>
int launch_ctile_b(pfem_solver *h, const CtileArgs &a)
{
    switch (h->ct_threads) {
    case 1024: return launch_ctile<KIND, 1024, SUBSYNC>(h, a);
    case 768: return launch_ctile<KIND, 768, SUBSYNC>(h, a);
    case 512: return launch_ctile<KIND, 512, SUBSYNC>(h, a);
    case 256: return launch_ctile<KIND, 256, SUBSYNC>(h, a);
    default: return launch_ctile<KIND, 384, SUBSYNC>(h, a);
    }
}

}  // namespace

// Tiles of the current pattern (built once per pattern; the applied values travel inside the node tables, so
// pfem_solver_set_applied marks them stale).  On return h->ct_ready tells whether the tile kernel can run; when it cannot
// (rows longer than 254 entries, more than 64 rounds, a halo beyond the hash set) the row kernels do the pass.
int build_ctiles(pfem_solver *h)
{
    StageTimer tm("value pass: tile construction (GPU)");
    if (h->ndof != 1) { h->ct_ready = false; return PFEM_OK; }
    return h->npe == 4 ? build_ctiles_kind<4>(h) : build_ctiles_kind<3>(h);
}

int assemble_values_ctile(pfem_solver *h, const double *dElemData, const double *dTimeData)
{
    CtileArgs a;
    a.ntiles = h->ct_ntiles;
    a.tdesc = h->ct_tdesc.p;
    a.trow = reinterpret_cast<const int4 *>(h->ct_trow.p);
    a.tnode = reinterpret_cast<const double4 *>(h->ct_tnode.p);
    a.round_off = h->ct_round_off.p;
    a.vnode = reinterpret_cast<const uint2 *>(h->ct_vnode.p);
    a.vslot = reinterpret_cast<const uint4 *>(h->ct_vslot.p);
    a.val = h->val.p; a.rhs = h->rhs.p;
    a.elemData = dElemData; a.timeData = dTimeData;
    a.neg_flag = h->neg_count.p;
    a.load_val = h->values_zero ? 0 : 1; a.load_rhs = h->rhs_zero ? 0 : 1;
    a.node_cap = h->ct_node_cap;
    a.acc_cap = h->ct_TR * h->ct_stride;
    a.row_cap = h->ct_TR;
    if (h->kind == PFEM_POISSON_TETRA)
        return h->ct_full ? launch_ctile_b<POISSON_TETRA, false>(h, a) : launch_ctile_b<POISSON_TETRA, true>(h, a);
    return h->ct_full ? launch_ctile_b<POISSON_TRIA, false>(h, a) : launch_ctile_b<POISSON_TRIA, true>(h, a);
}

}  // namespace pfem
