// internal.cuh -- solver handle and cross-TU declarations of libpfemb200.so (not part of the ABI)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/pfem_b200.h"

namespace pfem {

void set_error(const char *fmt, ...);

#define PFEM_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            pfem::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return PFEM_ERR_CUDA;                                                            \
        }                                                                                    \
    } while (0)

#define PFEM_TRY(call)                  \
    do {                                \
        int s_ = (call);                \
        if (s_ != PFEM_OK) return s_;   \
    } while (0)

// Simple owning device buffer.
template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    // Re-use the existing allocation when it is large enough (and not wastefully larger): repeated pattern passes
    // (every end-to-end step) then cost no cudaMalloc/cudaFree round trips.  Contents are NOT cleared.
    int alloc(size_t count) {
        if (count == 0) count = 1;
        if (p && n >= count && n <= 2 * count + 4096) return PFEM_OK;
        release();
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
            return PFEM_ERR_CUDA;
        }
        n = count;
        return PFEM_OK;
    }
    // grow-only variant for scratch space: never shrinks, so alternating uses of different sizes do not thrash
    // (a cudaFree + cudaMalloc pair of ~1 GB costs ~100 ms of page mapping work)
    int reserve(size_t count) {
        if (count == 0) count = 1;
        if (p && n >= count) return PFEM_OK;
        release();
        return alloc(count);
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

// Device-resident CG scalar state (one per solver).  Kernels early-out when reason != 0.
struct CgState {
    double beta, betaold, dpi, dpiold, dp, a, b, ttol, rnorm0, delta;
    double rtol, abstol, dtol;
    double red[4];          // reduction landing zone (local sums; all-reduced in place for nranks > 1)
    int its, reason, iter, max_it;
    unsigned int ticket[4]; // last-block-done counters
    unsigned int ticket2[4];
    unsigned long long seq; // solve sequence number: high half of every peer-exchange tag
};

struct NcclApi;             // comm.cu

// Peer-memory (NVLink) exchange area of one rank; every peer maps it through CUDA IPC and writes into it directly.
constexpr int P2P_MAX_RANKS = 16;
struct P2pSlot { double value; unsigned long long tag; };   // written with ONE 16-byte store: the tag validates the value
struct P2pMail {
    unsigned long long halo_flag[P2P_MAX_RANKS];        // peer q: "my boundary values for tag t are in your ghost buffer"
    P2pSlot red[2][P2P_MAX_RANKS][4];                   // [phase][peer][value index]
};
struct P2pCtx {
    int rank, nranks;
    unsigned long long seq;                             // solve sequence number (high half of every tag)
    P2pMail *mail[P2P_MAX_RANKS];                       // mail[rank] is the local one
    int sends_to[P2P_MAX_RANKS], recvs_from[P2P_MAX_RANKS];
};

// SELL-32 storage of the diagonal block (columns owned by this rank), built at pattern time.
struct SellMatrix {
    int nrows = 0, nslices = 0;
    long long nstored = 0;                    // padded entries
    DevBuf<long long> slice_off;              // [nslices+1] offsets (in entries)
    DevBuf<int> col;                          // [nstored] local column index
    DevBuf<double> val;                       // [nstored]
};

}  // namespace pfem

struct pfem_solver {
    int device = 0, rank = 0, nranks = 1;
    int sm_count = 148;
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_halo = nullptr, ev_pack = nullptr;
    int state = PFEM_SOLVER_EMPTY;
    bool initialised = false;
    long long launches = 0;

    // sizes (solverpetsc.F:116-131)
    int size_local = 0, size_global = 0, row_lo = 0, row_hi = 0;   // owned rows [row_lo,row_hi), 0-based
    std::vector<int> row_starts;                                   // [nranks+1]

    // options
    double rtol = 1e-5, abstol = 1e-50, dtol = 1e4;
    int max_it = 10000, pc_type = PFEM_PC_JACOBI;

    // mesh resident in HBM
    int kind = -1, npe = 0, ndof = 0, ndim = 0, nsize = 0;
    int nElem = 0, nNode = 0;
    int rec_ints = 0;                                              // ints per element record
    pfem::DevBuf<int> erec;        // AoS per element: conn[npe] (0-based NEW node id) then dof[nsize], padded to rec_ints
    pfem::DevBuf<double> xyz;      // AoS per NEW node: (x,y[,z,pad])
    pfem::DevBuf<double> applied;  // solnApplied, NEW numbering
    bool have_mesh = false, have_dofs = false;

    // pattern (owned rows, global columns)
    long long nnz = 0, ninc = 0;
    pfem::DevBuf<int> rowptr, col;
    pfem::DevBuf<double> val, rhs;
    pfem::DevBuf<int> rinc_ptr, rinc;      // row -> (e*nsize + k) incidences, ascending
    // value-pass streams: per 32-row slice, column-major incidence entries {code, slot bytes} (warp-coalesced),
    // and a conn-only int4 record per element
    pfem::DevBuf<long long> ainc_off;      // [nslices+1] entry offsets
    pfem::DevBuf<int> ainc;                // entries of ainc_words ints: code (or -1 = padding), then nsize slot bytes
    pfem::DevBuf<int> conn4;               // [nElem][4] 0-based NEW node ids
    int ainc_words = 0;
    bool asm_sell = false;                 // false: rows wider than 254 entries -> generic (binary search) kernel
    bool values_zero = true, rhs_zero = true;
    int asm_rows_per_cta = 0, asm_max_seg = 0;
    size_t asm_smem = 0;
    pfem::DevBuf<int> neg_count;
    // tiled (compute-once) value pass, opt-in (PFEM_ASM=tiled): tiles.hpp / assembly_tiled.cu
    bool tiles_ready = false, asm_tiled = false;
    int last_asm_mode = 0;                 // kernel of the last value pass: 0 generic row gather, 1 streamed row gather, 2 tiled
    int ntiles = 0, tile_threads = 0;
    size_t tile_smem = 0;
    long long tile_elem_visits = 0, tile_elems_touched = 0;
    int tile_mode = 0;                     // 1: gather kernel (PFEM_ASM=tiled), 2: scatter kernel (PFEM_ASM=tiled2)
    pfem::DevBuf<int> t_desc, t_rows, t_el, t_inc, t_crec, t_ts2;
    pfem::DevBuf<unsigned char> t_cnt;
    pfem::DevBuf<long long> t_slice_off;
    pfem::DevBuf<char> scratch[32];        // persistent, grow-only set-up scratch (sort buffers, upload staging, every temporary of the
                                           // pattern pass): a cudaMalloc/cudaFree pair costs ~0.1 ms per MB when the driver really maps
                                           // and unmaps, which made set_pattern bimodal (33 vs 80-110 ms on C5) while its temporaries were locals
    // colour-scheduled tile value pass (default for one dof per node): assembly_ctile.cu / .cuh
    bool ct_ready = false, ct_tried = false, rows_ready = false, ct_full = true;
    int asm_mode_req = 0;                  // 0 auto (tile kernel when it applies), 1 row kernels (bit-identical to the sequential order)
    int ct_ntiles = 0, ct_TR = 0, ct_stride = 0, ct_node_cap = 0, ct_max_rounds = 0, ct_threads = 0;
    long long ct_visits = 0;
    size_t ct_smem = 0;
    pfem::DevBuf<int> ct_tdesc, ct_trow, ct_round_off;
    pfem::DevBuf<unsigned int> ct_vnode, ct_vslot;
    pfem::DevBuf<double> ct_tnode;

    // solver
    pfem::SellMatrix A;                    // diagonal block
    // off-diagonal block (ghost columns), CSR over the boundary rows only
    int n_ghost = 0, n_brows = 0;
    long long nnz_off = 0;
    pfem::DevBuf<int> brow_ids, brow_ptr, bcol;   // boundary row ids (local), ptr, compact ghost index
    pfem::DevBuf<int> off_ptr;                    // [size_local+1] per-row pointer into bcol/bval (persistent CG kernel)
    // PCBJACOBI/ILU(0): diagonal-block row ranges, factor values, inverted pivots, tagged solve vectors, row-ready epochs
    pfem::DevBuf<int> ilu_dlo, ilu_ddiag, ilu_dhi, ilu_order_l, ilu_order_u;   // order_*: rows sorted by dependency level
    long long pattern_seq = 0, ilu_sched_seq = -1;
    int ilu_levels_l = 0, ilu_levels_u = 0;
    pfem::DevBuf<double> ilu_fval, ilu_invd, ilu_y, ilu_z, ilu_ticket;
    pfem::DevBuf<unsigned int> ilu_ready;
    unsigned long long ilu_tickets = 0, ilu_tag = 0;
    unsigned int ilu_epoch = 0;
    pfem::DevBuf<double> pcg_parts;               // persistent kernel, lean barrier: replicated per-CTA partial records
    pfem::DevBuf<double> pcg_bcast;               // persistent kernel: locally broadcast reduction results + flag + push ticket
    pfem::DevBuf<double> bval;
    pfem::DevBuf<int> csr2sell;            // per CSR slot: destination (>=0 SELL entry, <0: -(offdiag entry)-1)
    std::vector<int> ghost_cols;           // global ids (sorted): PETSc garray
    // halo plan
    std::vector<int> send_counts, recv_counts, send_displs, recv_displs;
    pfem::DevBuf<int> send_idx;            // local row indices to pack, grouped by destination rank
    // slow-path adds: off-rank (row, col, value) triples per owner (4 ints each), shipped at the next assembly point
    std::vector<std::vector<int>> stash;
    long long off_pattern_total = 0, stash_received = 0;
    pfem::DevBuf<double> send_buf, ghost_buf;
    size_t ghost_tag_off = 0;                     // ghost_buf: offset (in doubles) of the tagged {value, tag} entries
    pfem::DevBuf<double *> send_dst_t;            // per packed halo entry: address of the peer's tagged ghost entry
    pfem::DevBuf<double> x, r, z, p, w, dinv, sv;   // sv: s = A z of the single-reduction CG variant
    pfem::DevBuf<double> partials;         // [4][max_blocks]
    pfem::DevBuf<pfem::CgState> cg;
    pfem::CgState *cg_host = nullptr;      // pinned mirror
    int its = 0, reason = 0;
    double rnorm = 0.0, t_assemble = 0.0, t_solve = 0.0;
    // optional per-launch timing of the SpMV inside the solve (bench roofline evidence)
    bool profile = false;
    std::vector<cudaEvent_t> prof_ev;
    double prof_spmv_s = 0.0;
    long long prof_spmv_n = 0;

    // communicator
    pfem::NcclApi *nccl = nullptr;
    void *comm = nullptr;
    // peer-memory path (nranks > 1, all ranks on one NVLink/NVSwitch box)
    bool p2p = false;
    pfem::P2pMail *mail = nullptr;                       // own exchange area (cudaMalloc, IPC-exported)
    pfem::P2pMail *peer_mail[pfem::P2P_MAX_RANKS] = {};
    double *peer_ghost[pfem::P2P_MAX_RANKS] = {};
    bool peer_mail_open = false;
    pfem::DevBuf<pfem::P2pCtx> p2p_ctx;
    pfem::DevBuf<double *> send_dst;                     // per packed halo entry: address inside the peer's ghost buffer
    unsigned long long solve_seq = 0;
};

namespace pfem {

// elements.cu
int element_ke_batch(int kind, int n, const double *x, const double *y, const double *z, const double *elemData,
                     const double *timeData, const double *valC, double *K, double *F, int *jac_neg);
// pattern.cu
int upload_mesh(pfem_solver *h, int kind, int nElem, const int *conn, int nNode, const double *coords,
                const int *node_map_get_old);
int build_pattern(pfem_solver *h, int nElem, int nsize, const int *elemDof, const int *nodeDof);
// assembly.cu
int plan_assembly(pfem_solver *h);
int assemble_values(pfem_solver *h, const double *elemData, const double *timeData, int *n_neg);
int add_entries(pfem_solver *h, int n, const int *rows, const int *cols, const double *vals, bool transposed,
                const double *F);
// assembly_ctile.cu
int build_ctiles(pfem_solver *h);
int assemble_values_ctile(pfem_solver *h, const double *dElemData, const double *dTimeData);
// pattern.cu: streams of the row kernels (built on first use)
int build_asm_streams(pfem_solver *h);
// assembly_tiled.cu
int build_tiles_device(pfem_solver *h, int mode);
int assemble_values_tiled(pfem_solver *h, const double *dElemData, const double *dTimeData, bool unit);
// cg.cu
int build_solver_structures(pfem_solver *h);
int cg_solve(pfem_solver *h);
int time_spmv(pfem_solver *h, int reps, double *seconds);
// comm.cu
int comm_unique_id(void *id128);
int comm_init(pfem_solver *h, const void *id128);
void comm_destroy(pfem_solver *h);
int comm_allgather_int(pfem_solver *h, int value, std::vector<int> &out);
int stash_flush(pfem_solver *h);
int comm_alltoallv_int(pfem_solver *h, const std::vector<int> &sendbuf, const std::vector<int> &sendcounts,
                       std::vector<int> &recvbuf, std::vector<int> &recvcounts);
int comm_halo_exchange(pfem_solver *h, const double *sendbuf, double *recvbuf, cudaStream_t s);
int comm_allreduce_sum(pfem_solver *h, double *buf, int n, cudaStream_t s);
int comm_allgatherv_double(pfem_solver *h, const double *local, double *global_dev, cudaStream_t s);
int comm_p2p_setup(pfem_solver *h);
void comm_p2p_teardown(pfem_solver *h, bool final);

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// typed view of one of the handle's persistent scratch buffers
template <typename T> inline int scratch_get(pfem_solver *h, int idx, size_t count, T **out)
{
    int st = h->scratch[idx].reserve(count * sizeof(T) + 256);
    if (st != PFEM_OK) return st;
    *out = reinterpret_cast<T *>(h->scratch[idx].p);
    return PFEM_OK;
}

// set-up temporary in slot `slot` of the handle's persistent scratch: same use as a local DevBuf<T>, no allocation per call
template <typename T> struct Tmp {
    T *p = nullptr;
    int alloc(pfem_solver *h, int slot, size_t count) { return scratch_get<T>(h, slot, count, &p); }
};

// PFEM_TRACE=1: print host wall time of the set-up stages (stderr)
struct StageTimer {
    const char *name; double t0; bool on;
    static double now();
    explicit StageTimer(const char *n);
    ~StageTimer();
};

}  // namespace pfem
