// comm.cu -- NCCL plumbing (one process per GPU, NVLink 5 / NVSwitch underneath).
//
// Replaces the MPI traffic PETSc generates inside KSPSolve for the reference:
// VecScatter ghost gather of MatMult_MPIAIJ  -> grouped ncclSend/ncclRecv halo exchange,
// MPI_Allreduce of VecDot/VecNorm            -> ncclAllReduce on a device-resident scalar block,
// VecScatterCreateToAll                      -> grouped ncclBroadcast (all-gather-v).
// NCCL is loaded lazily with dlopen so that single-GPU use has no NCCL dependency at all, and so that a
// host process which already carries an NCCL (e.g. PyTorch's) shares that one copy.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>

#include "internal.cuh"

namespace pfem {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *g_api = nullptr;

static NcclApi *load_nccl()
{
    if (g_api) return g_api;
    const char *names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    void *lib = nullptr;
    for (const char *n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) { set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return nullptr; }
    NcclApi *a = new NcclApi;
    a->lib = lib;
#define LOAD(field, sym)                                                        \
    *(void **)(&a->field) = dlsym(lib, sym);                                    \
    if (!a->field) { set_error("NCCL symbol %s missing", sym); delete a; return nullptr; }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(AllGather, "ncclAllGather")
    LOAD(Broadcast, "ncclBroadcast")
    LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    g_api = a;
    return a;
}

#define PFEM_NCCL(h, call)                                                                       \
    do {                                                                                         \
        ncclResult_t r_ = (call);                                                                \
        if (r_ != ncclSuccess) {                                                                 \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, (h)->nccl->GetErrorString(r_)); \
            return PFEM_ERR_NCCL;                                                                \
        }                                                                                        \
    } while (0)

int comm_unique_id(void *id128)
{
    NcclApi *a = load_nccl();
    if (!a) return PFEM_ERR_NCCL;
    ncclUniqueId id;
    ncclResult_t r = a->GetUniqueId(&id);
    if (r != ncclSuccess) { set_error("ncclGetUniqueId: %s", a->GetErrorString(r)); return PFEM_ERR_NCCL; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return PFEM_OK;
}

int comm_init(pfem_solver *h, const void *id128)
{
    if (h->nranks == 1) return PFEM_OK;
    if (!id128) { set_error("pfem_solver_create: nranks > 1 needs the NCCL unique id"); return PFEM_ERR_ARG; }
    h->nccl = load_nccl();
    if (!h->nccl) return PFEM_ERR_NCCL;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    PFEM_NCCL(h, h->nccl->CommInitRank(&comm, h->nranks, id, h->rank));
    h->comm = comm;
    return PFEM_OK;
}

void comm_destroy(pfem_solver *h)
{
    if (h->comm && h->nccl) h->nccl->CommDestroy((ncclComm_t)h->comm);
    h->comm = nullptr;
}

int comm_allgather_int(pfem_solver *h, int value, std::vector<int> &out)
{
    out.assign(h->nranks, value);
    if (h->nranks == 1) return PFEM_OK;
    DevBuf<int> d;
    PFEM_TRY(d.alloc(h->nranks));
    PFEM_CUDA(cudaMemcpyAsync(d.p + h->rank, &value, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    PFEM_NCCL(h, h->nccl->AllGather(d.p + h->rank, d.p, 1, ncclInt32, (ncclComm_t)h->comm, h->stream));
    PFEM_CUDA(cudaMemcpyAsync(out.data(), d.p, h->nranks * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    PFEM_CUDA(cudaStreamSynchronize(h->stream));
    return PFEM_OK;
}

// sendbuf is grouped by destination rank (sendcounts[q] entries for rank q, in rank order)
int comm_alltoallv_int(pfem_solver *h, const std::vector<int> &sendbuf, const std::vector<int> &sendcounts,
                       std::vector<int> &recvbuf, std::vector<int> &recvcounts)
{
    const int P = h->nranks;
    recvcounts.assign(P, 0);
    recvbuf.clear();
    if (P == 1) return PFEM_OK;
    ncclComm_t comm = (ncclComm_t)h->comm;
    cudaStream_t s = h->stream;
    // counts matrix: row q = what rank q sends to everybody
    DevBuf<int> dcounts;
    PFEM_TRY(dcounts.alloc((size_t)P * P));
    PFEM_CUDA(cudaMemcpyAsync(dcounts.p + (size_t)h->rank * P, sendcounts.data(), P * sizeof(int), cudaMemcpyHostToDevice, s));
    PFEM_NCCL(h, h->nccl->AllGather(dcounts.p + (size_t)h->rank * P, dcounts.p, P, ncclInt32, comm, s));
    std::vector<int> all((size_t)P * P);
    PFEM_CUDA(cudaMemcpyAsync(all.data(), dcounts.p, (size_t)P * P * sizeof(int), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    size_t nrecv = 0, nsend = 0;
    for (int q = 0; q < P; q++) { recvcounts[q] = all[(size_t)q * P + h->rank]; nrecv += recvcounts[q]; nsend += sendcounts[q]; }
    DevBuf<int> dsend, drecv;
    PFEM_TRY(dsend.alloc(nsend + 1));
    PFEM_TRY(drecv.alloc(nrecv + 1));
    if (nsend) PFEM_CUDA(cudaMemcpyAsync(dsend.p, sendbuf.data(), nsend * sizeof(int), cudaMemcpyHostToDevice, s));
    PFEM_NCCL(h, h->nccl->GroupStart());
    size_t so = 0, ro = 0;
    for (int q = 0; q < P; q++) {
        if (q != h->rank) {
            if (sendcounts[q]) PFEM_NCCL(h, h->nccl->Send(dsend.p + so, sendcounts[q], ncclInt32, q, comm, s));
            if (recvcounts[q]) PFEM_NCCL(h, h->nccl->Recv(drecv.p + ro, recvcounts[q], ncclInt32, q, comm, s));
        }
        so += sendcounts[q];
        ro += recvcounts[q];
    }
    PFEM_NCCL(h, h->nccl->GroupEnd());
    recvbuf.resize(nrecv);
    if (nrecv) PFEM_CUDA(cudaMemcpyAsync(recvbuf.data(), drecv.p, nrecv * sizeof(int), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    return PFEM_OK;
}

int comm_halo_exchange(pfem_solver *h, const double *sendbuf, double *recvbuf, cudaStream_t s)
{
    const int P = h->nranks;
    ncclComm_t comm = (ncclComm_t)h->comm;
    PFEM_NCCL(h, h->nccl->GroupStart());
    for (int q = 0; q < P; q++) {
        if (q == h->rank) continue;
        if (h->send_counts[q]) PFEM_NCCL(h, h->nccl->Send(sendbuf + h->send_displs[q], h->send_counts[q], ncclFloat64, q, comm, s));
        if (h->recv_counts[q]) PFEM_NCCL(h, h->nccl->Recv(recvbuf + h->recv_displs[q], h->recv_counts[q], ncclFloat64, q, comm, s));
    }
    PFEM_NCCL(h, h->nccl->GroupEnd());
    return PFEM_OK;
}

int comm_allreduce_sum(pfem_solver *h, double *buf, int n, cudaStream_t s)
{
    if (h->nranks == 1) return PFEM_OK;
    PFEM_NCCL(h, h->nccl->AllReduce(buf, buf, n, ncclFloat64, ncclSum, (ncclComm_t)h->comm, s));
    return PFEM_OK;
}

int comm_allgatherv_double(pfem_solver *h, const double *local, double *global_dev, cudaStream_t s)
{
    const int P = h->nranks;
    if (P == 1) {
        PFEM_CUDA(cudaMemcpyAsync(global_dev, local, (size_t)h->size_local * sizeof(double), cudaMemcpyDeviceToDevice, s));
        return PFEM_OK;
    }
    ncclComm_t comm = (ncclComm_t)h->comm;
    PFEM_NCCL(h, h->nccl->GroupStart());
    for (int q = 0; q < P; q++) {
        const int cnt = h->row_starts[q + 1] - h->row_starts[q];
        if (cnt == 0) continue;
        PFEM_NCCL(h, h->nccl->Broadcast(local, global_dev + h->row_starts[q], cnt, ncclFloat64, q, comm, s));
    }
    PFEM_NCCL(h, h->nccl->GroupEnd());
    return PFEM_OK;
}

// ---- peer-memory (CUDA IPC over NVLink) set-up ------------------------------------------------------------------------
// Every rank exports its exchange area and its ghost buffer; the handles travel through one NCCL all-gather; each rank
// maps the peers' buffers and precomputes, for every packed halo entry, the address inside the owner's ghost buffer.
// Any failure (no peer access, IPC unavailable) leaves h->p2p false on ALL ranks and the NCCL path is used.

void comm_p2p_teardown(pfem_solver *h, bool final)
{
    for (int q = 0; q < h->nranks && q < P2P_MAX_RANKS; q++) {
        if (q == h->rank) continue;
        if (h->peer_ghost[q]) cudaIpcCloseMemHandle(h->peer_ghost[q]);
        h->peer_ghost[q] = nullptr;
        if (final && h->peer_mail[q]) { cudaIpcCloseMemHandle(h->peer_mail[q]); h->peer_mail[q] = nullptr; }
    }
    if (final) {
        h->peer_mail_open = false;
        if (h->mail) cudaFree(h->mail);
        h->mail = nullptr;
    }
    h->p2p = false;
}

int comm_p2p_setup(pfem_solver *h)
{
    const int P = h->nranks, me = h->rank;
    h->p2p = false;
    if (P == 1 || P > P2P_MAX_RANKS) return PFEM_OK;
    const char *env = getenv("PFEM_COMM");
    const bool want = !(env && strcmp(env, "nccl") == 0);
    cudaStream_t s = h->stream;
    ncclComm_t comm = (ncclComm_t)h->comm;
    int ok = want ? 1 : 0;
    if (ok && !h->mail) {
        if (cudaMalloc((void **)&h->mail, sizeof(P2pMail)) != cudaSuccess) { cudaGetLastError(); ok = 0; h->mail = nullptr; }
        else cudaMemset(h->mail, 0, sizeof(P2pMail));
    }
    // handles: [mail | ghost] per rank
    struct Pack { cudaIpcMemHandle_t mail, ghost; int ok; int pad[3]; };
    Pack mine;
    memset(&mine, 0, sizeof mine);
    if (ok && cudaIpcGetMemHandle(&mine.mail, h->mail) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    if (ok && cudaIpcGetMemHandle(&mine.ghost, h->ghost_buf.p) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    mine.ok = ok;
    DevBuf<Pack> dall;
    PFEM_TRY(dall.alloc(P));
    PFEM_CUDA(cudaMemcpyAsync(dall.p + me, &mine, sizeof(Pack), cudaMemcpyHostToDevice, s));
    PFEM_NCCL(h, h->nccl->AllGather(dall.p + me, dall.p, sizeof(Pack), ncclInt8, comm, s));
    std::vector<Pack> all(P);
    PFEM_CUDA(cudaMemcpyAsync(all.data(), dall.p, sizeof(Pack) * P, cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    for (int q = 0; q < P; q++) ok = ok && all[q].ok;
    if (ok) {
        for (int q = 0; q < P && ok; q++) {
            if (q == me) { h->peer_mail[q] = h->mail; h->peer_ghost[q] = h->ghost_buf.p; continue; }
            if (!h->peer_mail[q]) {
                void *ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, all[q].mail, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
                h->peer_mail[q] = (P2pMail *)ptr;
            }
            void *gp = nullptr;
            if (cudaIpcOpenMemHandle(&gp, all[q].ghost, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            h->peer_ghost[q] = (double *)gp;
        }
    }
    // every rank must take the same path
    std::vector<int> oks;
    PFEM_TRY(comm_allgather_int(h, ok, oks));
    for (int v : oks) ok = ok && v;
    if (!ok) {
        for (int q = 0; q < P; q++) {
            if (q != me && h->peer_ghost[q]) cudaIpcCloseMemHandle(h->peer_ghost[q]);
            h->peer_ghost[q] = nullptr;
        }
        return PFEM_OK;
    }
    // where do my values land in peer q's ghost buffer?  q's recv_displs[me]
    std::vector<int> send(P), ones(P, 1), recv, rc;
    for (int q = 0; q < P; q++) send[q] = h->recv_displs[q];
    PFEM_TRY(comm_alltoallv_int(h, send, ones, recv, rc));      // self entry is skipped by the exchange
    std::vector<int> their_displ(P, 0);
    {
        // comm_alltoallv_int leaves the self slot untouched in the packed receive buffer: rebuild by rank
        size_t o = 0;
        for (int q = 0; q < P; q++) { their_displ[q] = (q == me) ? 0 : recv[o]; o += 1; }
    }
    const int n_send = h->send_displs[P];
    std::vector<double *> dst((size_t)n_send + 1, nullptr);
    for (int q = 0; q < P; q++)
        for (int i = h->send_displs[q]; i < h->send_displs[q + 1]; i++)
            dst[i] = h->peer_ghost[q] + their_displ[q] + (i - h->send_displs[q]);
    PFEM_TRY(h->send_dst.alloc((size_t)n_send + 1));
    PFEM_CUDA(cudaMemcpy(h->send_dst.p, dst.data(), ((size_t)n_send + 1) * sizeof(double *), cudaMemcpyHostToDevice));
    {
        // the same for the tagged {value, tag} entries: they start ghost_tag_off doubles into the peer's buffer, and that
        // offset depends on the peer's ghost count
        std::vector<int> offs;
        PFEM_TRY(comm_allgather_int(h, (int)h->ghost_tag_off, offs));
        for (int q = 0; q < P; q++)
            for (int i = h->send_displs[q]; i < h->send_displs[q + 1]; i++)
                dst[i] = h->peer_ghost[q] + (size_t)offs[q] + 2 * ((size_t)their_displ[q] + (size_t)(i - h->send_displs[q]));
        PFEM_TRY(h->send_dst_t.alloc((size_t)n_send + 1));
        PFEM_CUDA(cudaMemcpy(h->send_dst_t.p, dst.data(), ((size_t)n_send + 1) * sizeof(double *), cudaMemcpyHostToDevice));
    }
    P2pCtx ctx;
    memset(&ctx, 0, sizeof ctx);
    ctx.rank = me; ctx.nranks = P; ctx.seq = 0;
    for (int q = 0; q < P; q++) {
        ctx.mail[q] = h->peer_mail[q];
        ctx.sends_to[q] = (q != me && h->send_counts[q] > 0) ? 1 : 0;
        ctx.recvs_from[q] = (q != me && h->recv_counts[q] > 0) ? 1 : 0;
    }
    PFEM_TRY(h->p2p_ctx.alloc(1));
    PFEM_CUDA(cudaMemcpy(h->p2p_ctx.p, &ctx, sizeof ctx, cudaMemcpyHostToDevice));
    h->p2p = true;
    return PFEM_OK;
}

}  // namespace pfem
