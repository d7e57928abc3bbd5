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

}  // namespace pfem
