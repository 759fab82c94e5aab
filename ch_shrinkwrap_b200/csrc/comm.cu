// comm.cu -- multi-GPU plumbing: points are sharded across ranks, the mesh is replicated, and one
// iteration exchanges (1) the fixed-point vertex accumulators [AH res | AH 1] = 4M int64 and (2) the
// 11 Gram scalars.  Integer summation makes (1) bitwise independent of the number of ranks.
// NCCL is loaded lazily with dlopen so that single-GPU use has no NCCL dependency at all.
#include <dlfcn.h>
#include <cstring>
#include "common.cuh"

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1, ncclInt64 = 4, ncclUint64 = 5, ncclFloat64 = 8 };
enum { ncclSum = 0 };
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool load_nccl(std::string &err) {
    if (g_nccl.lib) return true;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(lib, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(lib, "ncclAllReduce");
    g_nccl.Broadcast = (decltype(g_nccl.Broadcast))dlsym(lib, "ncclBroadcast");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.Broadcast) {
        err = "libnccl.so.2 lacks required symbols";
        return false;
    }
    g_nccl.lib = lib;
    return true;
}

// copies the 11 Gram sums into a contiguous staging area and back
__global__ void k_pack_scalars(SolverState *st, double *buf, int dir) {
    if (threadIdx.x != 0) return;
    if (dir == 0) {
        for (int k = 0; k < 6; ++k) buf[k] = st->hc[k];
        for (int k = 0; k < 3; ++k) buf[6 + k] = st->gc[k];
        buf[9] = st->c0; buf[10] = st->res2; buf[11] = (double)st->nan_flag;
    } else {
        for (int k = 0; k < 6; ++k) st->hc[k] = buf[k];
        for (int k = 0; k < 3; ++k) st->gc[k] = buf[6 + k];
        st->c0 = buf[9]; st->res2 = buf[10];
        if (buf[11] != 0.0 && st->nan_flag == 0) st->nan_flag = 1;
    }
}
}  // namespace

#define NW_NCCL(call)                                                                          \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != 0) {                                                                         \
            h->err = std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error"); \
            return NW_ERR_COMM;                                                                \
        }                                                                                      \
    } while (0)

extern "C" int nw_comm_unique_id(char id[128]) {
    std::string err;
    if (!load_nccl(err)) return NW_ERR_COMM;
    ncclUniqueId u;
    if (g_nccl.GetUniqueId(&u) != 0) return NW_ERR_COMM;
    memcpy(id, u.internal, 128);
    return NW_OK;
}

extern "C" int nw_comm_init(nw_ctx *h, int rank, int nranks, const char id[128]) {
    if (!h) return NW_ERR_ARG;
    NW_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "nw_comm_init: bad rank / nranks");
    h->rank = rank; h->nranks = nranks;
    if (nranks == 1) return NW_OK;
    if (!load_nccl(h->err)) return NW_ERR_COMM;
    NW_CUDA(cudaSetDevice(h->device));
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    ncclComm_t c = nullptr;
    NW_NCCL(g_nccl.CommInitRank(&c, nranks, u, rank));
    h->nccl = c;
    return NW_OK;
}

int nw_comm_destroy(nw_ctx *h) {
    if (h->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)h->nccl);
    h->nccl = nullptr;
    return NW_OK;
}

int nw_comm_allreduce_host_doubles(nw_ctx *h, double *vals, int n);

int nw_allreduce_acc(nw_ctx *h) {
    if (h->nranks <= 1) return NW_OK;
    NW_ARG(h->nccl, "communicator not initialised");
    NW_NCCL(g_nccl.AllReduce(h->acc, h->acc, (size_t)4 * h->M, ncclInt64, ncclSum, (ncclComm_t)h->nccl, h->stream));
    return NW_OK;
}

// The mesh is replicated: rank 0 uploads it from its host copy and the other ranks receive it over NVLink instead of
// all of them pulling the same 84 MB through the host at once (measured on 8 GPUs: the ranks left the upload several ms
// apart, and the first accumulator allreduce of every block waited for the last one -- 1.15 ms per iteration on average
// for a collective that takes 0.09 ms).  It also makes rank 0's mesh authoritative, so the replicas cannot differ.
int nw_upload_replicated(nw_ctx *h, void *dst, const void *src, size_t bytes, size_t stride) {
    if (h->nranks <= 1) return nw_h2d_strided32(h, dst, src, bytes, stride);
    NW_ARG(h->nccl, "communicator not initialised");
    // rank 0 enters the collective even if its own copy failed (the others are already waiting in it); the failure is
    // agreed on by nw_comm_agree at the end of the upload, so every rank returns an error instead of some hanging
    int rc = NW_OK;
    if (h->rank == 0) rc = nw_h2d_strided32(h, dst, src, bytes, stride);
    NW_NCCL(g_nccl.Broadcast(dst, dst, bytes, ncclUint8, 0, (ncclComm_t)h->nccl, h->stream));
    if (rc != NW_OK) h->comm_status = rc;
    return NW_OK;
}

// Largest error code any rank has latched since the last call (0 = all fine), with a host synchronisation.  Ranks that
// were fine themselves get NW_ERR_COMM and a message saying a peer failed.
int nw_comm_agree(nw_ctx *h) {
    if (h->nranks <= 1) { const int rc = h->comm_status; h->comm_status = NW_OK; return rc; }
    NW_ARG(h->nccl, "communicator not initialised");
    double v[8] = {0};
    v[h->comm_status < 8 ? h->comm_status : 7] = 1.0;
    const int mine = h->comm_status;
    h->comm_status = NW_OK;
    NW_CHECK(nw_comm_allreduce_host_doubles(h, v, 8));
    int worst = 0;
    for (int k = 1; k < 8; ++k) if (v[k] > 0.0) worst = k;
    if (worst == 0) return NW_OK;
    if (mine == NW_OK) { h->err = "a peer rank failed during a collective call (see its error)"; return NW_ERR_COMM; }
    return mine;
}
// every rank must be talking about the same mesh: compares a few integers with rank 0's
int nw_check_replicated(nw_ctx *h, const long long *vals, int n, const char *what) {
    if (h->nranks <= 1) return NW_OK;
    NW_ARG(h->nccl && n <= 8, "communicator not initialised");
    long long *d = (long long *)(h->partials ? (void *)h->partials : nullptr), host[8];
    long long *tmp = nullptr;
    if (!d) { NW_CUDA(cudaMalloc((void **)&tmp, sizeof(long long) * 8)); d = tmp; }
    NW_CUDA(cudaMemcpyAsync(d, vals, sizeof(long long) * n, cudaMemcpyHostToDevice, h->stream));
    NW_NCCL(g_nccl.Broadcast(d, d, sizeof(long long) * n, ncclUint8, 0, (ncclComm_t)h->nccl, h->stream));
    NW_CUDA(cudaMemcpyAsync(host, d, sizeof(long long) * n, cudaMemcpyDeviceToHost, h->stream));
    NW_CUDA(cudaStreamSynchronize(h->stream));
    if (tmp) cudaFree(tmp);
    for (int k = 0; k < n; ++k)
        if (host[k] != vals[k]) { h->err = std::string(what) + ": differs from rank 0's (the mesh must be the same on every rank)"; return NW_ERR_ARG; }
    return NW_OK;
}

int nw_allreduce_scalars(nw_ctx *h) {
    if (h->nranks <= 1) return NW_OK;
    NW_ARG(h->nccl, "communicator not initialised");
    double *buf = h->partials + (size_t)h->n_partials * 16 + (size_t)nw_grid(h->M, 256) * 16;   // 64-double tail
    k_pack_scalars<<<1, 32, 0, h->stream>>>(h->st, buf, 0);
    NW_LAUNCH_CHECK();
    NW_NCCL(g_nccl.AllReduce(buf, buf, 12, ncclFloat64, ncclSum, (ncclComm_t)h->nccl, h->stream));
    k_pack_scalars<<<1, 32, 0, h->stream>>>(h->st, buf, 1);
    NW_LAUNCH_CHECK();
    return NW_OK;
}

// small host-side sums at setup time (weight mean, point counts)
int nw_comm_allreduce_host_doubles(nw_ctx *h, double *vals, int n) {
    if (h->nranks <= 1) return NW_OK;
    NW_ARG(h->nccl, "communicator not initialised");
    double *d = nullptr;
    NW_CHECK(nw_alloc(h, &d, (size_t)n));
    NW_CUDA(cudaMemcpyAsync(d, vals, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    NW_NCCL(g_nccl.AllReduce(d, d, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)h->nccl, h->stream));
    NW_CUDA(cudaMemcpyAsync(vals, d, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    NW_CUDA(cudaStreamSynchronize(h->stream));
    nw_free(&d);
    return NW_OK;
}
