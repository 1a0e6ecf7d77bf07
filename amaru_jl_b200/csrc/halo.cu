// K10: multi-GPU plumbing of the PCG (one process per GPU): halo exchange of the ghost entries of a node-major vector
// and small all-reduces, over NCCL (NVLink 5 / NVSwitch).  The reference has no distributed path.
//
// Layout contract (amaru_jl_b200/partition.py): local nodes = owned nodes first, then ghosts grouped by owner rank; for every
// neighbour q the ghosts owned by q form one contiguous range [recv_start, recv_start+recv_count) and q sends exactly
// those nodes in the same (ascending global id) order, so received data lands in place and only the send side packs.
// NCCL is resolved with dlopen("libnccl.so.2") when the first partitioned handle is created (single-GPU use never loads
// it; inside a torch process the already-loaded bundled NCCL is picked up, so all ranks run the same version).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "amaru_internal.h"

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi &nccl() {
    static NcclApi api;
    if (api.lib) return api;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw AmaruError{AMARU_ERR_COMM, std::string("cannot load libnccl.so.2: ") + dlerror()};
    auto sym = [&](const char *n) {
        void *p = dlsym(h, n);
        if (!p) throw AmaruError{AMARU_ERR_COMM, std::string("libnccl: missing symbol ") + n};
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.lib = h;
    return api;
}

#define NCCL_CHECK(call)                                                                                   \
    do {                                                                                                   \
        ncclResult_t r__ = (call);                                                                         \
        if (r__ != ncclSuccess)                                                                            \
            throw AmaruError{AMARU_ERR_COMM, std::string(#call) + ": " + nccl().GetErrorString(r__)};      \
    } while (0)

struct HaloComm {
    ncclComm_t comm = nullptr;
    int nneigh = 0;
    std::vector<int> neigh;
    std::vector<int64_t> send_ptr, recv_start, recv_count;   // in nodes
    int32_t *d_send_nodes = nullptr;
    double *d_sendbuf = nullptr;
    int64_t nsend = 0;
};

__global__ void k_pack(int64_t n, int nd, const int32_t *__restrict__ nodes, const double *__restrict__ v, double *buf) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n * nd; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i / nd;
        const int d = (int)(i - k * nd);
        buf[i] = v[(int64_t)nodes[k] * nd + d];
    }
}

}  // namespace

void amaru_comm_setup(amaru_model *m, int nneigh, const int32_t *neigh_rank, const int64_t *send_ptr,
                      const int32_t *send_nodes, const int64_t *recv_start, const int64_t *recv_count, const void *uid) {
    NcclApi &api = nccl();
    HaloComm *hc = new HaloComm();
    m->comm = hc;
    hc->nneigh = nneigh;
    hc->neigh.assign(neigh_rank, neigh_rank + nneigh);
    hc->send_ptr.assign(send_ptr, send_ptr + nneigh + 1);
    hc->recv_start.assign(recv_start, recv_start + nneigh);
    hc->recv_count.assign(recv_count, recv_count + nneigh);
    hc->nsend = send_ptr[nneigh];
    for (int q = 0; q < nneigh; q++) {
        AMARU_REQUIRE(neigh_rank[q] >= 0 && neigh_rank[q] < m->nranks && neigh_rank[q] != m->rank, AMARU_ERR_ARG, "bad neighbour rank");
        AMARU_REQUIRE(recv_start[q] >= m->nowned && recv_start[q] + recv_count[q] <= m->nnodes, AMARU_ERR_ARG, "bad ghost range");
    }
    for (int64_t i = 0; i < hc->nsend; i++)
        AMARU_REQUIRE(send_nodes[i] >= 0 && send_nodes[i] < m->nowned, AMARU_ERR_ARG, "send list must hold owned nodes");
    CUDA_CHECK(cudaMalloc(&hc->d_send_nodes, std::max<int64_t>(hc->nsend, 1) * sizeof(int32_t)));
    CUDA_CHECK(cudaMemcpy(hc->d_send_nodes, send_nodes, hc->nsend * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&hc->d_sendbuf, std::max<int64_t>(hc->nsend, 1) * m->nd * sizeof(double)));
    ncclUniqueId id;
    std::memcpy(&id, uid, sizeof(id));
    NCCL_CHECK(api.CommInitRank(&hc->comm, m->nranks, id, m->rank));
}

void amaru_comm_destroy(amaru_model *m) {
    HaloComm *hc = static_cast<HaloComm *>(m->comm);
    if (!hc) return;
    if (hc->comm) nccl().CommDestroy(hc->comm);
    cudaFree(hc->d_send_nodes);
    cudaFree(hc->d_sendbuf);
    delete hc;
    m->comm = nullptr;
}

// ghost entries of the node-major vector d_v <- owners' values
void amaru_halo_exchange(amaru_model *m, double *d_v) {
    if (m->nranks <= 1) return;
    HaloComm *hc = static_cast<HaloComm *>(m->comm);
    AMARU_REQUIRE(hc && hc->comm, AMARU_ERR_COMM, "halo exchange: communicator not initialised");
    NcclApi &api = nccl();
    if (hc->nsend > 0) {
        const int64_t n = hc->nsend * m->nd;
        const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)m->nsm * 8);
        k_pack<<<blocks, 256, 0, m->stream>>>(hc->nsend, m->nd, hc->d_send_nodes, d_v, hc->d_sendbuf);
        m->launches++;
    }
    NCCL_CHECK(api.GroupStart());
    for (int q = 0; q < hc->nneigh; q++) {
        const int64_t ns = (hc->send_ptr[q + 1] - hc->send_ptr[q]) * m->nd, nr = hc->recv_count[q] * m->nd;
        if (ns > 0) NCCL_CHECK(api.Send(hc->d_sendbuf + hc->send_ptr[q] * m->nd, (size_t)ns, ncclDouble, hc->neigh[q], hc->comm, m->stream));
        if (nr > 0) NCCL_CHECK(api.Recv(d_v + hc->recv_start[q] * m->nd, (size_t)nr, ncclDouble, hc->neigh[q], hc->comm, m->stream));
    }
    NCCL_CHECK(api.GroupEnd());
}

void amaru_allreduce_sum(amaru_model *m, double *d_vals, int64_t n) {
    if (m->nranks <= 1) return;
    HaloComm *hc = static_cast<HaloComm *>(m->comm);
    AMARU_REQUIRE(hc && hc->comm, AMARU_ERR_COMM, "all-reduce: communicator not initialised");
    NCCL_CHECK(nccl().AllReduce(d_vals, d_vals, (size_t)n, ncclDouble, ncclSum, hc->comm, m->stream));
}

void amaru_allreduce_max(amaru_model *m, double *d_vals, int64_t n) {
    if (m->nranks <= 1) return;
    HaloComm *hc = static_cast<HaloComm *>(m->comm);
    AMARU_REQUIRE(hc && hc->comm, AMARU_ERR_COMM, "all-reduce: communicator not initialised");
    NCCL_CHECK(nccl().AllReduce(d_vals, d_vals, (size_t)n, ncclDouble, ncclMax, hc->comm, m->stream));
}

void amaru_allreduce_max_int(amaru_model *m, int *d_val) {
    if (m->nranks <= 1) return;
    HaloComm *hc = static_cast<HaloComm *>(m->comm);
    AMARU_REQUIRE(hc && hc->comm, AMARU_ERR_COMM, "all-reduce: communicator not initialised");
    NCCL_CHECK(nccl().AllReduce(d_val, d_val, 1, ncclInt, ncclMax, hc->comm, m->stream));
}

extern "C" int amaru_nccl_unique_id(void *uid128, char *msg, int msglen) {
    try {
        if (msg && msglen > 0) msg[0] = 0;
        if (!uid128) return AMARU_ERR_ARG;
        ncclUniqueId id;
        NCCL_CHECK(nccl().GetUniqueId(&id));
        std::memcpy(uid128, &id, sizeof(id));
        return AMARU_OK;
    } catch (const AmaruError &e) {
        if (msg && msglen > 0) snprintf(msg, (size_t)msglen, "%s", e.msg.c_str());
        return e.code;
    }
}
