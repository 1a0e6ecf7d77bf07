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

#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cstring>

#include "amaru_internal.h"
#include "p2p.cuh"

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
    // peer-memory path
    bool p2p = false, p2p_ready = false;
    void *d_win = nullptr;                 // this rank's P2PWin
    void *peer_win[16] = {nullptr};        // every rank's window (own entry = d_win)
    std::vector<void *> opened;            // cudaIpcOpenMemHandle results to close
    double **d_peer_p = nullptr;           // [nneigh] neighbours' p vectors
    double **d_peer_x = nullptr;           // [nneigh] neighbours' x vectors
    bool direct = false;                   // peers are devices of this process (no NCCL communicator, no cudaIpc)
    unsigned long long timeout_ns = 10000000000ull;
    int64_t *d_peer_start = nullptr;       // [nneigh] where this rank's nodes start in the neighbour's numbering
    int64_t *d_send_ptr = nullptr;
    int *d_neigh = nullptr;
    unsigned int *d_counter = nullptr;
    // fused CG loop (p2p.cuh): per-owned-node table of the ghost slots the node's p entry is pushed into
    std::vector<int32_t> h_send_nodes;
    int32_t *d_bidx = nullptr, *d_bent_ptr = nullptr, *d_bent_q = nullptr;
    int64_t *d_bent_remote = nullptr;
    unsigned int *d_fcounter = nullptr;
    bool fused_ready = false;
};

__global__ void k_pack(int64_t n, int nd, const int32_t *__restrict__ nodes, const double *__restrict__ v, double *buf) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n * nd; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i / nd;
        const int d = (int)(i - k * nd);
        buf[i] = v[(int64_t)nodes[k] * nd + d];
    }
}


// ---- peer-memory path: the per-iteration exchanges of the CG loop without NCCL ----------------------------------------
// Every rank owns a small window {scalar slots, flags, epochs, abort flag} and lets its peers address it and its p / x
// vectors (cudaIpc handles between processes, plain peer access inside one process).  Two kernels replace the three NCCL
// calls of an iteration:
//   k_p2p_halo       stores the boundary entries of a vector straight into the neighbours' ghost slots over NVLink, then
//                    (last block) raises its epoch flag in each neighbour's window and waits for the neighbours' flags;
//   k_p2p_allreduce  one warp: lane r stores this rank's partial values into rank r's window (slot of this rank, buffer
//                    epoch&1) and raises the flag, then the warp waits for all ranks' flags in its own window and combines
//                    the slots in rank order — every rank forms bitwise the same sum / max.
// Epochs live in the window and are advanced by the kernels themselves, so a batch of CG iterations is a fixed launch
// sequence and can be replayed as a CUDA graph.  Safety of the reuse: a slot/flag of parity e&1 is rewritten at epoch e+2,
// which a rank can only reach after it finished epoch e+1, i.e. after every peer entered epoch e+1, i.e. after every peer
// finished reading epoch e.  Ghost entries are rewritten only after an all-reduce that every rank enters after the
// product that read them.  Every wait is bounded (timeout_ns of %globaltimer): a rank that gives up raises the abort flag
// in every window, all spinning kernels leave, later ones return at once, and the host reports AMARU_ERR_COMM.
// (window layout and wait / release primitives: p2p.cuh)
// op 0: sum, 1: max
__global__ void k_p2p_allreduce(P2PDev pd, double *vals, int n, int op) {
    const int lane = threadIdx.x;
    P2PWin *me = pd.win[pd.rank];
    if (*reinterpret_cast<volatile int *>(&me->abort)) return;
    const unsigned long long epoch = me->scal_epoch + 1;
    const int par = (int)(epoch & 1ull);
    if (lane < pd.nranks) {
        P2PWin *w = pd.win[lane];
        for (int k = 0; k < n; k++) w->slot[par][pd.rank][k] = vals[k];
        __threadfence_system();
        st_release_sys(&w->sflag[par][pd.rank], epoch);
    }
    bool ok = true;
    if (lane < pd.nranks) ok = p2p_wait(pd, &me->sflag[par][lane], epoch);
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        if (ok) {
            for (int k = 0; k < n; k++) {
                double s = *reinterpret_cast<volatile double *>(&me->slot[par][0][k]);
                for (int r = 1; r < pd.nranks; r++) {
                    const double v = *reinterpret_cast<volatile double *>(&me->slot[par][r][k]);
                    s = op == 0 ? s + v : fmax(s, v);
                }
                vals[k] = s;
            }
        }
        me->scal_epoch = epoch;
    }
}

__global__ void k_p2p_halo(P2PDev pd, int nneigh, const int *__restrict__ neigh, const int64_t *__restrict__ send_ptr,
                           const int32_t *__restrict__ send_nodes, double *const *__restrict__ peer_v,
                           const int64_t *__restrict__ peer_start, int nd, const double *__restrict__ v, unsigned int *counter, int wait) {
    __shared__ bool last;
    P2PWin *me = pd.win[pd.rank];
    if (*reinterpret_cast<volatile int *>(&me->abort)) return;
    const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long *>(&me->halo_epoch) + 1;
    const int64_t total = send_ptr[nneigh] * nd;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i / nd;
        const int d = (int)(i - k * nd);
        int q = 0;
        while (k >= send_ptr[q + 1]) q++;
        peer_v[q][(peer_start[q] + (k - send_ptr[q])) * nd + d] = v[(int64_t)send_nodes[k] * nd + d];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (last) {   // every block has read the epoch and fenced its remote stores before its counter increment
        if (threadIdx.x < nneigh) {
            __threadfence_system();
            st_release_sys(&pd.win[neigh[threadIdx.x]]->hflag[pd.rank], epoch);
            if (wait) p2p_wait(pd, &me->hflag[neigh[threadIdx.x]], epoch);
        }
        __syncthreads();
        if (threadIdx.x == 0 && wait) me->halo_epoch = epoch;   // push-only: the consuming kernel advances the epoch
    }
}

}  // namespace

static P2PDev make_pd(amaru_model *m, HaloComm *hc) {
    P2PDev pd;
    pd.rank = m->rank;
    pd.nranks = m->nranks;
    pd.timeout_ns = hc->timeout_ns;
    for (int r = 0; r < P2P_MAXR; r++) pd.win[r] = static_cast<P2PWin *>(r < m->nranks ? hc->peer_win[r] : nullptr);
    return pd;
}

// `uid` == nullptr: peers are devices of this process (amaru_create with ngpus > 1): no NCCL communicator is made, every
// exchange goes through peer memory once amaru_p2p_connect_direct has run
void amaru_comm_setup(amaru_model *m, int nneigh, const int32_t *neigh_rank, const int64_t *send_ptr,
                      const int32_t *send_nodes, const int64_t *recv_start, const int64_t *recv_count, const void *uid) {
    HaloComm *hc = new HaloComm();
    m->comm = hc;
    hc->nneigh = nneigh;
    hc->neigh.assign(neigh_rank, neigh_rank + nneigh);
    hc->send_ptr.assign(send_ptr, send_ptr + nneigh + 1);
    hc->recv_start.assign(recv_start, recv_start + nneigh);
    hc->recv_count.assign(recv_count, recv_count + nneigh);
    hc->nsend = send_ptr[nneigh];
    hc->h_send_nodes.assign(send_nodes, send_nodes + hc->nsend);
    if (const char *e = getenv("AMARU_P2P_TIMEOUT_MS")) hc->timeout_ns = (unsigned long long)std::max(1, atoi(e)) * 1000000ull;
    for (int q = 0; q < nneigh; q++) {
        AMARU_REQUIRE(neigh_rank[q] >= 0 && neigh_rank[q] < m->nranks && neigh_rank[q] != m->rank, AMARU_ERR_ARG, "bad neighbour rank");
        AMARU_REQUIRE(recv_start[q] >= m->nowned && recv_start[q] + recv_count[q] <= m->nnodes, AMARU_ERR_ARG, "bad ghost range");
    }
    for (int64_t i = 0; i < hc->nsend; i++)
        AMARU_REQUIRE(send_nodes[i] >= 0 && send_nodes[i] < m->nowned, AMARU_ERR_ARG, "send list must hold owned nodes");
    CUDA_CHECK(cudaMalloc(&hc->d_send_nodes, std::max<int64_t>(hc->nsend, 1) * sizeof(int32_t)));
    CUDA_CHECK(cudaMemcpy(hc->d_send_nodes, send_nodes, hc->nsend * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&hc->d_sendbuf, std::max<int64_t>(hc->nsend, 1) * m->nd * sizeof(double)));
    if (uid) {
        ncclUniqueId id;
        std::memcpy(&id, uid, sizeof(id));
        NCCL_CHECK(nccl().CommInitRank(&hc->comm, m->nranks, id, m->rank));
    } else {
        hc->direct = true;
    }
}

void amaru_comm_destroy(amaru_model *m) {
    HaloComm *hc = static_cast<HaloComm *>(m->comm);
    if (!hc) return;
    if (hc->comm) nccl().CommDestroy(hc->comm);
    for (void *p : hc->opened) cudaIpcCloseMemHandle(p);
    cudaFree(hc->d_win);
    cudaFree(hc->d_peer_p);
    cudaFree(hc->d_peer_x);
    cudaFree(hc->d_peer_start);
    cudaFree(hc->d_send_ptr);
    cudaFree(hc->d_neigh);
    cudaFree(hc->d_counter);
    cudaFree(hc->d_bidx);
    cudaFree(hc->d_bent_ptr);
    cudaFree(hc->d_bent_q);
    cudaFree(hc->d_bent_remote);
    cudaFree(hc->d_fcounter);
    cudaFree(hc->d_send_nodes);
    cudaFree(hc->d_sendbuf);
    delete hc;
    m->comm = nullptr;
}

bool amaru_comm_is_p2p(const amaru_model *m) {
    const HaloComm *hc = static_cast<const HaloComm *>(m->comm);
    return hc && hc->p2p;
}

// after a batch of peer-memory kernels: did any rank give up waiting?
void amaru_comm_check(amaru_model *m) {
    HaloComm *hc = static_cast<HaloComm *>(m->comm);
    if (!hc || !hc->p2p || !hc->d_win) return;
    int ab = 0;
    CUDA_CHECK(cudaMemcpyAsync(&ab, reinterpret_cast<char *>(hc->d_win) + offsetof(P2PWin, abort), sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    CUDA_CHECK(cudaStreamSynchronize(m->stream));
    if (ab) throw AmaruError{AMARU_ERR_COMM, "peer-memory exchange timed out: a rank left the collective sequence"};
}

// ghost entries of the node-major vector d_v <- owners' values
void amaru_halo_exchange(amaru_model *m, double *d_v) {
    if (m->nranks <= 1) return;
    HaloComm *hc = static_cast<HaloComm *>(m->comm);
    AMARU_REQUIRE(hc && (hc->comm || hc->direct), AMARU_ERR_COMM, "halo exchange: communicator not initialised");
    double **peer = d_v == m->d_p ? hc->d_peer_p : (d_v == m->d_x ? hc->d_peer_x : nullptr);
    if (hc->p2p && peer) {   // push into the neighbours' ghost slots, flag, wait (no NCCL)
        const int64_t n = hc->nsend * m->nd;
        const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)m->nsm * 2));
        k_p2p_halo<<<blocks, 256, 0, m->stream>>>(make_pd(m, hc), hc->nneigh, hc->d_neigh, hc->d_send_ptr, hc->d_send_nodes, peer,
                                                  hc->d_peer_start, m->nd, d_v, hc->d_counter, 1);
        m->launches++;
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    AMARU_REQUIRE(hc->comm != nullptr, AMARU_ERR_COMM, "halo exchange: vector has no peer mapping and there is no NCCL communicator");
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

static void small_allreduce(amaru_model *m, double *d_vals, int64_t n, int op) {
    HaloComm *hc = static_cast<HaloComm *>(m->comm);
    AMARU_REQUIRE(hc && (hc->comm || hc->direct), AMARU_ERR_COMM, "all-reduce: communicator not initialised");
    if (hc->p2p && n <= 4) {   // the scalar all-reduces of the CG loop through peer memory
        k_p2p_allreduce<<<1, 32, 0, m->stream>>>(make_pd(m, hc), d_vals, (int)n, op);
        m->launches++;
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    AMARU_REQUIRE(hc->comm != nullptr, AMARU_ERR_COMM, "all-reduce of a long vector needs the NCCL communicator");
    NCCL_CHECK(nccl().AllReduce(d_vals, d_vals, (size_t)n, ncclDouble, op == 0 ? ncclSum : ncclMax, hc->comm, m->stream));
}

void amaru_allreduce_sum(amaru_model *m, double *d_vals, int64_t n) {
    if (m->nranks <= 1) return;
    small_allreduce(m, d_vals, n, 0);
}

void amaru_allreduce_max(amaru_model *m, double *d_vals, int64_t n) {
    if (m->nranks <= 1) return;
    small_allreduce(m, d_vals, n, 1);
}

void amaru_allreduce_max_int(amaru_model *m, int *d_val) {
    if (m->nranks <= 1) return;
    HaloComm *hc = static_cast<HaloComm *>(m->comm);
    AMARU_REQUIRE(hc && (hc->comm || hc->direct), AMARU_ERR_COMM, "all-reduce: communicator not initialised");
    if (hc->direct) return;   // one process: the caller combines the ranks' status words on the host
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

static void ensure_window(amaru_model *m, HaloComm *hc) {
    if (!hc->d_win) {
        CUDA_CHECK(cudaMalloc(&hc->d_win, sizeof(P2PWin)));
        CUDA_CHECK(cudaMemset(hc->d_win, 0, sizeof(P2PWin)));
    }
}

// device-side tables of the push kernel: neighbours' vectors, where this rank's nodes start in their numbering
static void upload_peer_tables(amaru_model *m, HaloComm *hc, const std::vector<double *> &peer_p, const std::vector<double *> &peer_x,
                               const int64_t *peer_recv_start) {
    const size_t nn = (size_t)std::max(hc->nneigh, 1);
    CUDA_CHECK(cudaMalloc(&hc->d_peer_p, nn * sizeof(double *)));
    CUDA_CHECK(cudaMalloc(&hc->d_peer_start, nn * sizeof(int64_t)));
    CUDA_CHECK(cudaMalloc(&hc->d_send_ptr, (nn + 1) * sizeof(int64_t)));
    CUDA_CHECK(cudaMalloc(&hc->d_neigh, nn * sizeof(int)));
    CUDA_CHECK(cudaMalloc(&hc->d_counter, sizeof(unsigned int)));
    CUDA_CHECK(cudaMemset(hc->d_counter, 0, sizeof(unsigned int)));
    CUDA_CHECK(cudaMemcpy(hc->d_peer_p, peer_p.data(), hc->nneigh * sizeof(double *), cudaMemcpyHostToDevice));
    if (!peer_x.empty()) {
        CUDA_CHECK(cudaMalloc(&hc->d_peer_x, nn * sizeof(double *)));
        CUDA_CHECK(cudaMemcpy(hc->d_peer_x, peer_x.data(), hc->nneigh * sizeof(double *), cudaMemcpyHostToDevice));
    }
    if (hc->nneigh) CUDA_CHECK(cudaMemcpy(hc->d_peer_start, peer_recv_start, hc->nneigh * sizeof(int64_t), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(hc->d_send_ptr, hc->send_ptr.data(), (hc->nneigh + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    if (hc->nneigh) CUDA_CHECK(cudaMemcpy(hc->d_neigh, hc->neigh.data(), hc->nneigh * sizeof(int), cudaMemcpyHostToDevice));
    // boundary table of the fused loop: for every owned node on a send list, the (neighbour, remote node) pairs it goes to
    {
        std::vector<int32_t> bidx((size_t)std::max<int64_t>(m->nowned, 1), -1), cnt;
        std::vector<int32_t> order;                       // boundary nodes in first-seen order
        for (int q = 0; q < hc->nneigh; q++)
            for (int64_t k = hc->send_ptr[(size_t)q]; k < hc->send_ptr[(size_t)q + 1]; k++) {
                const int32_t n = hc->h_send_nodes[(size_t)k];
                if (bidx[(size_t)n] < 0) {
                    bidx[(size_t)n] = (int32_t)order.size();
                    order.push_back(n);
                    cnt.push_back(0);
                }
                cnt[(size_t)bidx[(size_t)n]]++;
            }
        std::vector<int32_t> ptr(order.size() + 1, 0);
        for (size_t i = 0; i < order.size(); i++) ptr[i + 1] = ptr[i] + cnt[i];
        std::vector<int32_t> bq((size_t)std::max<int64_t>(hc->nsend, 1), 0), fill(ptr.begin(), ptr.end() - 1);
        std::vector<int64_t> br((size_t)std::max<int64_t>(hc->nsend, 1), 0);
        for (int q = 0; q < hc->nneigh; q++)
            for (int64_t k = hc->send_ptr[(size_t)q]; k < hc->send_ptr[(size_t)q + 1]; k++) {
                const int32_t b = bidx[(size_t)hc->h_send_nodes[(size_t)k]];
                const int32_t at = fill[(size_t)b]++;
                bq[(size_t)at] = q;
                br[(size_t)at] = peer_recv_start[q] + (k - hc->send_ptr[(size_t)q]);
            }
        CUDA_CHECK(cudaMalloc(&hc->d_bidx, bidx.size() * sizeof(int32_t)));
        CUDA_CHECK(cudaMalloc(&hc->d_bent_ptr, ptr.size() * sizeof(int32_t)));
        CUDA_CHECK(cudaMalloc(&hc->d_bent_q, bq.size() * sizeof(int32_t)));
        CUDA_CHECK(cudaMalloc(&hc->d_bent_remote, br.size() * sizeof(int64_t)));
        CUDA_CHECK(cudaMalloc(&hc->d_fcounter, sizeof(unsigned int)));
        CUDA_CHECK(cudaMemset(hc->d_fcounter, 0, sizeof(unsigned int)));
        CUDA_CHECK(cudaMemcpy(hc->d_bidx, bidx.data(), bidx.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(hc->d_bent_ptr, ptr.data(), ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(hc->d_bent_q, bq.data(), bq.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(hc->d_bent_remote, br.data(), br.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
        hc->fused_ready = true;
    }
    CUDA_CHECK(cudaDeviceSynchronize());
    hc->p2p_ready = true;
}

// the CG loop may fold its exchanges into the producing / consuming kernels (AMARU_P2P_FUSED=0 keeps the stand-alone kernels)
bool amaru_comm_fused(const amaru_model *m) {
    const HaloComm *hc = static_cast<const HaloComm *>(m->comm);
    if (!hc || !hc->p2p || !hc->fused_ready) return false;
    const char *e = getenv("AMARU_P2P_FUSED");
    return !(e && atoi(e) == 0);
}

P2PFused amaru_comm_fused_args(amaru_model *m) {
    HaloComm *hc = static_cast<HaloComm *>(m->comm);
    P2PFused f;
    f.pd = make_pd(m, hc);
    f.nneigh = hc->nneigh;
    f.neigh = hc->d_neigh;
    f.peer_p = hc->d_peer_p;
    f.bidx = hc->d_bidx;
    f.bent_ptr = hc->d_bent_ptr;
    f.bent_q = hc->d_bent_q;
    f.bent_remote = hc->d_bent_remote;
    f.counter = hc->d_fcounter;
    return f;
}

// producer half of a halo exchange of p (the consumer is the operator kernel of the fused loop): push + flags, no wait
void amaru_halo_push(amaru_model *m, double *d_v) {
    HaloComm *hc = static_cast<HaloComm *>(m->comm);
    AMARU_REQUIRE(hc && hc->p2p && d_v == m->d_p, AMARU_ERR_COMM, "halo push: peer-memory path of p only");
    const int64_t n = hc->nsend * m->nd;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)m->nsm * 2));
    k_p2p_halo<<<blocks, 256, 0, m->stream>>>(make_pd(m, hc), hc->nneigh, hc->d_neigh, hc->d_send_ptr, hc->d_send_nodes, hc->d_peer_p,
                                              hc->d_peer_start, m->nd, d_v, hc->d_counter, 0);
    m->launches++;
    CUDA_CHECK(cudaGetLastError());
}

// ---- in-process variant (amaru_create with ngpus > 1): the parts are devices of this process with peer access enabled;
// called once, from one thread, after every part's amaru_comm_setup(uid = nullptr)
void amaru_p2p_connect_direct(amaru_model *const *parts, int n) {
    AMARU_REQUIRE(n > 1 && n <= P2P_MAXR, AMARU_ERR_ARG, "peer-memory path: 2..16 GPUs");
    for (int r = 0; r < n; r++) {
        CUDA_CHECK(cudaSetDevice(parts[r]->device));
        ensure_window(parts[r], static_cast<HaloComm *>(parts[r]->comm));
    }
    for (int r = 0; r < n; r++) {
        amaru_model *m = parts[r];
        HaloComm *hc = static_cast<HaloComm *>(m->comm);
        CUDA_CHECK(cudaSetDevice(m->device));
        for (int q = 0; q < n; q++) hc->peer_win[q] = static_cast<HaloComm *>(parts[q]->comm)->d_win;
        std::vector<double *> pp((size_t)std::max(hc->nneigh, 1), nullptr), px((size_t)std::max(hc->nneigh, 1), nullptr);
        std::vector<int64_t> pstart((size_t)std::max(hc->nneigh, 1), 0);
        for (int i = 0; i < hc->nneigh; i++) {
            amaru_model *mq = parts[hc->neigh[(size_t)i]];
            HaloComm *hq = static_cast<HaloComm *>(mq->comm);
            pp[(size_t)i] = mq->d_p;
            px[(size_t)i] = mq->d_x;
            const auto it = std::find(hq->neigh.begin(), hq->neigh.end(), r);
            AMARU_REQUIRE(it != hq->neigh.end(), AMARU_ERR_ARG, "halo lists are not symmetric");
            pstart[(size_t)i] = hq->recv_start[(size_t)(it - hq->neigh.begin())];
        }
        upload_peer_tables(m, hc, pp, px, pstart.data());
        hc->p2p = true;
    }
}

// ---- between processes: handle exchange (the host all-gathers the 128-byte records, e.g. with torch.distributed / MPI.jl)
extern "C" int amaru_p2p_export(amaru_model *m, void *out128, char *msg, int msglen) {
    try {
        if (msg && msglen > 0) msg[0] = 0;
        AMARU_REQUIRE(m && out128, AMARU_ERR_ARG, "amaru_p2p_export: null argument");
        HaloComm *hc = static_cast<HaloComm *>(m->comm);
        AMARU_REQUIRE(hc && m->nranks > 1 && m->nranks <= P2P_MAXR, AMARU_ERR_ARG, "amaru_p2p_export: needs a partitioned handle of at most 16 ranks");
        CUDA_CHECK(cudaSetDevice(m->device));
        ensure_window(m, hc);
        cudaIpcMemHandle_t h[2];
        CUDA_CHECK(cudaIpcGetMemHandle(&h[0], hc->d_win));
        CUDA_CHECK(cudaIpcGetMemHandle(&h[1], m->d_p));
        static_assert(sizeof(h) == 128, "two IPC handles");
        std::memcpy(out128, h, sizeof(h));
        CUDA_CHECK(cudaDeviceSynchronize());
        return AMARU_OK;
    } catch (const AmaruError &e) {
        if (msg && msglen > 0) snprintf(msg, (size_t)msglen, "%s", e.msg.c_str());
        return e.code;
    }
}

// all_handles: nranks records of 128 bytes in rank order; peer_recv_start[i]: first local node id, in neighbour
// neigh_rank[i]'s numbering, of the ghost range that neighbour keeps for this rank (its recv_start for this rank)
extern "C" int amaru_p2p_connect(amaru_model *m, const void *all_handles, const int64_t *peer_recv_start, char *msg, int msglen) {
    try {
        if (msg && msglen > 0) msg[0] = 0;
        AMARU_REQUIRE(m && all_handles, AMARU_ERR_ARG, "amaru_p2p_connect: null argument");
        HaloComm *hc = static_cast<HaloComm *>(m->comm);
        AMARU_REQUIRE(hc && hc->d_win && m->nranks > 1 && m->nranks <= P2P_MAXR, AMARU_ERR_ARG, "amaru_p2p_connect: call amaru_p2p_export first");
        AMARU_REQUIRE(hc->nneigh == 0 || peer_recv_start, AMARU_ERR_ARG, "amaru_p2p_connect: null peer_recv_start");
        CUDA_CHECK(cudaSetDevice(m->device));
        const cudaIpcMemHandle_t *H = static_cast<const cudaIpcMemHandle_t *>(all_handles);
        std::vector<double *> peer_p((size_t)std::max(hc->nneigh, 1), nullptr);
        for (int r = 0; r < m->nranks; r++) {
            if (r == m->rank) {
                hc->peer_win[r] = hc->d_win;
                continue;
            }
            void *w = nullptr;
            CUDA_CHECK(cudaIpcOpenMemHandle(&w, H[2 * r], cudaIpcMemLazyEnablePeerAccess));
            hc->opened.push_back(w);
            hc->peer_win[r] = w;
        }
        for (int q = 0; q < hc->nneigh; q++) {
            void *pp = nullptr;
            CUDA_CHECK(cudaIpcOpenMemHandle(&pp, H[2 * hc->neigh[q] + 1], cudaIpcMemLazyEnablePeerAccess));
            hc->opened.push_back(pp);
            peer_p[(size_t)q] = static_cast<double *>(pp);
        }
        upload_peer_tables(m, hc, peer_p, {}, peer_recv_start);   // x keeps the NCCL exchange between processes
        return AMARU_OK;
    } catch (const AmaruError &e) {
        if (msg && msglen > 0) snprintf(msg, (size_t)msglen, "%s", e.msg.c_str());
        return e.code;
    }
}

// Collective switch: the host calls it with on=1 on EVERY rank only after every rank's amaru_p2p_connect succeeded (a mixed
// state would leave some ranks waiting on flags while others sit in NCCL calls).
extern "C" int amaru_p2p_enable(amaru_model *m, int on) {
    if (!m || !m->comm) return AMARU_ERR_ARG;
    HaloComm *hc = static_cast<HaloComm *>(m->comm);
    if (on && !hc->p2p_ready) return AMARU_ERR_ARG;
    hc->p2p = on != 0;
    return AMARU_OK;
}
