// Multi-GPU handles of ONE process: amaru_create(..., ngpus > 1) — the drop-in for solve!(ana) (reference
// src/mech/mech-solver.jl:172-181 is a single process) on N B200s of one box.  The reference has no distributed path; this
// is SURVEY.md §8(b)'s `ngpus` argument.
//
// The handle is a group of per-GPU parts (the same rank-level model the one-process-per-GPU launcher uses), each driven by
// its own host thread.  The mesh is partitioned inside the library (partition.cpp: RCB or METIS k-way, node ownership,
// duplicated halo elements); peers are addressed through plain peer access (cudaDeviceEnablePeerAccess), so
//   * the CG loop's halo push and scalar all-reduces are the peer-memory kernels of halo.cu (no NCCL anywhere), replayed
//     as one CUDA graph per GPU and batch;
//   * ABI vectors exist ONCE, on the first GPU: every rank gathers its inputs from them and stores the entries of the rows
//     it owns into them over NVLink (no global-length all-reduce, one H2D / D2H per call);
//   * status words are combined on the host.
// IP state in / out is scattered / gathered on the host by the rank threads (owner's copy wins).
#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <memory>
#include <functional>
#include <mutex>
#include <thread>

#include "partition.h"
#include "reduce.cuh"

struct GroupBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int n = 0, count = 0;
    uint64_t gen = 0;
    bool poisoned = false;
    bool wait() {
        std::unique_lock<std::mutex> lk(mu);
        if (poisoned) return false;
        const uint64_t g = gen;
        if (++count == n) {
            count = 0;
            gen++;
            cv.notify_all();
            return true;
        }
        cv.wait(lk, [&] { return gen != g || poisoned; });
        return !poisoned;
    }
    void poison() {
        std::lock_guard<std::mutex> lk(mu);
        poisoned = true;
        cv.notify_all();
    }
    void reset() {
        std::lock_guard<std::mutex> lk(mu);
        poisoned = false;
        count = 0;
    }
};

struct AmaruGroup {
    int n = 0;
    std::vector<int> dev;
    std::vector<amaru_model *> part;
    std::vector<AmaruLocalView> view;
    // layout of the caller's (global) model
    int nbatches = 0;
    std::vector<int> nn, nip;
    std::vector<int64_t> nelem, elem_off, ip_off;        // global, per batch
    std::vector<std::vector<int64_t>> lip_off;           // [part][batch] first IP of the batch in the part's ABI order
    GroupBarrier bar;
    // worker pool: one host thread per GPU
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    uint64_t job_gen = 0;
    int pending = 0;
    bool quit = false;
    std::function<int(int)> job;
    std::vector<int> status;
    std::vector<std::string> msg;
};

void amaru_group_barrier(amaru_model *part) {
    CUDA_CHECK(cudaStreamSynchronize(part->stream));
    if (!part->grp_of->bar.wait()) throw AmaruError{AMARU_ERR_COMM, "a peer GPU of the group failed"};
}

namespace {

void worker(AmaruGroup *g, int r) {
    cudaSetDevice(g->dev[(size_t)r]);
    uint64_t seen = 0;
    for (;;) {
        std::function<int(int)> job;
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_job.wait(lk, [&] { return g->quit || g->job_gen != seen; });
            if (g->quit) return;
            seen = g->job_gen;
            job = g->job;
        }
        int st = AMARU_OK;
        std::string msg;
        try {
            st = job(r);
        } catch (const AmaruError &e) {
            st = e.code;
            msg = e.msg;
            if (st < 0) g->bar.poison();   // peers blocked in a host barrier leave with AMARU_ERR_COMM
        } catch (const std::exception &e) {
            st = AMARU_ERR_ARG;
            msg = e.what();
            g->bar.poison();
        }
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->status[(size_t)r] = st;
            g->msg[(size_t)r] = msg;
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}

// run job(rank) on every worker; returns the combined status: the most severe error (< 0) if any, else the largest
// expected-failure code; *msg = message of the rank that produced it
int run(AmaruGroup *g, std::function<int(int)> job, std::string *msg = nullptr) {
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->job = std::move(job);
        g->pending = g->n;
        g->job_gen++;
    }
    g->cv_job.notify_all();
    {
        std::unique_lock<std::mutex> lk(g->mu);
        g->cv_done.wait(lk, [&] { return g->pending == 0; });
    }
    g->bar.reset();
    int st = AMARU_OK, who = -1;
    for (int r = 0; r < g->n; r++) {
        const int s = g->status[(size_t)r];
        // a rank's own failure outranks the AMARU_ERR_COMM its peers report for having been released from a barrier
        const bool worse = s != AMARU_OK && (st == AMARU_OK || (st == AMARU_ERR_COMM && s != AMARU_ERR_COMM) ||
                                             (st > 0 && (s < 0 || s > st)));
        if (worse) {
            st = s;
            who = r;
        }
    }
    if (msg) *msg = who >= 0 ? g->msg[(size_t)who] : std::string();
    return st;
}

void set_msg(char *msg, int msglen, const std::string &s) {
    if (msg && msglen > 0) std::snprintf(msg, (size_t)msglen, "%s", s.c_str());
}

int finish(int st, const std::string &m, char *msg, int msglen) {
    set_msg(msg, msglen, st == AMARU_OK ? std::string() : (m.empty() ? std::string(amaru_status_text(st)) : m));
    return st;
}

template <class T>
T *upload(const T *h, size_t n) {
    T *d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) CUDA_CHECK(cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

// Output side of a multi-GPU handle: load integration (loads.cu) and nodal patch recovery (recovery.cu) are surface- /
// output-sized work that wants the GLOBAL mesh, so the wrapper handle keeps, on the group's first GPU, exactly what those
// two files read from a model: global coordinates and eq ids, the caller's connectivity in ABI element order (identity
// permutation), materials, an ABI-order staging buffer; the IP state planes are allocated and refreshed from the parts
// (owner's copy of every element) when amaru_recover_nodal is called.
void output_side_setup(amaru_model *h, const CreateArgs &a) {
    CUDA_CHECK(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, h->device));
    h->nsm = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->d_coords = upload(a.coords, (size_t)a.nnodes * 3);
    h->d_eqid = upload(a.eqid, (size_t)a.nnodes * a.ndim);
    h->d_mat_kind = upload(a.mat_kind, (size_t)a.nmats);
    h->io_len = std::max<int64_t>(h->ndofs, 6 * h->nip_total);
    CUDA_CHECK(cudaMalloc(&h->d_io, (size_t)h->io_len * sizeof(double)));
    h->batches.resize((size_t)a.nbatches);
    int64_t eoff = 0, coff = 0, ipoff = 0;
    for (int b = 0; b < a.nbatches; b++) {
        ShapeInfo si;
        amaru_shape_info(a.batch_shape[b], si);
        Batch &B = h->batches[(size_t)b];
        B.shape = si.id; B.nn = si.nn; B.nd = si.nd; B.nip = si.nip;
        B.nelem = a.batch_nelem[b];
        B.elem_off = eoff;
        B.ip_off = ipoff;
        B.d_conn = upload(a.conn + coff, (size_t)B.nelem * B.nn);
        B.d_emat = upload(a.elem_mat + eoff, (size_t)B.nelem);
        std::vector<int64_t> perm((size_t)B.nelem);
        for (int64_t e = 0; e < B.nelem; e++) perm[(size_t)e] = e;
        B.d_perm = upload(perm.data(), perm.size());
        B.d_N = upload(si.N.data(), si.N.size());
        eoff += B.nelem;
        coff += B.nelem * B.nn;
        ipoff += B.nelem * B.nip;
    }
}

}  // namespace

// recovery.cu on a multi-GPU handle: bring the wrapper's IP state planes up to date (owner's copy of every element)
void amaru_group_refresh_output_state(amaru_model *h) {
    CUDA_CHECK(cudaSetDevice(h->device));
    const int64_t n = h->nip_total;
    if (!h->d_state) CUDA_CHECK(cudaMalloc(&h->d_state, std::max<size_t>((size_t)AMARU_NSTATE * n, 1) * sizeof(double)));
    std::vector<double> sig((size_t)n * 6), eps((size_t)n * 6), epa((size_t)n), dlam((size_t)n);
    char lm[256] = {0};
    const int st = amaru_group_state(h, false, sig.data(), eps.data(), epa.data(), dlam.data(), lm, sizeof(lm));
    if (st != AMARU_OK) throw AmaruError{st, lm};
    struct F { const double *p; int plane0, ncomp; } fields[4] = {{sig.data(), 0, 6}, {eps.data(), 6, 6}, {epa.data(), 12, 1}, {dlam.data(), 13, 1}};
    for (auto &f : fields) {
        CUDA_CHECK(cudaMemcpyAsync(h->d_io, f.p, (size_t)n * f.ncomp * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        amaru_state_permute(h, h->d_io, f.plane0, f.ncomp, true);
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
    }
}

// ------------------------------------------------------------------------------------------------ create / destroy
int amaru_group_create(const CreateArgs &a, int ngpus, const int32_t *devices, int partitioner, amaru_model **out, char *msg,
                       int msglen) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_msg(msg, msglen, "amaru_create: no CUDA device visible; this library has no CPU fallback");
        return AMARU_ERR_NO_DEVICE;
    }
    std::unique_ptr<AmaruGroup> gp(new AmaruGroup());
    AmaruGroup *g = gp.get();
    g->n = ngpus;
    for (int r = 0; r < ngpus; r++) g->dev.push_back(devices ? devices[r] : r);
    for (int r = 0; r < ngpus; r++) {
        if (g->dev[(size_t)r] < 0 || g->dev[(size_t)r] >= ndev) {
            set_msg(msg, msglen, "amaru_create: the box has fewer GPUs than ngpus asks for");
            return AMARU_ERR_ARG;
        }
        for (int q = 0; q < r; q++) {
            int can = 0;
            if (g->dev[(size_t)q] == g->dev[(size_t)r] ||
                cudaDeviceCanAccessPeer(&can, g->dev[(size_t)r], g->dev[(size_t)q]) != cudaSuccess || !can) {
                set_msg(msg, msglen, "amaru_create: ngpus > 1 needs distinct GPUs with peer access (NVLink) between all of them");
                return AMARU_ERR_UNSUPPORTED;
            }
        }
    }
    // global layout
    g->nbatches = a.nbatches;
    std::vector<const int32_t *> connp((size_t)a.nbatches);
    int64_t eoff = 0, coff = 0, ipoff = 0;
    for (int b = 0; b < a.nbatches; b++) {
        ShapeInfo si;
        if (a.batch_shape[b] < AMARU_SHAPE_QUAD4 || a.batch_shape[b] > AMARU_SHAPE_TET10 || !amaru_shape_info(a.batch_shape[b], si)) {
            set_msg(msg, msglen, "amaru_create: cell shape outside the hot path (QUAD4, QUAD8, HEX8, HEX20, TET10)");
            return AMARU_ERR_UNSUPPORTED;
        }
        g->nn.push_back(si.nn);
        g->nip.push_back(si.nip);
        g->nelem.push_back(a.batch_nelem[b]);
        g->elem_off.push_back(eoff);
        g->ip_off.push_back(ipoff);
        connp[(size_t)b] = a.conn + coff;
        for (int64_t i = 0; i < a.batch_nelem[b] * si.nn; i++)
            if (connp[(size_t)b][i] < 0 || connp[(size_t)b][i] >= a.nnodes) {
                set_msg(msg, msglen, "amaru_create: node id out of range");
                return AMARU_ERR_ARG;
            }
        eoff += a.batch_nelem[b];
        coff += a.batch_nelem[b] * si.nn;
        ipoff += a.batch_nelem[b] * si.nip;
    }
    // element partition and node owners (host, once), then one thread per GPU builds its view and its part
    std::vector<int32_t> epart, owner;
    try {
        amaru_partition_elements(partitioner, ngpus, a.nnodes, a.coords, a.nbatches, g->nn.data(), a.batch_shape, a.batch_nelem,
                                 connp.data(), epart);
        amaru_node_owners(a.nnodes, a.nbatches, g->nn.data(), a.batch_nelem, connp.data(), epart, owner);
    } catch (const AmaruError &e) {
        set_msg(msg, msglen, e.msg);
        return e.code;
    }
    for (int r = 0; r < ngpus; r++)
        if (std::find(owner.begin(), owner.end(), r) == owner.end()) {
            set_msg(msg, msglen, "amaru_create: the mesh is too small for ngpus parts (a GPU would own no node)");
            return AMARU_ERR_ARG;
        }
    g->part.assign((size_t)ngpus, nullptr);
    g->view.resize((size_t)ngpus);
    g->lip_off.resize((size_t)ngpus);
    g->status.assign((size_t)ngpus, 0);
    g->msg.assign((size_t)ngpus, std::string());
    g->bar.n = ngpus;
    for (int r = 0; r < ngpus; r++) g->th.emplace_back(worker, g, r);
    const int nd = a.ndim;
    std::string emsg;
    int st = run(g, [&](int r) {
        for (int q = 0; q < ngpus; q++) {
            if (q == r) continue;
            const cudaError_t e = cudaDeviceEnablePeerAccess(g->dev[(size_t)q], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(e);
            cudaGetLastError();
        }
        AmaruLocalView &v = g->view[(size_t)r];
        amaru_local_view(r, ngpus, a.nnodes, a.nbatches, g->nn.data(), a.batch_nelem, connp.data(), owner, v);
        const int64_t nl = (int64_t)v.node_gid.size();
        std::vector<double> coords((size_t)nl * 3);
        std::vector<int32_t> eq((size_t)nl * nd);
        for (int64_t i = 0; i < nl; i++) {
            std::memcpy(&coords[(size_t)i * 3], a.coords + v.node_gid[(size_t)i] * 3, 3 * sizeof(double));
            std::memcpy(&eq[(size_t)i * nd], a.eqid + v.node_gid[(size_t)i] * nd, (size_t)nd * sizeof(int32_t));
        }
        std::vector<int32_t> conn, emat;
        std::vector<int64_t> bn((size_t)a.nbatches);
        g->lip_off[(size_t)r].assign((size_t)a.nbatches, 0);
        int64_t lip = 0;
        for (int b = 0; b < a.nbatches; b++) {
            bn[(size_t)b] = (int64_t)v.elem_gid[(size_t)b].size();
            conn.insert(conn.end(), v.conn[(size_t)b].begin(), v.conn[(size_t)b].end());
            for (int64_t ge : v.elem_gid[(size_t)b]) emat.push_back(a.elem_mat[g->elem_off[(size_t)b] + ge]);
            g->lip_off[(size_t)r][(size_t)b] = lip;
            lip += bn[(size_t)b] * g->nip[(size_t)b];
        }
        CreateArgs la = a;
        la.nnodes = nl;
        la.nowned = v.nowned;
        la.coords = coords.data();
        la.batch_nelem = bn.data();
        la.conn = conn.data();
        la.elem_mat = emat.data();
        la.eqid = eq.data();
        la.prescribed = nullptr;
        la.device = g->dev[(size_t)r];
        la.rank = r;
        la.nranks = ngpus;
        amaru_model *m = amaru_create_impl(la);
        g->part[(size_t)r] = m;
        m->grp_of = g;
        amaru_comm_setup(m, (int)v.neigh.size(), v.neigh.data(), v.send_ptr.data(), v.send_nodes.data(), v.recv_start.data(),
                         v.recv_count.data(), nullptr);
        amaru_set_element_ownership(m, conn.data(), (int)v.neigh.size(), v.neigh.data(), v.recv_start.data(), v.recv_count.data());
        if (r > 0) {   // ABI-order vectors exist once, on the first GPU
            cudaFree(m->d_U);
            cudaFree(m->d_F);
            m->d_U = m->d_F = nullptr;
        }
        m->io_shared = true;
        return AMARU_OK;
    }, &emsg);
    if (st == AMARU_OK) {
        try {
            for (int r = 1; r < ngpus; r++) {
                g->part[(size_t)r]->d_U = g->part[0]->d_U;
                g->part[(size_t)r]->d_F = g->part[0]->d_F;
            }
            amaru_p2p_connect_direct(g->part.data(), ngpus);
        } catch (const AmaruError &e) {
            st = e.code;
            emsg = e.msg;
        }
    }
    amaru_model *h = new amaru_model();
    h->grp = g;
    h->ndim = h->nd = a.ndim;
    h->stressmodel = a.stressmodel;
    h->th = a.thickness;
    h->nnodes = h->nowned = a.nnodes;
    h->ndofs = a.ndofs;
    h->nu = a.nu;
    h->nelem_total = eoff;
    h->nip_total = ipoff;
    h->nmats = a.nmats;
    h->nranks = ngpus;
    h->device = g->dev[0];
    gp.release();
    if (st == AMARU_OK) {
        try {
            output_side_setup(h, a);
        } catch (const AmaruError &e) {
            st = e.code;
            emsg = e.msg;
        }
    }
    if (st != AMARU_OK) {
        amaru_group_destroy(h);
        return finish(st, emsg, msg, msglen);
    }
    *out = h;
    set_msg(msg, msglen, "");
    return AMARU_OK;
}

int amaru_group_destroy(amaru_model *h) {
    AmaruGroup *g = h->grp;
    run(g, [&](int r) {
        amaru_model *m = g->part[(size_t)r];
        if (m) amaru_free_model(m);
        g->part[(size_t)r] = nullptr;
        return AMARU_OK;
    });
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->quit = true;
    }
    g->cv_job.notify_all();
    for (auto &t : g->th) t.join();
    delete g;
    // the wrapper's own (output-side) allocations on the first GPU
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    amaru_recovery_destroy(h);
    for (Batch &B : h->batches) {
        cudaFree(B.d_conn); cudaFree(B.d_emat); cudaFree(B.d_perm); cudaFree(B.d_N);
    }
    cudaFree(h->d_coords); cudaFree(h->d_eqid); cudaFree(h->d_mat_kind); cudaFree(h->d_io); cudaFree(h->d_state);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return AMARU_OK;
}

// ------------------------------------------------------------------------------------------------ sizes
int64_t amaru_group_sum(const amaru_model *h, int what) {
    int64_t s = 0;
    for (amaru_model *m : h->grp->part) s += what == 0 ? m->nblk : (what == 1 ? m->launches : (int64_t)m->ncolors);
    return s;
}
amaru_model *amaru_group_part(const amaru_model *h, int r) { return h->grp->part[(size_t)r]; }

// ------------------------------------------------------------------------------------------------ IP state
int amaru_group_state(amaru_model *h, bool set, double *sigma, double *eps, double *epa, double *dlam, char *msg, int msglen) {
    AmaruGroup *g = h->grp;
    std::string emsg;
    const int st = run(g, [&](int r) {
        amaru_model *m = g->part[(size_t)r];
        const AmaruLocalView &v = g->view[(size_t)r];
        const int64_t n = m->nip_total;
        struct Fd { double *h; int nc; } fields[4] = {{sigma, 6}, {eps, 6}, {epa, 1}, {dlam, 1}};
        std::vector<double> buf[4];
        double *lp[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int f = 0; f < 4; f++)
            if (fields[f].h) {
                buf[f].resize((size_t)std::max<int64_t>(n, 1) * fields[f].nc);
                lp[f] = buf[f].data();
            }
        auto copy_rows = [&](bool to_local) {
            for (int b = 0; b < g->nbatches; b++) {
                const int nip = g->nip[(size_t)b];
                const auto &eg = v.elem_gid[(size_t)b];
                for (size_t le = 0; le < eg.size(); le++) {
                    if (!to_local && !v.elem_owned[(size_t)b][le]) continue;   // the owner's copy is the authoritative one
                    const int64_t gi = g->ip_off[(size_t)b] + eg[le] * nip, li = g->lip_off[(size_t)r][(size_t)b] + (int64_t)le * nip;
                    for (int f = 0; f < 4; f++) {
                        if (!fields[f].h) continue;
                        const size_t bytes = (size_t)nip * fields[f].nc * sizeof(double);
                        if (to_local) std::memcpy(lp[f] + li * fields[f].nc, fields[f].h + gi * fields[f].nc, bytes);
                        else std::memcpy(fields[f].h + gi * fields[f].nc, lp[f] + li * fields[f].nc, bytes);
                    }
                }
            }
        };
        char lm[256] = {0};
        int s;
        if (set) {
            copy_rows(true);
            s = amaru_set_state(m, lp[0], lp[1], lp[2], lp[3], lm, sizeof(lm));
        } else {
            s = amaru_get_state(m, lp[0], lp[1], lp[2], lp[3], lm, sizeof(lm));
            if (s == AMARU_OK) copy_rows(false);
        }
        if (s != AMARU_OK) throw AmaruError{s, lm};
        return AMARU_OK;
    }, &emsg);
    return finish(st, emsg, msg, msglen);
}

int amaru_group_simple(amaru_model *h, int what, double a, double b, char *msg, int msglen) {
    AmaruGroup *g = h->grp;
    std::string emsg;
    const int st = run(g, [&](int r) {
        amaru_model *m = g->part[(size_t)r];
        char lm[256] = {0};
        int s = AMARU_OK;
        switch (what) {
        case 0: s = amaru_state_backup(m); break;
        case 1: s = amaru_state_restore(m); break;
        case 2: s = amaru_assemble_K(m, lm, sizeof(lm)); break;
        case 3: s = amaru_set_system_matrix(m, a, b, lm, sizeof(lm)); break;
        case 4: s = amaru_set_operator(m, (int)a); break;
        case 5: s = amaru_set_profiling(m, (int)a); break;
        case 6: s = amaru_tangent_save(m, lm, sizeof(lm)); break;
        case 7: s = amaru_tangent_blend(m, a, b, lm, sizeof(lm)); break;
        }
        if (s != AMARU_OK) throw AmaruError{s, lm};
        return AMARU_OK;
    }, &emsg);
    return finish(st, emsg, msg, msglen);
}

int amaru_group_assemble_M(amaru_model *h, const double *rho, char *msg, int msglen) {
    AmaruGroup *g = h->grp;
    std::string emsg;
    const int st = run(g, [&](int r) {
        amaru_model *m = g->part[(size_t)r];
        const AmaruLocalView &v = g->view[(size_t)r];
        std::vector<double> lr;
        for (int b = 0; b < g->nbatches; b++)
            for (int64_t ge : v.elem_gid[(size_t)b]) lr.push_back(rho[g->elem_off[(size_t)b] + ge]);
        if (lr.empty()) lr.push_back(0.0);
        char lm[256] = {0};
        const int s = amaru_assemble_M(m, lr.data(), lm, sizeof(lm));
        if (s != AMARU_OK) throw AmaruError{s, lm};
        return AMARU_OK;
    }, &emsg);
    return finish(st, emsg, msg, msglen);
}

// ------------------------------------------------------------------------------------------------ the hot calls
int amaru_group_solve(amaru_model *h, double *U, double *F, double cg_rtol, int cg_maxit, int precond, int *iters, double *relres,
                      char *msg, int msglen) {
    AmaruGroup *g = h->grp;
    const size_t bytes = (size_t)h->ndofs * sizeof(double);
    SolveInfo info0;
    std::string emsg;
    const int st = run(g, [&](int r) {
        amaru_model *m = g->part[(size_t)r];
        if (r == 0) {
            CUDA_CHECK(cudaMemcpyAsync(m->d_U, U, bytes, cudaMemcpyHostToDevice, m->stream));
            CUDA_CHECK(cudaMemcpyAsync(m->d_F, F, bytes, cudaMemcpyHostToDevice, m->stream));
        }
        amaru_group_barrier(m);
        SolveInfo info;
        const int s = amaru_solve_device(m, cg_rtol, cg_maxit, precond, info);
        if (r == 0) {
            info0 = info;
            if (s == AMARU_OK) {
                if (h->nu > 0) CUDA_CHECK(cudaMemcpyAsync(U, m->d_U, (size_t)h->nu * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
                if (h->ndofs > h->nu)
                    CUDA_CHECK(cudaMemcpyAsync(F + h->nu, m->d_F + h->nu, (size_t)(h->ndofs - h->nu) * sizeof(double),
                                               cudaMemcpyDeviceToHost, m->stream));
                CUDA_CHECK(cudaStreamSynchronize(m->stream));
            }
        }
        return s;
    }, &emsg);
    if (iters) *iters = info0.iters;
    if (relres) *relres = info0.relres;
    return finish(st, emsg, msg, msglen);
}

int amaru_group_update(amaru_model *h, const double *dU, double *dFin, int mode, char *msg, int msglen) {
    AmaruGroup *g = h->grp;
    const size_t bytes = (size_t)h->ndofs * sizeof(double);
    std::string emsg;
    const int st = run(g, [&](int r) {
        amaru_model *m = g->part[(size_t)r];
        int s = AMARU_OK;
        if (mode == 0) {   // update_state!
            if (r == 0) CUDA_CHECK(cudaMemcpyAsync(m->d_U, dU, bytes, cudaMemcpyHostToDevice, m->stream));
            amaru_group_barrier(m);
            s = amaru_update_device(m);
        } else {           // elem_internal_forces of the current stress
            amaru_launch_update(m, m->d_x, m->d_f, 1);
            amaru_nodes_to_eq(m, m->d_f, m->d_F, 0);
            amaru_group_barrier(m);
        }
        if (r == 0) {
            CUDA_CHECK(cudaMemcpyAsync(dFin, m->d_F, bytes, cudaMemcpyDeviceToHost, m->stream));
            CUDA_CHECK(cudaStreamSynchronize(m->stream));
        }
        return s;
    }, &emsg);
    return finish(st, emsg, msg, msglen);
}

// y = a*(K x) + b*(M x) (what = 0) or y = A x with the CG operator (what = 1: masked / fused dot as in the CG loop)
int amaru_group_product(amaru_model *h, int what, double a, double b, const double *x, double *y, int masked, double *pAp, char *msg,
                        int msglen) {
    AmaruGroup *g = h->grp;
    const size_t bytes = (size_t)h->ndofs * sizeof(double);
    double pq0 = 0.0;
    std::string emsg;
    const int st = run(g, [&](int r) {
        amaru_model *m = g->part[(size_t)r];
        if (r == 0) CUDA_CHECK(cudaMemcpyAsync(m->d_U, x, bytes, cudaMemcpyHostToDevice, m->stream));
        amaru_group_barrier(m);
        if (what == 0) {
            AMARU_REQUIRE(b == 0.0 || m->d_M != nullptr, AMARU_ERR_ARG, "amaru_matvec: mass matrix not assembled");
            amaru_eq_to_nodes(m, m->d_U, m->d_x);
            amaru_halo_exchange(m, m->d_x);
            amaru_spmv(m, m->d_K, m->d_x, m->d_q, 0);
            if (b != 0.0) amaru_spmv(m, m->d_M, m->d_x, m->d_r, 0);
            amaru_axpby(m, m->nowned * m->nd, a, m->d_q, b, b != 0.0 ? m->d_r : m->d_q, m->d_q);
        } else {
            amaru_eq_to_nodes(m, m->d_U, m->d_p);
            const double pq = amaru_operator_product(m, masked);
            if (r == 0) pq0 = pq;
        }
        amaru_nodes_to_eq(m, m->d_q, m->d_F, 0);
        amaru_group_barrier(m);
        if (r == 0) {
            CUDA_CHECK(cudaMemcpyAsync(y, m->d_F, bytes, cudaMemcpyDeviceToHost, m->stream));
            CUDA_CHECK(cudaStreamSynchronize(m->stream));
        }
        return AMARU_OK;
    }, &emsg);
    if (pAp) *pAp = pq0;
    return finish(st, emsg, msg, msglen);
}

// ------------------------------------------------------------------------------------------------ measurement hooks
int amaru_group_set_device_vectors(amaru_model *h, const double *U, const double *F, char *msg, int msglen) {
    AmaruGroup *g = h->grp;
    const size_t bytes = (size_t)h->ndofs * sizeof(double);
    std::string emsg;
    int st = run(g, [&](int r) {
        amaru_model *m = g->part[(size_t)r];
        if (r == 0) {
            if (!m->d_U0) CUDA_CHECK(cudaMalloc(&m->d_U0, bytes));
            if (!m->d_F0) CUDA_CHECK(cudaMalloc(&m->d_F0, bytes));
            CUDA_CHECK(cudaMemcpyAsync(m->d_U0, U, bytes, cudaMemcpyHostToDevice, m->stream));
            CUDA_CHECK(cudaMemcpyAsync(m->d_F0, F, bytes, cudaMemcpyHostToDevice, m->stream));
            CUDA_CHECK(cudaStreamSynchronize(m->stream));
        }
        return AMARU_OK;
    }, &emsg);
    for (int r = 1; r < g->n; r++) {
        g->part[(size_t)r]->d_U0 = g->part[0]->d_U0;
        g->part[(size_t)r]->d_F0 = g->part[0]->d_F0;
    }
    return finish(st, emsg, msg, msglen);
}

int amaru_group_newton_iteration(amaru_model *h, double cg_rtol, int cg_maxit, int precond, double *phase_ms, int *iters,
                                 double *relres, char *msg, int msglen) {
    AmaruGroup *g = h->grp;
    const size_t bytes = (size_t)h->ndofs * sizeof(double);
    SolveInfo info0;
    std::vector<float> ms((size_t)g->n * 4, 0.f);
    std::string emsg;
    const int st = run(g, [&](int r) {
        amaru_model *m = g->part[(size_t)r];
        AMARU_REQUIRE(m->d_U0 && m->d_F0, AMARU_ERR_ARG, "amaru_newton_iteration_device: call amaru_set_device_vectors first");
        struct Ev {
            cudaEvent_t ev[4] = {};
            ~Ev() {
                for (auto &e : ev)
                    if (e) cudaEventDestroy(e);
            }
        } evs;
        cudaEvent_t *ev = evs.ev;
        for (int i = 0; i < 4; i++) CUDA_CHECK(cudaEventCreate(&ev[i]));
        amaru_group_barrier(m);
        CUDA_CHECK(cudaEventRecord(ev[0], m->stream));
        amaru_reset_status(m);
        amaru_launch_assemble(m, 0);
        m->blended = false;
        amaru_combine_matrix(m);
        amaru_ebe_refresh(m);
        const int as = amaru_read_status(m);
        CUDA_CHECK(cudaEventRecord(ev[1], m->stream));
        if (r == 0) {
            CUDA_CHECK(cudaMemcpyAsync(m->d_U, m->d_U0, bytes, cudaMemcpyDeviceToDevice, m->stream));
            CUDA_CHECK(cudaMemcpyAsync(m->d_F, m->d_F0, bytes, cudaMemcpyDeviceToDevice, m->stream));
        }
        amaru_group_barrier(m);
        SolveInfo info;
        int s = amaru_solve_device(m, cg_rtol, cg_maxit, precond, info);
        if (r == 0) info0 = info;
        CUDA_CHECK(cudaEventRecord(ev[2], m->stream));
        amaru_state_restore(m);
        const int s2 = amaru_update_device(m);
        CUDA_CHECK(cudaEventRecord(ev[3], m->stream));
        CUDA_CHECK(cudaEventSynchronize(ev[3]));
        for (int i = 0; i < 3; i++) CUDA_CHECK(cudaEventElapsedTime(&ms[(size_t)r * 4 + i], ev[i], ev[i + 1]));
        CUDA_CHECK(cudaEventElapsedTime(&ms[(size_t)r * 4 + 3], ev[0], ev[3]));
        if (as) s = as;
        if (s == AMARU_OK) s = s2;
        return s;
    }, &emsg);
    if (phase_ms)
        for (int i = 0; i < 4; i++) {   // max over the GPUs
            float t = 0.f;
            for (int r = 0; r < g->n; r++) t = std::max(t, ms[(size_t)r * 4 + i]);
            phase_ms[i] = t;
        }
    if (iters) *iters = info0.iters;
    if (relres) *relres = info0.relres;
    return finish(st, emsg, msg, msglen);
}

// failure injection for the bounded peer-memory waits: every rank but `skip_rank` enters one scalar all-reduce
int amaru_group_comm_selftest(amaru_model *h, int skip_rank, char *msg, int msglen) {
    AmaruGroup *g = h->grp;
    std::string emsg;
    const int st = run(g, [&](int r) {
        amaru_model *m = g->part[(size_t)r];
        if (r == skip_rank) return (int)AMARU_OK;
        CUDA_CHECK(cudaMemsetAsync(m->d_scal, 0, sizeof(CgScalars), m->stream));
        amaru_allreduce_sum(m, m->d_scal->acc, 1);
        CUDA_CHECK(cudaStreamSynchronize(m->stream));
        amaru_comm_check(m);
        return (int)AMARU_OK;
    }, &emsg);
    return finish(st, emsg, msg, msglen);
}
