// extern "C" surface of libamaru_b200.so (include/amaru_b200.h).  Each entry point replaces one Julia call of the
// reference's mech_stage_solver! (src/mech/mech-solver.jl:186-492); see the header for the mapping.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <numeric>

#include "amaru_internal.h"
#include "partition.h"

namespace {

void set_msg(char *msg, int msglen, const std::string &s) {
    if (msg && msglen > 0) {
        std::snprintf(msg, (size_t)msglen, "%s", s.c_str());
    }
}

template <class F>
int guarded(char *msg, int msglen, F f) {
    try {
        set_msg(msg, msglen, "");
        return f();
    } catch (const AmaruError &e) {
        set_msg(msg, msglen, e.msg);
        return e.code;
    } catch (const std::exception &e) {
        set_msg(msg, msglen, e.what());
        return AMARU_ERR_ARG;
    }
}

template <class T>
T *upload(const T *h, size_t n) {
    T *d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) CUDA_CHECK(cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

void use_device(const amaru_model *m) { CUDA_CHECK(cudaSetDevice(m->device)); }

template <int N>
struct EventSet {   // CUDA events released on every exit path
    cudaEvent_t ev[N] = {};
    EventSet() {
        for (auto &e : ev) CUDA_CHECK(cudaEventCreate(&e));
    }
    ~EventSet() {
        for (auto &e : ev)
            if (e) cudaEventDestroy(e);
    }
};

const char *status_text(int st) {
    switch (st) {
    case AMARU_FAIL_MATERIAL: return "VonMisses: Negative value for √J2D";   // von-mises.jl:146
    case AMARU_FAIL_NAN: return "solve_system!: NaN values in internal forces vector";  // mech-solver.jl:142
    case AMARU_FAIL_SINGULAR: return "solve_system!: Possible syngular matrix";          // solver.jl:70
    case AMARU_FAIL_NEG_JACOBIAN: return "Negative Jacobian determinant in cell";        // mech-solid.jl:150
    case AMARU_FAIL_CG_NOCONV: return "solve_system!: PCG did not reach cg_rtol within cg_maxit iterations";
    case AMARU_FAIL_TANGENT: return "AssertionError: j2d > 0";                           // von-mises.jl:117
    }
    return "";
}

int read_status(amaru_model *m) {
    int st = 0;
    CUDA_CHECK(cudaMemcpyAsync(&st, m->d_status, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    CUDA_CHECK(cudaStreamSynchronize(m->stream));
    return st;
}
void reset_status(amaru_model *m) { CUDA_CHECK(cudaMemsetAsync(m->d_status, 0, sizeof(int), m->stream)); }

}  // namespace

amaru_model *amaru_create_impl(const CreateArgs &a) {
    AMARU_REQUIRE(a.ndim == 2 || a.ndim == 3, AMARU_ERR_ARG, "amaru_create: ndim must be 2 or 3");
    AMARU_REQUIRE(a.stressmodel >= AMARU_STRESS_D3 && a.stressmodel <= AMARU_STRESS_AXISYMMETRIC, AMARU_ERR_UNSUPPORTED,
                  "amaru_create: unknown stress model (d3, planestrain, planestress, axisymmetric)");
    AMARU_REQUIRE((a.stressmodel != AMARU_STRESS_PLANESTRESS && a.stressmodel != AMARU_STRESS_AXISYMMETRIC) || a.ndim == 2,
                  AMARU_ERR_ARG, "amaru_create: planestress / axisymmetric need ndim == 2");
    AMARU_REQUIRE(a.nnodes > 0 && a.nbatches > 0 && a.nmats > 0, AMARU_ERR_ARG, "amaru_create: empty model");
    AMARU_REQUIRE(a.coords && a.conn && a.elem_mat && a.mat_kind && a.mat_params && a.eqid && a.batch_shape &&
                      a.batch_nelem, AMARU_ERR_ARG, "amaru_create: null pointer");
    AMARU_REQUIRE(a.thickness > 0, AMARU_ERR_ARG, "amaru_create: thickness must be > 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw AmaruError{AMARU_ERR_NO_DEVICE, "amaru_create: no CUDA device visible; this library has no CPU fallback"};
    AMARU_REQUIRE(a.device >= 0 && a.device < ndev, AMARU_ERR_ARG, "amaru_create: bad device ordinal");
    for (int i = 0; i < a.nmats; i++) {
        const int k = a.mat_kind[i];
        AMARU_REQUIRE(k == AMARU_MAT_LINEAR_ELASTIC || k == AMARU_MAT_VON_MISES || k == AMARU_MAT_DRUCKER_PRAGER,
                      AMARU_ERR_UNSUPPORTED, "amaru_create: material outside the hot path (LinearElastic, VonMises, DruckerPrager)");
        AMARU_REQUIRE(a.stressmodel != AMARU_STRESS_PLANESTRESS || k == AMARU_MAT_LINEAR_ELASTIC, AMARU_ERR_UNSUPPORTED,
                      "amaru_create: planestress is available for LinearElastic only (linear-elastic.jl:99-108)");
    }

    // on a throw below, release every device allocation made so far (not only the struct)
    struct ModelGuard {
        amaru_model *m;
        ~ModelGuard() { if (m) amaru_free_model(m); }
        amaru_model *release() { amaru_model *r = m; m = nullptr; return r; }
    } mp{new amaru_model()};
    amaru_model *m = mp.m;
    m->device = a.device;
    CUDA_CHECK(cudaSetDevice(a.device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, a.device));
    m->nsm = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    m->ndim = m->nd = a.ndim;
    m->stressmodel = a.stressmodel;
    m->th = a.thickness;
    m->nnodes = a.nnodes;
    m->nowned = a.nowned;
    m->ndofs = a.ndofs;
    m->nu = a.nu;
    m->nmats = a.nmats;
    m->rank = a.rank;
    m->nranks = a.nranks;

    // batches
    std::vector<ShapeInfo> info(a.nbatches);
    std::vector<int> nn(a.nbatches);
    std::vector<const int32_t *> connp(a.nbatches);
    int64_t eoff = 0, coff = 0, ipoff = 0;
    m->batches.resize(a.nbatches);
    for (int b = 0; b < a.nbatches; b++) {
        AMARU_REQUIRE(a.batch_shape[b] >= AMARU_SHAPE_QUAD4 && a.batch_shape[b] <= AMARU_SHAPE_TET10 &&
                          amaru_shape_info(a.batch_shape[b], info[b]), AMARU_ERR_UNSUPPORTED,
                      "amaru_create: cell shape outside the hot path (QUAD4, QUAD8, HEX8, HEX20, TET10)");
        AMARU_REQUIRE(info[b].nd == a.ndim, AMARU_ERR_ARG, "amaru_create: cell shape dimension differs from ndim");
        Batch &B = m->batches[b];
        B.shape = info[b].id; B.nn = info[b].nn; B.nd = info[b].nd; B.nip = info[b].nip;
        B.nelem = a.batch_nelem[b];
        B.elem_off = eoff;
        B.ip_off = ipoff;
        nn[b] = B.nn;
        connp[b] = a.conn + coff;
        for (int64_t i = 0; i < B.nelem * B.nn; i++)
            AMARU_REQUIRE(connp[b][i] >= 0 && connp[b][i] < a.nnodes, AMARU_ERR_ARG, "amaru_create: node id out of range");
        eoff += B.nelem;
        coff += B.nelem * B.nn;
        ipoff += B.nelem * B.nip;
    }
    m->nelem_total = eoff;
    m->nip_total = ipoff;
    for (int64_t e = 0; e < eoff; e++)
        AMARU_REQUIRE(a.elem_mat[e] >= 0 && a.elem_mat[e] < a.nmats, AMARU_ERR_ARG, "amaru_create: material index out of range");

    // host preprocessing: adjacency, colouring, symbolic pattern
    std::vector<int64_t> adj_ptr, adj;
    amaru_build_adjacency(a.nnodes, a.nbatches, nn.data(), a.batch_nelem, connp.data(), adj_ptr, adj);
    std::vector<int32_t> color;
    m->ncolors = amaru_color_elements(a.nnodes, a.nbatches, nn.data(), a.batch_nelem, connp.data(), adj_ptr, adj, color);
    AMARU_REQUIRE(m->ncolors > 0, AMARU_ERR_ARG, "amaru_create: element colouring needs more than 512 colours");
    HostPattern pat;
    amaru_build_pattern(a.nowned, a.nbatches, nn.data(), a.batch_nelem, connp.data(), adj_ptr, adj, pat);
    m->nblk = (int64_t)pat.col.size();

    // colour-sort each batch (stable: ascending element id inside a colour)
    for (int b = 0; b < a.nbatches; b++) {
        Batch &B = m->batches[b];
        const int32_t *col_b = color.data() + B.elem_off;
        std::vector<int64_t> cnt((size_t)m->ncolors + 1, 0);
        for (int64_t e = 0; e < B.nelem; e++) cnt[(size_t)col_b[e] + 1]++;
        for (int c = 0; c < m->ncolors; c++) cnt[c + 1] += cnt[c];
        B.color_off.assign(cnt.begin(), cnt.end());
        std::vector<int64_t> perm((size_t)B.nelem);
        {
            std::vector<int64_t> fill(cnt.begin(), cnt.end() - 1);
            for (int64_t e = 0; e < B.nelem; e++) perm[(size_t)fill[col_b[e]]++] = e;
        }
        std::vector<int32_t> sconn((size_t)B.nelem * B.nn), smat((size_t)B.nelem);
        for (int64_t s = 0; s < B.nelem; s++) {
            const int64_t e = perm[(size_t)s];
            std::memcpy(&sconn[(size_t)s * B.nn], connp[b] + e * B.nn, sizeof(int32_t) * B.nn);
            smat[(size_t)s] = a.elem_mat[B.elem_off + e];
        }
        B.d_conn = upload(sconn.data(), sconn.size());
        B.d_emat = upload(smat.data(), smat.size());
        B.d_perm = upload(perm.data(), perm.size());
        B.d_dNdR = upload(info[b].dNdR.data(), info[b].dNdR.size());
        B.d_N = upload(info[b].N.data(), info[b].N.size());
        std::vector<double> w((size_t)B.nip);
        for (int q = 0; q < B.nip; q++) w[q] = info[b].ips[4 * q + 3];
        B.d_w = upload(w.data(), w.size());
        CUDA_CHECK(cudaMalloc(&B.d_map, std::max<size_t>((size_t)B.nelem * B.nn * B.nn, 1) * sizeof(int32_t)));
    }

    // model arrays
    m->d_coords = upload(a.coords, (size_t)a.nnodes * 3);
    m->d_eqid = upload(a.eqid, (size_t)a.nnodes * m->nd);
    {
        std::vector<uint8_t> fx((size_t)a.nnodes * m->nd);
        for (size_t i = 0; i < fx.size(); i++) {
            AMARU_REQUIRE(a.eqid[i] >= 0 && a.eqid[i] < a.ndofs, AMARU_ERR_ARG, "amaru_create: eq_id out of range");
            fx[i] = a.prescribed ? a.prescribed[i] : (uint8_t)(a.eqid[i] >= a.nu);
        }
        m->d_fixed = upload(fx.data(), fx.size());
        m->h_fixed.swap(fx);
    }
    m->d_mat_kind = upload(a.mat_kind, (size_t)a.nmats);
    {
        // Plane stress (linear-elastic.jl:99-108: c = E/(1-ν²); c, cν, c(1-ν)) restricted to the in-plane components IS the
        // 3D / plane-strain matrix of the material E* = E(1+2ν)/(1+ν)², ν* = ν/(1+ν) (same shear modulus): every kernel keeps
        // its one constitutive form and the device gets the equivalent constants; σzz, the only place where the two differ,
        // is cleared after every state update (amaru_update_device).
        std::vector<double> par(a.mat_params, a.mat_params + (size_t)a.nmats * AMARU_MAT_NPARAMS);
        if (a.stressmodel == AMARU_STRESS_PLANESTRESS)
            for (int i = 0; i < a.nmats; i++) {
                const double E = par[(size_t)i * AMARU_MAT_NPARAMS], nu = par[(size_t)i * AMARU_MAT_NPARAMS + 1];
                par[(size_t)i * AMARU_MAT_NPARAMS] = E * (1.0 + 2.0 * nu) / ((1.0 + nu) * (1.0 + nu));
                par[(size_t)i * AMARU_MAT_NPARAMS + 1] = nu / (1.0 + nu);
            }
        m->d_mat_par = upload(par.data(), par.size());
    }
    m->h_eqid.assign(a.eqid, a.eqid + (size_t)a.nnodes * m->nd);

    // pattern + matrix
    m->d_rowptr = upload(pat.rowptr.data(), pat.rowptr.size());
    m->d_col = upload(pat.col.data(), pat.col.size());
    m->d_diag = upload(pat.diag.data(), pat.diag.size());
    m->h_rowptr.swap(pat.rowptr);
    m->h_col.swap(pat.col);
    // +256 B slack: the streamed SpMV reads the value array in 16-byte granules (bulk async copies)
    const size_t kbytes = std::max<size_t>((size_t)m->nblk * m->nd * m->nd, 1) * sizeof(double) + 256;
    CUDA_CHECK(cudaMalloc(&m->d_K, kbytes));
    CUDA_CHECK(cudaMemset(m->d_K, 0, kbytes));
    m->d_A = m->d_K;
    for (Batch &B : m->batches) amaru_build_map(m, B);

    // IP state
    const size_t sbytes = std::max<size_t>((size_t)AMARU_NSTATE * m->nip_total, 1) * sizeof(double);
    CUDA_CHECK(cudaMalloc(&m->d_state, sbytes));
    CUDA_CHECK(cudaMalloc(&m->d_statebk, sbytes));
    CUDA_CHECK(cudaMemset(m->d_state, 0, sbytes));
    CUDA_CHECK(cudaMemset(m->d_statebk, 0, sbytes));

    // staging
    m->io_len = std::max<int64_t>(m->ndofs, 6 * m->nip_total);
    CUDA_CHECK(cudaMalloc(&m->d_io, (size_t)m->io_len * sizeof(double)));
    CUDA_CHECK(cudaMalloc(&m->d_U, (size_t)m->ndofs * sizeof(double)));
    CUDA_CHECK(cudaMalloc(&m->d_F, (size_t)m->ndofs * sizeof(double)));
    CUDA_CHECK(cudaMemset(m->d_U, 0, (size_t)m->ndofs * sizeof(double)));
    CUDA_CHECK(cudaMemset(m->d_F, 0, (size_t)m->ndofs * sizeof(double)));
    CUDA_CHECK(cudaMalloc(&m->d_status, sizeof(int)));
    CUDA_CHECK(cudaMemset(m->d_status, 0, sizeof(int)));
    amaru_ebe_setup(m);
    amaru_pcg_setup(m);
    CUDA_CHECK(cudaStreamSynchronize(m->stream));
    return mp.release();
}

namespace {

// An element is counted (p.Ap of the matrix-free operator) on the lowest rank that owns one of its nodes — the rank whose
// copy of the element's IP state is authoritative (amaru_jl_b200/partition.py).  Node owners follow from the halo lists:
// local nodes below nowned are ours, every ghost range belongs to one neighbour.
void set_element_ownership(amaru_model *m, const int32_t *conn, int nneigh, const int32_t *neigh_rank,
                           const int64_t *recv_start, const int64_t *recv_count) {
    std::vector<int32_t> owner((size_t)m->nnodes, m->rank);
    for (int q = 0; q < nneigh; q++)
        for (int64_t i = 0; i < recv_count[q]; i++) owner[(size_t)(recv_start[q] + i)] = neigh_rank[q];
    int64_t coff = 0;
    for (size_t bi = 0; bi < m->batches.size(); bi++) {
        Batch &B = m->batches[bi];
        std::vector<int64_t> perm((size_t)B.nelem);
        if (B.nelem) CUDA_CHECK(cudaMemcpy(perm.data(), B.d_perm, perm.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
        std::vector<uint8_t> own((size_t)B.nelem);
        for (int64_t s = 0; s < B.nelem; s++) {
            const int32_t *c = conn + coff + perm[(size_t)s] * B.nn;
            int32_t lo = owner[(size_t)c[0]];
            for (int a = 1; a < B.nn; a++) lo = std::min(lo, owner[(size_t)c[a]]);
            own[(size_t)s] = (uint8_t)(lo == m->rank);
        }
        amaru_ebe_set_owned(m, (int)bi, own.data());
        coff += B.nelem * B.nn;
    }
}

void free_model(amaru_model *m) {
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    amaru_comm_destroy(m);
    amaru_recovery_destroy(m);
    amaru_ebe_destroy(m);
    for (Batch &B : m->batches) {
        cudaFree(B.d_conn); cudaFree(B.d_emat); cudaFree(B.d_map); cudaFree(B.d_perm); cudaFree(B.d_owned);
        cudaFree(B.d_rho); cudaFree(B.d_dNdR); cudaFree(B.d_N); cudaFree(B.d_w);
    }
    cudaFree(m->d_Abuf);
    if (m->io_shared && m->rank > 0) m->d_U = m->d_F = m->d_U0 = m->d_F0 = nullptr;   // they live on the group's first GPU
    for (void *p : {(void *)m->d_coords, (void *)m->d_eqid, (void *)m->d_fixed, (void *)m->d_mat_kind, (void *)m->d_mat_par,
                    (void *)m->d_rowptr, (void *)m->d_col, (void *)m->d_diag, (void *)m->d_K, (void *)m->d_Ksave, (void *)m->d_M,
                    (void *)m->d_Minv, (void *)m->d_state, (void *)m->d_statebk, (void *)m->d_x, (void *)m->d_r,
                    (void *)m->d_z, (void *)m->d_p, (void *)m->d_q, (void *)m->d_b, (void *)m->d_f, (void *)m->d_io,
                    (void *)m->d_U, (void *)m->d_F, (void *)m->d_U0, (void *)m->d_F0, (void *)m->d_tiles, (void *)m->d_tmeta,
                    (void *)m->d_stiles, (void *)m->d_stmeta, (void *)m->d_usrc, (void *)m->d_Asym,
                    (void *)m->d_partial, (void *)m->d_scal, (void *)m->d_status})
        cudaFree(p);
    if (m->h_pinned) cudaFreeHost(m->h_pinned);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

// core of amaru_solve on device-resident ABI-order vectors d_U / d_F
int solve_device(amaru_model *m, double cg_rtol, int cg_maxit, int precond, SolveInfo &info) {
    amaru_eq_to_nodes(m, m->d_U, m->d_x);
    amaru_eq_to_nodes(m, m->d_F, m->d_b);
    amaru_pcg_solve(m, cg_rtol, cg_maxit, precond, info);
    if (m->nranks == 1) {
        amaru_nodes_to_eq(m, m->d_x, m->d_U, 1);   // U[1:nu]     .= U1   (solver.jl:74)
        amaru_nodes_to_eq(m, m->d_q, m->d_F, 2);   // F[nu+1:end] .= F2   (solver.jl:75)
    } else if (m->io_shared) {
        // one process, d_U / d_F are ONE pair of ABI-order vectors on the first GPU: once every rank has read its inputs,
        // each rank stores the entries of the rows it owns straight into them over NVLink (disjoint, no reduction)
        amaru_group_barrier(m);
        amaru_nodes_to_eq(m, m->d_x, m->d_U, 1);
        amaru_nodes_to_eq(m, m->d_q, m->d_F, 2);
        amaru_group_barrier(m);
    } else {
        // every rank contributes the entries of the rows it owns; one all-reduce per vector rebuilds the global ones
        const size_t bytes = (size_t)m->ndofs * sizeof(double);
        CUDA_CHECK(cudaMemsetAsync(m->d_io, 0, bytes, m->stream));
        amaru_nodes_to_eq(m, m->d_x, m->d_io, 1);
        amaru_allreduce_sum(m, m->d_io, m->ndofs);
        if (m->nu > 0) CUDA_CHECK(cudaMemcpyAsync(m->d_U, m->d_io, (size_t)m->nu * sizeof(double), cudaMemcpyDeviceToDevice, m->stream));
        CUDA_CHECK(cudaMemsetAsync(m->d_io, 0, bytes, m->stream));
        amaru_nodes_to_eq(m, m->d_q, m->d_io, 2);
        amaru_allreduce_sum(m, m->d_io, m->ndofs);
        if (m->ndofs > m->nu)
            CUDA_CHECK(cudaMemcpyAsync(m->d_F + m->nu, m->d_io + m->nu, (size_t)(m->ndofs - m->nu) * sizeof(double),
                                       cudaMemcpyDeviceToDevice, m->stream));
    }
    if (!info.converged) return AMARU_FAIL_CG_NOCONV;
    if (!(info.maxabs <= 1e8)) return AMARU_FAIL_SINGULAR;   // solver.jl:68-71 (NaN also lands here)
    return AMARU_OK;
}

int update_device(amaru_model *m) {
    reset_status(m);
    amaru_eq_to_nodes(m, m->d_U, m->d_x);
    amaru_launch_update(m, m->d_x, m->d_f, 0);
    if (m->io_shared) {   // every dof is owned by exactly one rank: the owners' stores rebuild the ABI vector on the first GPU
        amaru_group_barrier(m);                      // every rank has read ΔU from d_U
        amaru_nodes_to_eq(m, m->d_f, m->d_F, 0);
        amaru_group_barrier(m);
        const int st = read_status(m);               // the caller takes the maximum over the ranks
        if (st) return st;
        return amaru_check_nan(m, m->d_f, m->nowned * m->nd) ? AMARU_FAIL_NAN : AMARU_OK;
    }
    CUDA_CHECK(cudaMemsetAsync(m->d_F, 0, (size_t)m->ndofs * sizeof(double), m->stream));
    amaru_nodes_to_eq(m, m->d_f, m->d_F, 0);
    if (m->nranks > 1) {
        amaru_allreduce_sum(m, m->d_F, m->ndofs);
        amaru_allreduce_max_int(m, m->d_status);
    }
    const int st = read_status(m);
    if (st) return st;
    const int nan = m->nranks > 1 ? amaru_check_nan(m, m->d_F, m->ndofs) : amaru_check_nan(m, m->d_f, m->nowned * m->nd);
    if (nan) return AMARU_FAIL_NAN;
    return AMARU_OK;
}

}  // namespace

// internals shared with group.cu (multi-GPU handles of one process)
void amaru_set_element_ownership(amaru_model *m, const int32_t *conn, int nneigh, const int32_t *neigh_rank,
                                 const int64_t *recv_start, const int64_t *recv_count) {
    set_element_ownership(m, conn, nneigh, neigh_rank, recv_start, recv_count);
}
void amaru_free_model(amaru_model *m) { free_model(m); }
int amaru_solve_device(amaru_model *m, double cg_rtol, int cg_maxit, int precond, SolveInfo &info) {
    return solve_device(m, cg_rtol, cg_maxit, precond, info);
}
int amaru_update_device(amaru_model *m) { return update_device(m); }
const char *amaru_status_text(int st) { return status_text(st); }
int amaru_read_status(amaru_model *m) { return read_status(m); }
void amaru_reset_status(amaru_model *m) { reset_status(m); }

extern "C" {

int amaru_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *amaru_version(void) { return "amaru_b200 0.1 (sm_100a)"; }

int amaru_create(int ndim, int stressmodel, double thickness, int64_t nnodes, const double *coords, int nbatches,
                 const int32_t *batch_shape, const int64_t *batch_nelem, const int32_t *conn, const int32_t *elem_mat,
                 int nmats, const int32_t *mat_kind, const double *mat_params, const int32_t *eqid, int64_t ndofs,
                 int64_t nu, int ngpus, const int32_t *devices, int partitioner, amaru_model **out, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(out != nullptr, AMARU_ERR_ARG, "amaru_create: out is NULL");
        *out = nullptr;
        AMARU_REQUIRE(ndofs == nnodes * ndim, AMARU_ERR_ARG, "amaru_create: ndofs must equal nnodes*ndim");
        AMARU_REQUIRE(nu >= 0 && nu <= ndofs, AMARU_ERR_ARG, "amaru_create: nu out of range");
        AMARU_REQUIRE(ngpus >= 1 && ngpus <= 16, AMARU_ERR_ARG, "amaru_create: ngpus must be 1..16");
        AMARU_REQUIRE(partitioner == AMARU_PARTITION_RCB || partitioner == AMARU_PARTITION_METIS, AMARU_ERR_ARG, "amaru_create: bad partitioner");
        CreateArgs a{ndim, stressmodel, thickness, nnodes, nnodes, coords, nbatches, batch_shape, batch_nelem, conn,
                     elem_mat, nmats, mat_kind, mat_params, eqid, nullptr, ndofs, nu, devices ? devices[0] : 0, 0, 1};
        if (ngpus > 1) {
            AMARU_REQUIRE(nnodes > 0 && nbatches > 0 && nmats > 0 && coords && conn && elem_mat && mat_kind && mat_params && eqid &&
                              batch_shape && batch_nelem, AMARU_ERR_ARG, "amaru_create: empty model / null pointer");
            return amaru_group_create(a, ngpus, devices, partitioner, out, msg, msglen);
        }
        *out = amaru_create_impl(a);
        return AMARU_OK;
    });
}

int amaru_partition_elements_abi(int partitioner, int nparts, int64_t nnodes, const double *coords, int nbatches,
                                 const int32_t *batch_shape, const int64_t *batch_nelem, const int32_t *conn, int32_t *elem_part,
                                 char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(coords && batch_shape && batch_nelem && conn && elem_part && nparts >= 1 && nbatches > 0, AMARU_ERR_ARG,
                      "amaru_partition_elements: bad argument");
        AMARU_REQUIRE(partitioner == AMARU_PARTITION_RCB || partitioner == AMARU_PARTITION_METIS, AMARU_ERR_ARG, "bad partitioner");
        std::vector<int> nn((size_t)nbatches);
        std::vector<const int32_t *> connp((size_t)nbatches);
        int64_t coff = 0, total = 0;
        for (int b = 0; b < nbatches; b++) {
            ShapeInfo si;
            AMARU_REQUIRE(amaru_shape_info(batch_shape[b], si), AMARU_ERR_UNSUPPORTED, "unknown cell shape");
            nn[(size_t)b] = si.nn;
            connp[(size_t)b] = conn + coff;
            for (int64_t i = 0; i < batch_nelem[b] * si.nn; i++)
                AMARU_REQUIRE(conn[coff + i] >= 0 && conn[coff + i] < nnodes, AMARU_ERR_ARG, "node id out of range");
            coff += batch_nelem[b] * si.nn;
            total += batch_nelem[b];
        }
        std::vector<int32_t> part;
        amaru_partition_elements(partitioner, nparts, nnodes, coords, nbatches, nn.data(), batch_shape, batch_nelem, connp.data(), part);
        std::memcpy(elem_part, part.data(), (size_t)total * sizeof(int32_t));
        return AMARU_OK;
    });
}

int amaru_create_partitioned(int ndim, int stressmodel, double thickness, int64_t nnodes, int64_t nowned,
                             const double *coords, int nbatches, const int32_t *batch_shape, const int64_t *batch_nelem,
                             const int32_t *conn, const int32_t *elem_mat, int nmats, const int32_t *mat_kind,
                             const double *mat_params, const int32_t *eqid, int64_t ndofs, int64_t nu, int rank, int nranks,
                             int nneigh, const int32_t *neigh_rank, const int64_t *send_ptr, const int32_t *send_nodes,
                             const int64_t *recv_start, const int64_t *recv_count, const void *nccl_uid, int device,
                             amaru_model **out, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(out != nullptr, AMARU_ERR_ARG, "amaru_create_partitioned: out is NULL");
        *out = nullptr;
        AMARU_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, AMARU_ERR_ARG, "amaru_create_partitioned: bad rank");
        AMARU_REQUIRE(nowned >= 0 && nowned <= nnodes, AMARU_ERR_ARG, "amaru_create_partitioned: nowned out of range");
        AMARU_REQUIRE(nu >= 0 && nu <= ndofs, AMARU_ERR_ARG, "amaru_create_partitioned: nu out of range");
        AMARU_REQUIRE(nneigh == 0 || (neigh_rank && send_ptr && recv_start && recv_count && nccl_uid), AMARU_ERR_ARG,
                      "amaru_create_partitioned: null halo lists");
        CreateArgs a{ndim, stressmodel, thickness, nnodes, nowned, coords, nbatches, batch_shape, batch_nelem, conn,
                     elem_mat, nmats, mat_kind, mat_params, eqid, nullptr, ndofs, nu, device, rank, nranks};
        amaru_model *m = amaru_create_impl(a);
        try {
            if (nranks > 1) {
                amaru_comm_setup(m, nneigh, neigh_rank, send_ptr, send_nodes, recv_start, recv_count, nccl_uid);
                set_element_ownership(m, conn, nneigh, neigh_rank, recv_start, recv_count);
            }
        } catch (...) {
            free_model(m);
            throw;
        }
        *out = m;
        return AMARU_OK;
    });
}

int amaru_destroy(amaru_model *m) {
    if (!m) return AMARU_ERR_ARG;
    if (m->grp) return amaru_group_destroy(m);
    free_model(m);
    return AMARU_OK;
}

int64_t amaru_nip_total(const amaru_model *m) { return m ? m->nip_total : -1; }
int64_t amaru_nnz(const amaru_model *m) { return m ? amaru_nblocks(m) * m->nd * m->nd : -1; }
int64_t amaru_nblocks(const amaru_model *m) { return !m ? -1 : (m->grp ? amaru_group_sum(m, 0) : m->nblk); }
int amaru_ncolors(const amaru_model *m) { return !m ? -1 : (m->grp ? amaru_group_part(m, 0)->ncolors : m->ncolors); }
int64_t amaru_launch_count(const amaru_model *m) { return !m ? -1 : (m->grp ? amaru_group_sum(m, 1) : m->launches); }
int amaru_ngpus(const amaru_model *m) { return !m ? -1 : (m->grp ? m->nranks : 1); }
int64_t amaru_spmv_bytes(const amaru_model *m) {
    if (!m) return -1;
    if (m->grp) return amaru_spmv_bytes(amaru_group_part(m, 0));
    const int64_t b2 = (int64_t)m->nd * m->nd, n = m->nowned * m->nd;
    if (m->op_ebe && !m->blended) return amaru_ebe_bytes(m);
    if (m->use_sym)   // CG product from the symmetric storage: upper blocks + records + x once + y zeroed and reduced into once
        return m->nublk * b2 * 8 + m->sym_meta_bytes + 8 * n + 16 * n;
    const int64_t meta = m->use_tma ? m->spmv_meta_bytes : m->nblk * 4 + (m->nowned + 1) * 4 + n;   // + fixed mask
    return m->nblk * b2 * 8 + meta + 8 * n + 8 * n;
}

const char *amaru_spmv_kernel(const amaru_model *m) {
    if (!m) return "";
    if (m->grp) return amaru_spmv_kernel(amaru_group_part(m, 0));
    if (m->op_ebe && !m->blended) return amaru_ebe_kernel(m);
    if (m->use_sym) return m->nd == 3 ? "k_spmv_sym<3,true>" : "k_spmv_sym<2,true>";
    if (m->use_tma) return m->nd == 3 ? "k_spmv_stream2<3,true>" : "k_spmv_stream2<2,true>";
    return m->nd == 3 ? "k_spmv<3,true>" : "k_spmv<2,true>";
}

int amaru_set_state(amaru_model *m, const double *sigma, const double *eps, const double *epa, const double *dlam,
                    char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m, AMARU_ERR_ARG, "null handle");
        if (m->grp)
            return amaru_group_state(m, true, const_cast<double *>(sigma), const_cast<double *>(eps), const_cast<double *>(epa),
                                     const_cast<double *>(dlam), msg, msglen);
        use_device(m);
        const int64_t n = m->nip_total;
        struct F { const double *h; int plane0, ncomp; } fields[4] = {{sigma, 0, 6}, {eps, 6, 6}, {epa, 12, 1}, {dlam, 13, 1}};
        for (auto &f : fields) {
            if (!f.h) continue;
            CUDA_CHECK(cudaMemcpyAsync(m->d_io, f.h, (size_t)n * f.ncomp * sizeof(double), cudaMemcpyHostToDevice, m->stream));
            amaru_state_permute(m, m->d_io, f.plane0, f.ncomp, true);
            CUDA_CHECK(cudaStreamSynchronize(m->stream));
        }
        return AMARU_OK;
    });
}

int amaru_get_state(amaru_model *m, double *sigma, double *eps, double *epa, double *dlam, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m, AMARU_ERR_ARG, "null handle");
        if (m->grp) return amaru_group_state(m, false, sigma, eps, epa, dlam, msg, msglen);
        use_device(m);
        const int64_t n = m->nip_total;
        struct F { double *h; int plane0, ncomp; } fields[4] = {{sigma, 0, 6}, {eps, 6, 6}, {epa, 12, 1}, {dlam, 13, 1}};
        for (auto &f : fields) {
            if (!f.h) continue;
            amaru_state_permute(m, m->d_io, f.plane0, f.ncomp, false);
            CUDA_CHECK(cudaMemcpyAsync(f.h, m->d_io, (size_t)n * f.ncomp * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
            CUDA_CHECK(cudaStreamSynchronize(m->stream));
        }
        return AMARU_OK;
    });
}

int amaru_state_backup(amaru_model *m) {
    if (!m) return AMARU_ERR_ARG;
    if (m->grp) return amaru_group_simple(m, 0, 0, 0, nullptr, 0);
    cudaSetDevice(m->device);
    const size_t bytes = (size_t)AMARU_NSTATE * m->nip_total * sizeof(double);
    if (cudaMemcpyAsync(m->d_statebk, m->d_state, bytes, cudaMemcpyDeviceToDevice, m->stream) != cudaSuccess) return AMARU_ERR_CUDA;
    return AMARU_OK;
}

int amaru_state_restore(amaru_model *m) {
    if (!m) return AMARU_ERR_ARG;
    if (m->grp) return amaru_group_simple(m, 1, 0, 0, nullptr, 0);
    cudaSetDevice(m->device);
    const size_t bytes = (size_t)AMARU_NSTATE * m->nip_total * sizeof(double);
    if (cudaMemcpyAsync(m->d_state, m->d_statebk, bytes, cudaMemcpyDeviceToDevice, m->stream) != cudaSuccess) return AMARU_ERR_CUDA;
    return AMARU_OK;
}

int amaru_assemble_K(amaru_model *m, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m, AMARU_ERR_ARG, "null handle");
        if (m->grp) return amaru_group_simple(m, 2, 0, 0, msg, msglen);
        use_device(m);
        reset_status(m);
        amaru_launch_assemble(m, 0);
        m->blended = false;
        amaru_combine_matrix(m);
        amaru_ebe_refresh(m);
        if (m->nranks > 1) amaru_allreduce_max_int(m, m->d_status);
        const int st = read_status(m);
        if (st) throw AmaruError{st, status_text(st)};
        return AMARU_OK;
    });
}

int amaru_tangent_save(amaru_model *m, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m, AMARU_ERR_ARG, "null handle");
        if (m->grp) return amaru_group_simple(m, 6, 0, 0, msg, msglen);
        use_device(m);
        const size_t bytes = (size_t)m->nblk * m->nd * m->nd * sizeof(double);
        if (!m->d_Ksave) CUDA_CHECK(cudaMalloc(&m->d_Ksave, bytes + 256));
        CUDA_CHECK(cudaMemcpyAsync(m->d_Ksave, m->d_K, bytes, cudaMemcpyDeviceToDevice, m->stream));
        return AMARU_OK;
    });
}

int amaru_tangent_blend(amaru_model *m, double a1, double a2, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m, AMARU_ERR_ARG, "null handle");
        if (m->grp) return amaru_group_simple(m, 7, a1, a2, msg, msglen);
        use_device(m);
        AMARU_REQUIRE(m->d_Ksave != nullptr, AMARU_ERR_ARG, "amaru_tangent_blend: call amaru_tangent_save first");
        if (a1 == 0.0 && a2 == 1.0) return AMARU_OK;   // :BE — K2 itself, the per-IP tangent data still describes it
        amaru_axpby(m, m->nblk * m->nd * m->nd, a1, m->d_Ksave, a2, m->d_K, m->d_K);
        m->blended = true;
        amaru_combine_matrix(m);
        return AMARU_OK;
    });
}

int amaru_assemble_M(amaru_model *m, const double *rho, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m && rho, AMARU_ERR_ARG, "null argument");
        if (m->grp) return amaru_group_assemble_M(m, rho, msg, msglen);
        use_device(m);
        reset_status(m);
        if (!m->d_M) CUDA_CHECK(cudaMalloc(&m->d_M, (size_t)m->nblk * m->nd * m->nd * sizeof(double) + 256));
        for (Batch &B : m->batches) {
            std::vector<int64_t> perm((size_t)B.nelem);
            CUDA_CHECK(cudaMemcpy(perm.data(), B.d_perm, perm.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
            std::vector<double> r((size_t)B.nelem);
            for (int64_t s = 0; s < B.nelem; s++) r[(size_t)s] = rho[B.elem_off + perm[(size_t)s]];
            if (B.d_rho) cudaFree(B.d_rho);
            B.d_rho = upload(r.data(), r.size());
        }
        amaru_launch_assemble(m, 1);
        if (m->sysB != 0.0) amaru_combine_matrix(m);    // a*K + b*M in use: keep it in step with the new M
        if (m->nranks > 1) amaru_allreduce_max_int(m, m->d_status);   // every rank takes the same exit
        const int st = read_status(m);
        if (st) throw AmaruError{st, status_text(st)};
        return AMARU_OK;
    });
}

int amaru_set_system_matrix(amaru_model *m, double a, double b, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m, AMARU_ERR_ARG, "null handle");
        if (m->grp) return amaru_group_simple(m, 3, a, b, msg, msglen);
        use_device(m);
        m->sysA = a;
        m->sysB = b;
        amaru_combine_matrix(m);
        CUDA_CHECK(cudaStreamSynchronize(m->stream));
        return AMARU_OK;
    });
}

int amaru_get_csr(amaru_model *m, int64_t *rowptr, int32_t *colind, double *val, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m && rowptr && colind, AMARU_ERR_ARG, "null argument");
        AMARU_REQUIRE(m->nranks == 1 && !m->grp, AMARU_ERR_UNSUPPORTED, "amaru_get_csr: single-GPU handles only");
        use_device(m);
        const int bs = m->nd, b2 = bs * bs;
        std::vector<double> K;
        if (val) {
            K.resize((size_t)m->nblk * b2);
            CUDA_CHECK(cudaStreamSynchronize(m->stream));
            CUDA_CHECK(cudaMemcpy(K.data(), m->d_A, K.size() * sizeof(double), cudaMemcpyDeviceToHost));
        }
        const int64_t n = m->ndofs;
        std::vector<int64_t> cnt((size_t)n + 1, 0);
        for (int64_t A = 0; A < m->nowned; A++) {
            const int64_t nb = m->h_rowptr[A + 1] - m->h_rowptr[A];
            for (int r = 0; r < bs; r++) cnt[(size_t)m->h_eqid[A * bs + r] + 1] = nb * bs;
        }
        for (int64_t i = 0; i < n; i++) cnt[i + 1] += cnt[i];
        std::memcpy(rowptr, cnt.data(), (size_t)(n + 1) * sizeof(int64_t));
        std::vector<std::pair<int32_t, double>> row;
        for (int64_t A = 0; A < m->nowned; A++) {
            for (int r = 0; r < bs; r++) {
                row.clear();
                for (int32_t k = m->h_rowptr[A]; k < m->h_rowptr[A + 1]; k++) {
                    const int64_t B = m->h_col[k];
                    for (int c = 0; c < bs; c++)
                        row.emplace_back(m->h_eqid[B * bs + c], val ? K[(size_t)k * b2 + r * bs + c] : 0.0);
                }
                std::sort(row.begin(), row.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
                int64_t o = cnt[(size_t)m->h_eqid[A * bs + r]];
                for (auto &e : row) {
                    colind[o] = e.first;
                    if (val) val[o] = e.second;
                    o++;
                }
            }
        }
        return AMARU_OK;
    });
}

int amaru_solve(amaru_model *m, double *U, double *F, double cg_rtol, int cg_maxit, int precond, int *iters,
                double *relres, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m && U && F, AMARU_ERR_ARG, "null argument");
        AMARU_REQUIRE(cg_rtol > 0 && cg_maxit > 0, AMARU_ERR_ARG, "amaru_solve: cg_rtol and cg_maxit must be > 0");
        AMARU_REQUIRE(precond == AMARU_PRECOND_JACOBI || precond == AMARU_PRECOND_BLOCK_JACOBI, AMARU_ERR_ARG, "bad preconditioner");
        if (m->grp) return amaru_group_solve(m, U, F, cg_rtol, cg_maxit, precond, iters, relres, msg, msglen);
        use_device(m);
        const size_t bytes = (size_t)m->ndofs * sizeof(double);
        CUDA_CHECK(cudaMemcpyAsync(m->d_U, U, bytes, cudaMemcpyHostToDevice, m->stream));
        CUDA_CHECK(cudaMemcpyAsync(m->d_F, F, bytes, cudaMemcpyHostToDevice, m->stream));
        SolveInfo info;
        const int st = solve_device(m, cg_rtol, cg_maxit, precond, info);
        if (iters) *iters = info.iters;
        if (relres) *relres = info.relres;
        if (st == AMARU_FAIL_CG_NOCONV || st == AMARU_FAIL_SINGULAR) throw AmaruError{st, status_text(st)};
        if (m->nu > 0) CUDA_CHECK(cudaMemcpyAsync(U, m->d_U, (size_t)m->nu * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
        if (m->ndofs > m->nu)
            CUDA_CHECK(cudaMemcpyAsync(F + m->nu, m->d_F + m->nu, (size_t)(m->ndofs - m->nu) * sizeof(double),
                                       cudaMemcpyDeviceToHost, m->stream));
        CUDA_CHECK(cudaStreamSynchronize(m->stream));
        return st;
    });
}

int amaru_update_state(amaru_model *m, const double *dU, double *dFin, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m && dU && dFin, AMARU_ERR_ARG, "null argument");
        if (m->grp) return amaru_group_update(m, dU, dFin, 0, msg, msglen);
        use_device(m);
        const size_t bytes = (size_t)m->ndofs * sizeof(double);
        CUDA_CHECK(cudaMemcpyAsync(m->d_U, dU, bytes, cudaMemcpyHostToDevice, m->stream));
        const int st = update_device(m);
        CUDA_CHECK(cudaMemcpyAsync(dFin, m->d_F, bytes, cudaMemcpyDeviceToHost, m->stream));
        CUDA_CHECK(cudaStreamSynchronize(m->stream));
        if (st) throw AmaruError{st, status_text(st)};
        return AMARU_OK;
    });
}

int amaru_internal_forces(amaru_model *m, double *Fin, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m && Fin, AMARU_ERR_ARG, "null argument");
        if (m->grp) return amaru_group_update(m, nullptr, Fin, 1, msg, msglen);
        use_device(m);
        amaru_launch_update(m, m->d_x, m->d_f, 1);
        CUDA_CHECK(cudaMemsetAsync(m->d_F, 0, (size_t)m->ndofs * sizeof(double), m->stream));
        amaru_nodes_to_eq(m, m->d_f, m->d_F, 0);
        if (m->nranks > 1) amaru_allreduce_sum(m, m->d_F, m->ndofs);   // every rank contributes the rows it owns
        CUDA_CHECK(cudaMemcpyAsync(Fin, m->d_F, (size_t)m->ndofs * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
        CUDA_CHECK(cudaStreamSynchronize(m->stream));
        return AMARU_OK;
    });
}

int amaru_matvec(amaru_model *m, double a, double b, const double *x, double *y, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m && x && y, AMARU_ERR_ARG, "null argument");
        if (m->grp) return amaru_group_product(m, 0, a, b, x, y, 0, nullptr, msg, msglen);
        use_device(m);
        AMARU_REQUIRE(b == 0.0 || m->d_M != nullptr, AMARU_ERR_ARG, "amaru_matvec: mass matrix not assembled");
        // y = a*(K x) + b*(M x): two products on the stored matrices, no combined matrix is formed
        const size_t bytes = (size_t)m->ndofs * sizeof(double);
        CUDA_CHECK(cudaMemcpyAsync(m->d_U, x, bytes, cudaMemcpyHostToDevice, m->stream));
        amaru_eq_to_nodes(m, m->d_U, m->d_x);
        if (m->nranks > 1) amaru_halo_exchange(m, m->d_x);
        amaru_spmv(m, m->d_K, m->d_x, m->d_q, 0);
        if (b != 0.0) amaru_spmv(m, m->d_M, m->d_x, m->d_r, 0);
        amaru_axpby(m, m->nowned * m->nd, a, m->d_q, b, b != 0.0 ? m->d_r : m->d_q, m->d_q);
        if (m->nranks == 1) {
            amaru_nodes_to_eq(m, m->d_q, m->d_F, 0);
        } else {
            CUDA_CHECK(cudaMemsetAsync(m->d_F, 0, bytes, m->stream));
            amaru_nodes_to_eq(m, m->d_q, m->d_F, 0);
            amaru_allreduce_sum(m, m->d_F, m->ndofs);
        }
        CUDA_CHECK(cudaMemcpyAsync(y, m->d_F, bytes, cudaMemcpyDeviceToHost, m->stream));
        CUDA_CHECK(cudaStreamSynchronize(m->stream));
        return AMARU_OK;
    });
}

int amaru_operator_apply(amaru_model *m, const double *x, double *y, int masked, double *pAp, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m && x && y, AMARU_ERR_ARG, "null argument");
        if (m->grp) return amaru_group_product(m, 1, 0, 0, x, y, masked, pAp, msg, msglen);
        use_device(m);
        const size_t bytes = (size_t)m->ndofs * sizeof(double);
        CUDA_CHECK(cudaMemcpyAsync(m->d_U, x, bytes, cudaMemcpyHostToDevice, m->stream));
        amaru_eq_to_nodes(m, m->d_U, m->d_p);
        const double pq = amaru_operator_product(m, masked);
        if (pAp) *pAp = pq;
        CUDA_CHECK(cudaMemsetAsync(m->d_F, 0, bytes, m->stream));
        amaru_nodes_to_eq(m, m->d_q, m->d_F, 0);
        if (m->nranks > 1) amaru_allreduce_sum(m, m->d_F, m->ndofs);
        CUDA_CHECK(cudaMemcpyAsync(y, m->d_F, bytes, cudaMemcpyDeviceToHost, m->stream));
        CUDA_CHECK(cudaStreamSynchronize(m->stream));
        return AMARU_OK;
    });
}

int amaru_set_device_vectors(amaru_model *m, const double *U, const double *F, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m && U && F, AMARU_ERR_ARG, "null argument");
        if (m->grp) return amaru_group_set_device_vectors(m, U, F, msg, msglen);
        use_device(m);
        const size_t bytes = (size_t)m->ndofs * sizeof(double);
        if (!m->d_U0) CUDA_CHECK(cudaMalloc(&m->d_U0, bytes));
        if (!m->d_F0) CUDA_CHECK(cudaMalloc(&m->d_F0, bytes));
        CUDA_CHECK(cudaMemcpyAsync(m->d_U0, U, bytes, cudaMemcpyHostToDevice, m->stream));
        CUDA_CHECK(cudaMemcpyAsync(m->d_F0, F, bytes, cudaMemcpyHostToDevice, m->stream));
        CUDA_CHECK(cudaStreamSynchronize(m->stream));
        return AMARU_OK;
    });
}

int amaru_newton_iteration_device(amaru_model *m, double cg_rtol, int cg_maxit, int precond, double *phase_ms, int *iters,
                                  double *relres, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        if (m && m->grp) return amaru_group_newton_iteration(m, cg_rtol, cg_maxit, precond, phase_ms, iters, relres, msg, msglen);
        AMARU_REQUIRE(m && m->d_U0 && m->d_F0, AMARU_ERR_ARG, "amaru_newton_iteration_device: call amaru_set_device_vectors first");
        use_device(m);
        const size_t bytes = (size_t)m->ndofs * sizeof(double);
        EventSet<4> evs;
        cudaEvent_t *ev = evs.ev;
        CUDA_CHECK(cudaEventRecord(ev[0], m->stream));
        reset_status(m);
        amaru_launch_assemble(m, 0);                                   // mount_K            (mech-solver.jl:327)
        m->blended = false;
        amaru_combine_matrix(m);
        amaru_ebe_refresh(m);
        if (m->nranks > 1) amaru_allreduce_max_int(m, m->d_status);
        const int as = read_status(m);                                 // assembly status (update_device resets the flag)
        CUDA_CHECK(cudaEventRecord(ev[1], m->stream));
        CUDA_CHECK(cudaMemcpyAsync(m->d_U, m->d_U0, bytes, cudaMemcpyDeviceToDevice, m->stream));
        CUDA_CHECK(cudaMemcpyAsync(m->d_F, m->d_F0, bytes, cudaMemcpyDeviceToDevice, m->stream));
        SolveInfo info;
        int st = solve_device(m, cg_rtol, cg_maxit, precond, info);    // solve_system!      (:331)
        CUDA_CHECK(cudaEventRecord(ev[2], m->stream));
        amaru_state_restore(m);                                        // copyto!(State,Bk)  (:333)
        const int st2 = update_device(m);                              // update_state!      (:335)
        CUDA_CHECK(cudaEventRecord(ev[3], m->stream));
        CUDA_CHECK(cudaEventSynchronize(ev[3]));
        if (phase_ms) {
            float t;
            for (int i = 0; i < 3; i++) {
                CUDA_CHECK(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
                phase_ms[i] = t;
            }
            CUDA_CHECK(cudaEventElapsedTime(&t, ev[0], ev[3]));
            phase_ms[3] = t;
        }
        if (iters) *iters = info.iters;
        if (relres) *relres = info.relres;
        if (as) st = as;
        if (st == AMARU_OK) st = st2;
        if (st) throw AmaruError{st, status_text(st)};
        return AMARU_OK;
    });
}

int amaru_time_kernel(amaru_model *m, int kind, int precond, int reps, double *avg_ms, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m && avg_ms && reps > 0, AMARU_ERR_ARG, "bad argument");
        AMARU_REQUIRE(!m->grp, AMARU_ERR_UNSUPPORTED, "amaru_time_kernel: single-GPU handles only");
        use_device(m);
        EventSet<2> evs;
        cudaEvent_t e0 = evs.ev[0], e1 = evs.ev[1];
        if (kind >= 3 || kind == 0) amaru_time_cg_kernel(m, kind, precond, 1);   // warm-up
        CUDA_CHECK(cudaEventRecord(e0, m->stream));
        if (kind == 1) {
            for (int i = 0; i < reps; i++) amaru_launch_assemble(m, 0);
        } else if (kind == 2) {
            for (int i = 0; i < reps; i++) amaru_launch_update(m, m->d_x, m->d_f, 1);
        } else {
            amaru_time_cg_kernel(m, kind, precond, reps);
        }
        CUDA_CHECK(cudaEventRecord(e1, m->stream));
        CUDA_CHECK(cudaEventSynchronize(e1));
        float t = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&t, e0, e1));
        *avg_ms = (double)t / reps;
        return AMARU_OK;
    });
}

int amaru_set_operator(amaru_model *m, int kind) {
    if (!m || (kind != AMARU_OPERATOR_CSR && kind != AMARU_OPERATOR_EBE)) return AMARU_ERR_ARG;
    if (kind == AMARU_OPERATOR_EBE && m->stressmodel == AMARU_STRESS_AXISYMMETRIC) return AMARU_ERR_UNSUPPORTED;
    if (m->grp) return amaru_group_simple(m, 4, kind, 0, nullptr, 0);
    m->op_ebe = kind == AMARU_OPERATOR_EBE;
    return AMARU_OK;
}

int amaru_set_profiling(amaru_model *m, int on) {
    if (!m) return AMARU_ERR_ARG;
    if (m->grp) return amaru_group_simple(m, 5, on, 0, nullptr, 0);
    m->profiling = on != 0;
    m->prof_spmv_ms = 0.0;
    m->prof_spmv_n = 0;
    return AMARU_OK;
}

int amaru_get_profile(amaru_model *m, double *spmv_ms_total, int64_t *spmv_launches) {
    if (!m) return AMARU_ERR_ARG;
    if (m->grp) m = amaru_group_part(m, 0);
    if (spmv_ms_total) *spmv_ms_total = m->prof_spmv_ms;
    if (spmv_launches) *spmv_launches = m->prof_spmv_n;
    return AMARU_OK;
}

int amaru_comm_selftest(amaru_model *m, int skip_rank, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m && m->grp, AMARU_ERR_UNSUPPORTED, "amaru_comm_selftest: multi-GPU handles (ngpus > 1) only");
        return amaru_group_comm_selftest(m, skip_rank, msg, msglen);
    });
}

}  // extern "C"
