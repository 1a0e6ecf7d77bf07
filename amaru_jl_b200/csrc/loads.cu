// K11: natural boundary conditions integrated on the device — the step right before the Newton loop.
// Replaces mech_boundary_forces / mech_solid_body_forces (reference src/mech/elem/distributed.jl:76-152,157-217) and
// the `F[map] += Fd` loops of setup_bc!/compute_bc_vals! (src/bc.jl:116-136,175-194).
//
// One load set = the entities (facets or cells) one boundary condition selected.  Two kernels per application:
//   k_load_forces<NN,CD,ND>  one thread per entity: gathers the entity's coordinates, loops over the default quadrature
//                            of the shape (J = C'D, coef = norm2(J)|det(J) * w * th, Q from key/value), keeps the
//                            NN x ND nodal forces in registers and writes them entity-major;
//   k_load_gather            one thread per distinct node of the set: adds the entity contributions in ascending entity
//                            order into F (eq_id order) — the same accumulation order as the reference's serial loop, no
//                            atomics, bitwise deterministic.
// The gather lists (node -> [entity, local node]) are built once on the host when the set is created.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "amaru_internal.h"

struct amaru_loadset {
    amaru_model *m = nullptr;
    int shape = 0, nn = 0, cd = 0, nd = 0, nip = 0;
    int64_t nents = 0, nuniq = 0;
    int32_t *d_nodes = nullptr;   // [nents*nn]
    double *d_N = nullptr;        // [nip*nn]
    double *d_D = nullptr;        // [nip*nn*cd]
    double *d_w = nullptr;        // [nip]
    double *d_Fd = nullptr;       // [nents*nn*nd]
    double *d_vip = nullptr;      // [nents*nip]
    int64_t *d_gptr = nullptr;    // [nuniq+1]
    int32_t *d_gnode = nullptr;   // [nuniq]
    int32_t *d_gsrc = nullptr;    // [nents*nn]  entity*nn + local node, grouped by node, ascending entity
};

namespace {

template <int NN, int CD, int ND>
__global__ void __launch_bounds__(128)
k_load_forces(int64_t nents, int nip, const int32_t *__restrict__ nodes, const double *__restrict__ coords,
              const double *__restrict__ Ntab, const double *__restrict__ Dtab, const double *__restrict__ wtab, double th,
              int key, double cval, const double *__restrict__ vip, double *__restrict__ Fd) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nents; e += (int64_t)gridDim.x * blockDim.x) {
        double C[NN][ND], F[NN][ND];
#pragma unroll
        for (int a = 0; a < NN; a++) {
            const int64_t n = nodes[e * NN + a];
#pragma unroll
            for (int i = 0; i < ND; i++) {
                C[a][i] = coords[n * 3 + i];
                F[a][i] = 0.0;
            }
        }
        for (int q = 0; q < nip; q++) {
            const double *N = Ntab + q * NN, *D = Dtab + q * NN * CD;
            double J[ND][CD];   // J = C'D  (distributed.jl:122,193)
#pragma unroll
            for (int i = 0; i < ND; i++)
#pragma unroll
                for (int j = 0; j < CD; j++) {
                    double s = 0.0;
#pragma unroll
                    for (int a = 0; a < NN; a++) s += C[a][i] * D[a * CD + j];
                    J[i][j] = s;
                }
            double nrm[3] = {0.0, 0.0, 0.0}, jac;
            if constexpr (CD == ND && ND == 2) {          // det(J) (distributed.jl:211)
                jac = J[0][0] * J[1][1] - J[0][1] * J[1][0];
            } else if constexpr (CD == ND && ND == 3) {
                jac = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                      J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
            } else if constexpr (CD == 1) {               // norm2 of a column = its 2-norm (tools/linalg.jl:56); n = [J2,-J1] (distributed.jl:136)
                jac = sqrt(J[0][0] * J[0][0] + J[1][0] * J[1][0]);
                nrm[0] = J[1][0];
                nrm[1] = -J[0][0];
            } else {                                      // 3x2: n = J[:,1] x J[:,2] (distributed.jl:138); norm2 = |minors| (tools/linalg.jl:64-69)
                nrm[0] = J[1][0] * J[2][1] - J[2][0] * J[1][1];
                nrm[1] = J[2][0] * J[0][1] - J[0][0] * J[2][1];
                nrm[2] = J[0][0] * J[1][1] - J[1][0] * J[0][1];
                const double j1 = J[0][0] * J[1][1] - J[0][1] * J[1][0];
                const double j2 = J[0][0] * J[2][1] - J[0][1] * J[2][0];
                const double j3 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
                jac = sqrt(j1 * j1 + j2 * j2 + j3 * j3);
            }
            const double v = vip ? vip[e * nip + q] : cval;
            double Q[ND];
#pragma unroll
            for (int i = 0; i < ND; i++) Q[i] = 0.0;
            if (key == AMARU_LOAD_NORMAL) {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < ND; i++) s += nrm[i] * nrm[i];
                s = sqrt(s);
#pragma unroll
                for (int i = 0; i < ND; i++) Q[i] = v * (nrm[i] / s);
            } else {
#pragma unroll
                for (int i = 0; i < ND; i++)
                    if (i == key) Q[i] = v;
            }
            const double coef = jac * wtab[q] * th;
#pragma unroll
            for (int a = 0; a < NN; a++)
#pragma unroll
                for (int i = 0; i < ND; i++) F[a][i] += coef * N[a] * Q[i];
        }
#pragma unroll
        for (int a = 0; a < NN; a++)
#pragma unroll
            for (int i = 0; i < ND; i++) Fd[(e * NN + a) * ND + i] = F[a][i];
    }
}

// X = C'N at every integration point of every entity (distributed.jl:123,194)
template <int NN, int ND>
__global__ void k_load_ipcoords(int64_t nents, int nip, const int32_t *__restrict__ nodes, const double *__restrict__ coords,
                                const double *__restrict__ Ntab, double *__restrict__ X) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nents * nip; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / nip;
        const int q = (int)(t - e * nip);
        double x[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int a = 0; a < NN; a++) {
            const int64_t n = nodes[e * NN + a];
            const double Na = Ntab[q * NN + a];
#pragma unroll
            for (int i = 0; i < ND; i++) x[i] += coords[n * 3 + i] * Na;
        }
        X[t * 3 + 0] = x[0];
        X[t * 3 + 1] = x[1];
        X[t * 3 + 2] = x[2];
    }
}

__global__ void k_load_gather(int64_t nuniq, const int64_t *__restrict__ gptr, const int32_t *__restrict__ gnode,
                              const int32_t *__restrict__ gsrc, int nd, const double *__restrict__ Fd,
                              const int32_t *__restrict__ eqid, double *F) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nuniq * nd; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = t / nd;
        const int d = (int)(t - u * nd);
        const int64_t dst = eqid[(int64_t)gnode[u] * nd + d];
        double s = F[dst];
        for (int64_t k = gptr[u]; k < gptr[u + 1]; k++) s += Fd[(int64_t)gsrc[k] * nd + d];
        F[dst] = s;
    }
}

void set_msg(char *msg, int msglen, const std::string &s) {
    if (msg && msglen > 0) std::snprintf(msg, (size_t)msglen, "%s", s.c_str());
}
template <class Fn>
int guarded(char *msg, int msglen, Fn f) {
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
int grid_for(const amaru_model *m, int64_t n, int threads) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, (int64_t)m->nsm * 16));
}

void launch_forces(amaru_loadset *ls, int key, double cval, const double *d_vip) {
    amaru_model *m = ls->m;
    const int g = grid_for(m, ls->nents, 128);
#define LOADCASE(NN, CD, ND)                                                                                         \
    k_load_forces<NN, CD, ND><<<g, 128, 0, m->stream>>>(ls->nents, ls->nip, ls->d_nodes, m->d_coords, ls->d_N, ls->d_D, \
                                                         ls->d_w, m->th, key, cval, d_vip, ls->d_Fd)
    switch (ls->shape) {
    case AMARU_SHAPE_LIN2: LOADCASE(2, 1, 2); break;
    case AMARU_SHAPE_LIN3: LOADCASE(3, 1, 2); break;
    case AMARU_SHAPE_TRI6: LOADCASE(6, 2, 3); break;
    case AMARU_SHAPE_QUAD4:
        if (ls->nd == 3) LOADCASE(4, 2, 3);
        else LOADCASE(4, 2, 2);
        break;
    case AMARU_SHAPE_QUAD8:
        if (ls->nd == 3) LOADCASE(8, 2, 3);
        else LOADCASE(8, 2, 2);
        break;
    case AMARU_SHAPE_HEX8: LOADCASE(8, 3, 3); break;
    case AMARU_SHAPE_HEX20: LOADCASE(20, 3, 3); break;
    case AMARU_SHAPE_TET10: LOADCASE(10, 3, 3); break;
    }
#undef LOADCASE
    m->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void launch_ipcoords(amaru_loadset *ls, double *d_X) {
    amaru_model *m = ls->m;
    const int g = grid_for(m, ls->nents * ls->nip, 256);
#define IPCASE(NN, ND) k_load_ipcoords<NN, ND><<<g, 256, 0, m->stream>>>(ls->nents, ls->nip, ls->d_nodes, m->d_coords, ls->d_N, d_X)
    switch (ls->shape) {
    case AMARU_SHAPE_LIN2: IPCASE(2, 2); break;
    case AMARU_SHAPE_LIN3: IPCASE(3, 2); break;
    case AMARU_SHAPE_TRI6: IPCASE(6, 3); break;
    case AMARU_SHAPE_QUAD4:
        if (ls->nd == 3) IPCASE(4, 3);
        else IPCASE(4, 2);
        break;
    case AMARU_SHAPE_QUAD8:
        if (ls->nd == 3) IPCASE(8, 3);
        else IPCASE(8, 2);
        break;
    case AMARU_SHAPE_HEX8: IPCASE(8, 3); break;
    case AMARU_SHAPE_HEX20: IPCASE(20, 3); break;
    case AMARU_SHAPE_TET10: IPCASE(10, 3); break;
    }
#undef IPCASE
    m->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void free_loadset(amaru_loadset *ls) {
    if (!ls) return;
    if (ls->m) cudaSetDevice(ls->m->device);
    for (void *p : {(void *)ls->d_nodes, (void *)ls->d_N, (void *)ls->d_D, (void *)ls->d_w, (void *)ls->d_Fd, (void *)ls->d_vip,
                    (void *)ls->d_gptr, (void *)ls->d_gnode, (void *)ls->d_gsrc})
        cudaFree(p);
    delete ls;
}

}  // namespace

extern "C" {

int amaru_loadset_create(amaru_model *m, int shape, int64_t nents, const int32_t *nodes, amaru_loadset **out, char *msg,
                         int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m && out, AMARU_ERR_ARG, "amaru_loadset_create: null argument");
        *out = nullptr;
        AMARU_REQUIRE(nents >= 0 && (nents == 0 || nodes), AMARU_ERR_ARG, "amaru_loadset_create: null node list");
        AMARU_REQUIRE(m->stressmodel != AMARU_STRESS_AXISYMMETRIC, AMARU_ERR_UNSUPPORTED,
                      "amaru_loadset_create: axisymmetric loads (th = 2*pi*r, distributed.jl:121,193) are integrated by the host glue");
        // a multi-GPU handle of one process integrates on its first GPU with the global arrays its wrapper keeps (group.cu)
        AMARU_REQUIRE(m->nranks == 1 || m->grp, AMARU_ERR_UNSUPPORTED,
                      "amaru_loadset_create: rank-level partitioned handles integrate loads per rank on the host");
        ShapeInfo si;
        AMARU_REQUIRE(amaru_shape_info(shape, si), AMARU_ERR_UNSUPPORTED, "amaru_loadset_create: unknown shape");
        // facets: shape dimension ndim-1 (distributed.jl:92-93 rejects surfaces loads on 3D edges); cells: dimension ndim
        AMARU_REQUIRE(si.nd == m->nd || si.nd == m->nd - 1, AMARU_ERR_ARG,
                      "mech_boundary_forces: shape dimension does not fit a facet or a cell of this analysis");
        AMARU_REQUIRE(nents * si.nn < 2147483647LL, AMARU_ERR_ARG, "amaru_loadset_create: too many entities");
        for (int64_t i = 0; i < nents * si.nn; i++)
            AMARU_REQUIRE(nodes[i] >= 0 && nodes[i] < m->nnodes, AMARU_ERR_ARG, "amaru_loadset_create: node id out of range");
        CUDA_CHECK(cudaSetDevice(m->device));
        amaru_loadset *ls = new amaru_loadset();
        try {
            ls->m = m;
            ls->shape = shape; ls->nn = si.nn; ls->cd = si.nd; ls->nd = m->nd; ls->nip = si.nip;
            ls->nents = nents;
            ls->d_nodes = upload(nodes, (size_t)nents * si.nn);
            ls->d_N = upload(si.N.data(), si.N.size());
            ls->d_D = upload(si.dNdR.data(), si.dNdR.size());
            std::vector<double> w((size_t)si.nip);
            for (int q = 0; q < si.nip; q++) w[q] = si.ips[4 * q + 3];
            ls->d_w = upload(w.data(), w.size());
            CUDA_CHECK(cudaMalloc(&ls->d_Fd, std::max<size_t>((size_t)nents * si.nn * m->nd, 1) * sizeof(double)));
            CUDA_CHECK(cudaMalloc(&ls->d_vip, std::max<size_t>((size_t)nents * si.nip, 1) * sizeof(double)));
            // gather lists: counting sort of (node, entity*nn+a) by node keeps ascending entity order inside a node
            std::vector<int64_t> cnt((size_t)m->nnodes + 1, 0);
            const int64_t tot = nents * si.nn;
            for (int64_t i = 0; i < tot; i++) cnt[(size_t)nodes[i] + 1]++;
            std::vector<int32_t> gnode;
            std::vector<int64_t> gptr(1, 0);
            std::vector<int64_t> start((size_t)m->nnodes, -1);
            int64_t acc = 0;
            for (int64_t n = 0; n < m->nnodes; n++) {
                if (cnt[(size_t)n + 1] == 0) continue;
                start[(size_t)n] = acc;
                acc += cnt[(size_t)n + 1];
                gnode.push_back((int32_t)n);
                gptr.push_back(acc);
            }
            std::vector<int32_t> gsrc((size_t)tot);
            for (int64_t i = 0; i < tot; i++) gsrc[(size_t)start[(size_t)nodes[i]]++] = (int32_t)i;
            ls->nuniq = (int64_t)gnode.size();
            ls->d_gptr = upload(gptr.data(), gptr.size());
            ls->d_gnode = upload(gnode.data(), gnode.size());
            ls->d_gsrc = upload(gsrc.data(), gsrc.size());
        } catch (...) {
            free_loadset(ls);
            throw;
        }
        *out = ls;
        return AMARU_OK;
    });
}

int64_t amaru_loadset_nip(const amaru_loadset *ls) { return ls ? ls->nents * ls->nip : -1; }

int amaru_loadset_ip_coords(amaru_loadset *ls, double *X, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(ls && X, AMARU_ERR_ARG, "amaru_loadset_ip_coords: null argument");
        amaru_model *m = ls->m;
        CUDA_CHECK(cudaSetDevice(m->device));
        const int64_t n = ls->nents * ls->nip;
        if (n == 0) return AMARU_OK;
        double *d_X = nullptr;
        CUDA_CHECK(cudaMalloc(&d_X, (size_t)n * 3 * sizeof(double)));
        try {
            launch_ipcoords(ls, d_X);
            CUDA_CHECK(cudaMemcpyAsync(X, d_X, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
            CUDA_CHECK(cudaStreamSynchronize(m->stream));
        } catch (...) {
            cudaFree(d_X);
            throw;
        }
        cudaFree(d_X);
        return AMARU_OK;
    });
}

int amaru_loadset_apply(amaru_loadset *ls, int key, double cval, const double *vip, double *F, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(ls && F, AMARU_ERR_ARG, "amaru_loadset_apply: null argument");
        amaru_model *m = ls->m;
        const bool facet = ls->cd == ls->nd - 1;
        if (facet) {   // distributed.jl:79-87
            AMARU_REQUIRE(key == AMARU_LOAD_X || key == AMARU_LOAD_Y || key == AMARU_LOAD_Z || key == AMARU_LOAD_NORMAL, AMARU_ERR_ARG,
                          "mech_boundary_forces: boundary condition is not applicable as distributed bc. Suitable keys are tx, ty, tz, tn");
            AMARU_REQUIRE(!(key == AMARU_LOAD_Z && ls->nd == 2), AMARU_ERR_ARG,
                          "mech_boundary_forces: boundary condition tz is not applicable in a 2D analysis");
        } else {       // distributed.jl:163-165
            AMARU_REQUIRE(key == AMARU_LOAD_X || key == AMARU_LOAD_Y || key == AMARU_LOAD_Z, AMARU_ERR_ARG,
                          "mech_solid_body_forces: condition is not applicable as distributed bc. Suitable keys are wx, wy, wz");
            AMARU_REQUIRE(!(key == AMARU_LOAD_Z && ls->nd == 2), AMARU_ERR_ARG,
                          "mech_solid_body_forces: key wz is not applicable in a 2D analysis");
        }
        CUDA_CHECK(cudaSetDevice(m->device));
        if (ls->nents == 0) return AMARU_OK;
        const double *d_vip = nullptr;
        if (vip) {
            CUDA_CHECK(cudaMemcpyAsync(ls->d_vip, vip, (size_t)ls->nents * ls->nip * sizeof(double), cudaMemcpyHostToDevice, m->stream));
            d_vip = ls->d_vip;
        }
        const size_t bytes = (size_t)m->ndofs * sizeof(double);
        CUDA_CHECK(cudaMemcpyAsync(m->d_io, F, bytes, cudaMemcpyHostToDevice, m->stream));
        launch_forces(ls, key, cval, d_vip);
        k_load_gather<<<grid_for(m, ls->nuniq * ls->nd, 256), 256, 0, m->stream>>>(ls->nuniq, ls->d_gptr, ls->d_gnode, ls->d_gsrc,
                                                                                    ls->nd, ls->d_Fd, m->d_eqid, m->d_io);
        m->launches++;
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemcpyAsync(F, m->d_io, bytes, cudaMemcpyDeviceToHost, m->stream));
        CUDA_CHECK(cudaStreamSynchronize(m->stream));
        return AMARU_OK;
    });
}

int amaru_loadset_destroy(amaru_loadset *ls) {
    if (!ls) return AMARU_ERR_ARG;
    free_loadset(ls);
    return AMARU_OK;
}

}  // extern "C"
