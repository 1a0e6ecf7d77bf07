// K12: output side — integration-point fields recovered to the nodes on the device.
// Replaces nodal_patch_recovery (reference src/fe-model.jl:506-692) fed by ip_state_vals / stress_strain_dict
// (src/tools/tensors.jl:162-216; src/mech/mat/von-mises.jl:159-168, drucker-prager.jl:152-163), which in the reference
// builds one OrderedDict per integration point and one pinv per corner-node patch on every output.
//
// Host, once per handle (amaru_recovery_create): corner-node patches exactly as the reference forms them (internal
// patches, boundary patches adopted for orphan nodes with >= 3, 2, 1 elements, fe-model.jl:529-582), patch -> element
// lists and node -> patch lists.
// Device, per output (amaru_recover_nodal):
//   k_ip_coords<NN,ND>  ip.coord = C'N of every integration point (element.jl:160-164), once
//   k_patch_fit<ND>     one warp per patch, lane = field: least-squares fit of the regression polynomial
//                       (reg_terms, fe-model.jl:490-503) through the patch's integration points.  The reference forms
//                       pinv(M) in global coordinates; here the same polynomial space is fitted in coordinates centred at the
//                       patch node and scaled by the patch size (the spaces {1,x,y,z,xy,yz,xz} etc. are invariant under
//                       shift and scaling, so the fitted polynomial is the same one) through the normal equations and a
//                       Cholesky factorisation, which is well conditioned in the scaled variables.
//   k_node_eval         one warp per node, lane = field: evaluates the polynomials of all patches whose elements contain
//                       the node and averages them (the V_vals ./ V_reps of fe-model.jl:664-687), fixed order, no atomics.
// Deviation (documented in DESIGN.md): a rank-deficient patch (possible only for adopted boundary patches of meshes one
// element thick) falls back to the next smaller term count instead of the reference's minimum-norm pinv solution.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>

#include "amaru_internal.h"

namespace {

constexpr int MAXT = 7;       // regression terms
constexpr int NGROUPS = 3;    // 0: stress/strain fields (all materials), 1: ep (VonMises), 2: epa j1 srj2d (DruckerPrager)
constexpr double SR2 = 1.4142135623730951;

struct PatchElem {
    int64_t ip0;   // first integration point (device order)
    int32_t nip;
    int32_t kind;  // AMARU_MAT_*
};

struct Recovery {
    int64_t npatch = 0;
    int nfields = 0;
    int32_t codes[32];
    int group_f0[NGROUPS + 1];
    int group_mask[NGROUPS];
    int32_t *d_codes = nullptr;
    int32_t *d_pcorner = nullptr;
    int64_t *d_pptr = nullptr;
    PatchElem *d_pe = nullptr;
    int64_t *d_nptr = nullptr;
    int32_t *d_npatch = nullptr;
    uint8_t *d_nbits = nullptr;
    double *d_ipx = nullptr;     // [3][nip_total]
    double *d_coef = nullptr;    // [npatch][MAXT][nfields]
    double *d_pcs = nullptr;     // [NGROUPS][npatch][4]  centre xyz, 1/scale
    int32_t *d_pnt = nullptr;    // [NGROUPS][npatch]     terms used (0 = the patch carries no data of the group)
    double *d_V = nullptr;       // [nfields][nnodes]
};

const char *FIELD_NAMES[20] = {"σxx", "σyy", "σzz", "σyz", "σxz", "σxy", "σvm", "σ1", "σ2", "σ3",
                               "εxx", "εyy", "εzz", "εyz", "εxz", "εxy", "ep", "epa", "j1", "srj2d"};

int ncorner(int shape) { return (shape == AMARU_SHAPE_HEX8 || shape == AMARU_SHAPE_HEX20) ? 8 : 4; }

// eigenvalues of the symmetric 3x3 tensor, highest first (tensors.jl:80-93): cyclic Jacobi, accurate to eps*|T|
__device__ void eig3(double a00, double a11, double a22, double a12, double a02, double a01, double *L) {
    for (int sweep = 0; sweep < 8; sweep++) {
        const double off = a01 * a01 + a02 * a02 + a12 * a12;
        if (off == 0.0) break;
        // rotate (0,1), (0,2), (1,2)
#define JROT(app, aqq, apq, arp, arq)                                        \
    if (apq != 0.0) {                                                        \
        const double th = (aqq - app) / (2.0 * apq);                         \
        const double t = copysign(1.0, th) / (fabs(th) + sqrt(th * th + 1.0)); \
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;                 \
        app -= t * apq;                                                      \
        aqq += t * apq;                                                      \
        apq = 0.0;                                                           \
        const double rp = c * arp - s * arq, rq = s * arp + c * arq;         \
        arp = rp;                                                            \
        arq = rq;                                                            \
    }
        JROT(a00, a11, a01, a02, a12)
        JROT(a00, a22, a02, a01, a12)
        JROT(a11, a22, a12, a01, a02)
#undef JROT
    }
    double s1 = a00, s2 = a11, s3 = a22, t;
    if (s1 < s2) { t = s1; s1 = s2; s2 = t; }
    if (s2 < s3) { t = s2; s2 = s3; s3 = t; }
    if (s1 < s2) { t = s1; s1 = s2; s2 = t; }
    L[0] = s1; L[1] = s2; L[2] = s3;
}

__device__ __forceinline__ double j2_of(const double *s) {   // tensors.jl:24-27
    const double t23 = s[3] / SR2, t13 = s[4] / SR2, t12 = s[5] / SR2;
    const double a = s[0] - s[1], b = s[1] - s[2], c = s[2] - s[0];
    return 1.0 / 6.0 * (a * a + b * b + c * c) + t23 * t23 + t13 * t13 + t12 * t12;
}

// one field of ip_state_vals (tensors.jl:162-216, von-mises.jl:165, drucker-prager.jl:158-160)
__device__ double field_value(int code, const double *__restrict__ state, int64_t nipt, int64_t ip) {
    if (code < 0) return 0.0;
    if (code <= 2) return state[(int64_t)code * nipt + ip];
    if (code <= 5) return state[(int64_t)code * nipt + ip] / SR2;
    if (code >= 10 && code <= 12) return state[(int64_t)(code - 4) * nipt + ip];
    if (code >= 13 && code <= 15) return state[(int64_t)(code - 4) * nipt + ip] / SR2;
    if (code == 16 || code == 17) return state[12 * nipt + ip];
    double s[6];
#pragma unroll
    for (int i = 0; i < 6; i++) s[i] = state[(int64_t)i * nipt + ip];
    if (code == 6) return sqrt(3.0 * j2_of(s));
    if (code == 18) return s[0] + s[1] + s[2];
    if (code == 19) return sqrt(j2_of(s));
    double L[3];
    eig3(s[0], s[1], s[2], s[3] / SR2, s[4] / SR2, s[5] / SR2, L);
    return L[code - 7];
}

template <int ND>
__device__ __forceinline__ void basis(const double *xi, double *t) {   // reg_terms (fe-model.jl:490-503), nested prefixes
    t[0] = 1.0;
    t[1] = xi[0];
    t[2] = xi[1];
    if (ND == 3) {
        t[3] = xi[2];
        t[4] = xi[0] * xi[1];
        t[5] = xi[1] * xi[2];
        t[6] = xi[0] * xi[2];
    } else {
        t[3] = xi[0] * xi[1];
        t[4] = xi[0] * xi[0];
        t[5] = xi[1] * xi[1];
        t[6] = 0.0;
    }
}

template <int ND>
__device__ __forceinline__ int nterms_for(int m) {   // fe-model.jl:628-633
    if (ND == 3) return m >= 7 ? 7 : m >= 4 ? 4 : 1;
    return m >= 6 ? 6 : m >= 4 ? 4 : m >= 3 ? 3 : 1;
}
template <int ND>
__device__ __forceinline__ int nterms_below(int nt) {
    if (ND == 3) return nt > 4 ? 4 : 1;
    return nt > 4 ? 4 : nt > 3 ? 3 : 1;
}

template <int ND>
__global__ void __launch_bounds__(128)
k_patch_fit(int64_t npatch, const int64_t *__restrict__ pptr, const PatchElem *__restrict__ pe,
            const int32_t *__restrict__ pcorner, const double *__restrict__ coords, const double *__restrict__ ipx,
            const double *__restrict__ state, int64_t nipt, int kindmask, int f0, int nfg, int nftot,
            const int32_t *__restrict__ codes, double *__restrict__ coef, double *__restrict__ pcs, int32_t *__restrict__ pnt) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int code = lane < nfg ? codes[f0 + lane] : -1;
    for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < npatch; p += nwarps) {
        const int32_t c = pcorner[p];
        double xc[3];
#pragma unroll
        for (int d = 0; d < 3; d++) xc[d] = coords[(int64_t)c * 3 + d];
        // pass 1: number of integration points of the sub-patch and its extent around the patch node
        int m = 0;
        double h = 0.0;
        for (int64_t k = pptr[p]; k < pptr[p + 1]; k++) {
            const PatchElem E = pe[k];
            if (!((1 << E.kind) & kindmask)) continue;
            for (int q = lane; q < E.nip; q += 32)
#pragma unroll
                for (int d = 0; d < ND; d++) h = fmax(h, fabs(ipx[(int64_t)d * nipt + E.ip0 + q] - xc[d]));
            m += E.nip;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) h = fmax(h, __shfl_xor_sync(0xffffffffu, h, o));
        if (m == 0) {
            if (lane == 0) pnt[p] = 0;
            continue;
        }
        const double hinv = h > 0.0 ? 1.0 / h : 1.0;
        // pass 2: normal equations G = M'M (same on every lane) and r = M'W of this lane's field
        double G[MAXT * (MAXT + 1) / 2], r[MAXT];
#pragma unroll
        for (int i = 0; i < MAXT * (MAXT + 1) / 2; i++) G[i] = 0.0;
#pragma unroll
        for (int i = 0; i < MAXT; i++) r[i] = 0.0;
        for (int64_t k = pptr[p]; k < pptr[p + 1]; k++) {
            const PatchElem E = pe[k];
            if (!((1 << E.kind) & kindmask)) continue;
            for (int q = 0; q < E.nip; q++) {
                const int64_t ip = E.ip0 + q;
                double xi[3] = {0.0, 0.0, 0.0}, t[MAXT];
#pragma unroll
                for (int d = 0; d < ND; d++) xi[d] = (ipx[(int64_t)d * nipt + ip] - xc[d]) * hinv;
                basis<ND>(xi, t);
                const double w = field_value(code, state, nipt, ip);
#pragma unroll
                for (int i = 0; i < MAXT; i++) {
                    r[i] += t[i] * w;
#pragma unroll
                    for (int j = 0; j <= i; j++) G[i * (i + 1) / 2 + j] += t[i] * t[j];
                }
            }
        }
        // Cholesky on the leading nt x nt block; a (near-)singular pivot drops to the next smaller basis
        int nt = nterms_for<ND>(m);
        double Lc[MAXT * (MAXT + 1) / 2], a[MAXT];
        while (true) {
            bool ok = true;
            double dmax = 0.0;
            for (int i = 0; i < nt; i++) dmax = fmax(dmax, G[i * (i + 1) / 2 + i]);
            for (int i = 0; i < nt && ok; i++) {
                for (int j = 0; j <= i; j++) {
                    double s = G[i * (i + 1) / 2 + j];
                    for (int k2 = 0; k2 < j; k2++) s -= Lc[i * (i + 1) / 2 + k2] * Lc[j * (j + 1) / 2 + k2];
                    if (i == j) {
                        if (!(s > 1e-11 * dmax)) { ok = false; break; }
                        Lc[i * (i + 1) / 2 + i] = sqrt(s);
                    } else {
                        Lc[i * (i + 1) / 2 + j] = s / Lc[j * (j + 1) / 2 + j];
                    }
                }
            }
            if (ok || nt == 1) break;
            nt = nterms_below<ND>(nt);
        }
        for (int i = 0; i < nt; i++) {   // L y = r
            double s = r[i];
            for (int k2 = 0; k2 < i; k2++) s -= Lc[i * (i + 1) / 2 + k2] * a[k2];
            a[i] = s / Lc[i * (i + 1) / 2 + i];
        }
        for (int i = nt - 1; i >= 0; i--) {   // L' a = y
            double s = a[i];
            for (int k2 = i + 1; k2 < nt; k2++) s -= Lc[k2 * (k2 + 1) / 2 + i] * a[k2];
            a[i] = s / Lc[i * (i + 1) / 2 + i];
        }
        if (lane < nfg)
            for (int i = 0; i < MAXT; i++) coef[((int64_t)p * MAXT + i) * nftot + f0 + lane] = i < nt ? a[i] : 0.0;
        if (lane == 0) {
            pcs[p * 4 + 0] = xc[0]; pcs[p * 4 + 1] = xc[1]; pcs[p * 4 + 2] = xc[2]; pcs[p * 4 + 3] = hinv;
            pnt[p] = nt;
        }
    }
}

template <int ND>
__global__ void __launch_bounds__(128)
k_node_eval(int64_t nnodes, int64_t npatch, const int64_t *__restrict__ nptr, const int32_t *__restrict__ npatchl,
            const uint8_t *__restrict__ nbits, const double *__restrict__ coords, const double *__restrict__ coef,
            const double *__restrict__ pcs, const int32_t *__restrict__ pnt, int nftot, int f1, int f2, double *__restrict__ V) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int g = lane < f1 ? 0 : lane < f2 ? 1 : 2;   // group of this lane's field
    for (int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < nnodes; n += nwarps) {
        double x[3];
#pragma unroll
        for (int d = 0; d < 3; d++) x[d] = coords[n * 3 + d];
        double acc = 0.0;
        int reps = 0;
        for (int64_t k = nptr[n]; k < nptr[n + 1]; k++) {
            const int64_t p = npatchl[k];
            if (!((nbits[k] >> g) & 1)) continue;
            const int nt = pnt[(int64_t)g * npatch + p];
            if (nt == 0 || lane >= nftot) continue;
            const double *cs = pcs + ((int64_t)g * npatch + p) * 4;
            double xi[3] = {0.0, 0.0, 0.0}, t[MAXT];
#pragma unroll
            for (int d = 0; d < ND; d++) xi[d] = (x[d] - cs[d]) * cs[3];
            basis<ND>(xi, t);
            double v = 0.0;
            for (int i = 0; i < nt; i++) v += t[i] * coef[(p * MAXT + i) * nftot + lane];
            acc += v;
            reps++;
        }
        if (lane < nftot) V[(int64_t)lane * nnodes + n] = reps > 0 ? acc / reps : 0.0;   // NaN -> 0 (fe-model.jl:688)
    }
}

template <int NN, int ND>
__global__ void k_ip_coords(int64_t nelem, int nip, const int32_t *__restrict__ conn, const double *__restrict__ coords,
                            const double *__restrict__ Ntab, int64_t ip_off, int64_t nipt, double *__restrict__ ipx) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nelem * nip; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = t / nip;
        const int q = (int)(t - s * nip);
        double x[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int a = 0; a < NN; a++) {
            const int64_t nd = conn[s * NN + a];
            const double Na = Ntab[q * NN + a];
#pragma unroll
            for (int d = 0; d < ND; d++) x[d] += coords[nd * 3 + d] * Na;
        }
#pragma unroll
        for (int d = 0; d < 3; d++) ipx[(int64_t)d * nipt + ip_off + t] = x[d];
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
T *upload(const std::vector<T> &h) {
    T *d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(T)));
    if (!h.empty()) CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}
template <class F>
void parallel_chunks(int64_t n, F f) {
    int nt = amaru_host_threads();
    if (n < 8192) nt = 1;
    if (nt == 1) {
        f(0, n);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([=] { f(n * t / nt, n * (t + 1) / nt); });
    for (auto &x : th) x.join();
}

void free_recovery(Recovery *r) {
    if (!r) return;
    for (void *p : {(void *)r->d_codes, (void *)r->d_pcorner, (void *)r->d_pptr, (void *)r->d_pe, (void *)r->d_nptr,
                    (void *)r->d_npatch, (void *)r->d_nbits, (void *)r->d_ipx, (void *)r->d_coef, (void *)r->d_pcs,
                    (void *)r->d_pnt, (void *)r->d_V})
        cudaFree(p);
    delete r;
}

}  // namespace

void amaru_recovery_destroy(amaru_model *m) {
    free_recovery(static_cast<Recovery *>(m->recovery));
    m->recovery = nullptr;
}

extern "C" {

int amaru_recovery_create(amaru_model *m, const uint8_t *at_bound, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m && at_bound, AMARU_ERR_ARG, "amaru_recovery_create: null argument");
        AMARU_REQUIRE(m->nranks == 1 || m->grp, AMARU_ERR_UNSUPPORTED, "amaru_recovery_create: not for rank-level partitioned handles");
        CUDA_CHECK(cudaSetDevice(m->device));
        CUDA_CHECK(cudaStreamSynchronize(m->stream));
        amaru_recovery_destroy(m);
        Recovery *r = new Recovery();
        m->recovery = r;
        const int64_t nnodes = m->nnodes;
        const int nb = (int)m->batches.size();
        // host copies in ABI element order: connectivity, device position and material kind of every element
        std::vector<int32_t> mat_kind((size_t)m->nmats);
        CUDA_CHECK(cudaMemcpy(mat_kind.data(), m->d_mat_kind, mat_kind.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        std::vector<std::vector<int32_t>> conn(nb);          // [elem (ABI order)][nn]
        std::vector<std::vector<int64_t>> pos(nb);           // ABI element -> colour-sorted position
        std::vector<std::vector<int32_t>> kind(nb);
        for (int b = 0; b < nb; b++) {
            Batch &B = m->batches[b];
            std::vector<int32_t> sconn((size_t)B.nelem * B.nn), smat((size_t)B.nelem);
            std::vector<int64_t> perm((size_t)B.nelem);
            CUDA_CHECK(cudaMemcpy(sconn.data(), B.d_conn, sconn.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
            CUDA_CHECK(cudaMemcpy(smat.data(), B.d_emat, smat.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
            CUDA_CHECK(cudaMemcpy(perm.data(), B.d_perm, perm.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
            conn[b].resize(sconn.size());
            pos[b].resize((size_t)B.nelem);
            kind[b].resize((size_t)B.nelem);
            for (int64_t s = 0; s < B.nelem; s++) {
                const int64_t e = perm[(size_t)s];
                std::memcpy(&conn[b][(size_t)e * B.nn], &sconn[(size_t)s * B.nn], sizeof(int32_t) * B.nn);
                pos[b][(size_t)e] = s;
                kind[b][(size_t)e] = mat_kind[(size_t)smat[(size_t)s]];
            }
        }
        // fields, ordered like the OrderedSet of fe-model.jl:588-596: stress/strain fields, then the extra fields of each
        // material kind in order of first appearance
        {
            static const int d3[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};
            static const int ps[13] = {0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 15};
            const bool plane = m->stressmodel == AMARU_STRESS_PLANESTRAIN;
            int n = 0;
            for (int i = 0; i < (plane ? 13 : 16); i++) r->codes[n++] = plane ? ps[i] : d3[i];
            bool has_vm = false, has_dp = false;
            for (int b = 0; b < nb; b++)
                for (int32_t k : kind[b]) {
                    if (k == AMARU_MAT_VON_MISES) has_vm = true;
                    if (k == AMARU_MAT_DRUCKER_PRAGER) has_dp = true;
                }
            // groups are stored base | VonMises | DruckerPrager; amaru_recovery_field reports them in first-appearance order
            r->group_f0[0] = 0;
            r->group_f0[1] = n;
            if (has_vm) r->codes[n++] = 16;
            r->group_f0[2] = n;
            if (has_dp) { r->codes[n++] = 17; r->codes[n++] = 18; r->codes[n++] = 19; }
            r->group_f0[3] = n;
            r->nfields = n;
            r->group_mask[0] = (1 << AMARU_MAT_LINEAR_ELASTIC) | (1 << AMARU_MAT_VON_MISES) | (1 << AMARU_MAT_DRUCKER_PRAGER);
            r->group_mask[1] = 1 << AMARU_MAT_VON_MISES;
            r->group_mask[2] = 1 << AMARU_MAT_DRUCKER_PRAGER;
            m->recovery_vm_first = true;
            if (has_vm && has_dp) {   // which of the two kinds shows up first in element order
                bool done = false;
                for (int b = 0; b < nb && !done; b++)
                    for (int32_t k : kind[b])
                        if (k == AMARU_MAT_VON_MISES || k == AMARU_MAT_DRUCKER_PRAGER) {
                            m->recovery_vm_first = k == AMARU_MAT_VON_MISES;
                            done = true;
                            break;
                        }
            }
        }
        // corner lists (ascending global element id = the reference's push! order)
        std::vector<int64_t> cptr((size_t)nnodes + 1, 0), aptr((size_t)nnodes + 1, 0);
        for (int b = 0; b < nb; b++) {
            const Batch &B = m->batches[b];
            const int nc = ncorner(B.shape);
            for (int64_t e = 0; e < B.nelem; e++)
                for (int a = 0; a < B.nn; a++) {
                    const int32_t n = conn[b][(size_t)e * B.nn + a];
                    aptr[(size_t)n + 1]++;
                    if (a < nc) cptr[(size_t)n + 1]++;
                }
        }
        for (int64_t i = 0; i < nnodes; i++) { cptr[i + 1] += cptr[i]; aptr[i + 1] += aptr[i]; }
        struct ERef { int32_t b; int64_t e; };
        std::vector<ERef> celem((size_t)cptr[nnodes]), aelem((size_t)aptr[nnodes]);
        {
            std::vector<int64_t> cf(cptr.begin(), cptr.end() - 1), af(aptr.begin(), aptr.end() - 1);
            for (int b = 0; b < nb; b++) {
                const Batch &B = m->batches[b];
                const int nc = ncorner(B.shape);
                for (int64_t e = 0; e < B.nelem; e++)
                    for (int a = 0; a < B.nn; a++) {
                        const int32_t n = conn[b][(size_t)e * B.nn + a];
                        aelem[(size_t)af[n]++] = ERef{b, e};
                        if (a < nc) celem[(size_t)cf[n]++] = ERef{b, e};
                    }
            }
        }
        // active patches (fe-model.jl:529-582)
        std::vector<uint8_t> active((size_t)nnodes, 0), haspatch((size_t)nnodes, 0);
        auto mark = [&](int64_t c) {
            for (int64_t k = cptr[c]; k < cptr[c + 1]; k++) {
                const ERef E = celem[(size_t)k];
                const int nn = m->batches[E.b].nn;
                for (int a = 0; a < nn; a++) haspatch[(size_t)conn[E.b][(size_t)E.e * nn + a]] = 1;
            }
        };
        bool any_bound = false;
        for (int64_t n = 0; n < nnodes; n++) {
            if (at_bound[n]) any_bound = true;
            if (!at_bound[n] && cptr[n + 1] > cptr[n]) {
                active[(size_t)n] = 1;
                mark(n);
            }
        }
        std::vector<int64_t> orphans;
        for (int64_t n = 0; n < nnodes; n++)
            if (!haspatch[(size_t)n] && at_bound[n]) orphans.push_back(n);
        for (int need = 3; need >= 1 && !orphans.empty(); need--) {
            for (int64_t n : orphans)
                if (cptr[n + 1] - cptr[n] >= need) active[(size_t)n] = 1;
            for (int64_t n : orphans)
                if (active[(size_t)n]) mark(n);
            std::vector<int64_t> rest;
            for (int64_t n : orphans)
                if (!haspatch[(size_t)n]) rest.push_back(n);
            orphans.swap(rest);
        }
        if (!any_bound) std::fill(active.begin(), active.end(), 0);   // no faces: nothing is recovered (fe-model.jl:511)
        std::vector<int32_t> pidx((size_t)nnodes, -1), pcorner;
        std::vector<int64_t> pptr(1, 0);
        std::vector<PatchElem> pe;
        for (int64_t n = 0; n < nnodes; n++) {
            if (!active[(size_t)n]) continue;
            pidx[(size_t)n] = (int32_t)pcorner.size();
            pcorner.push_back((int32_t)n);
            for (int64_t k = cptr[n]; k < cptr[n + 1]; k++) {
                const ERef E = celem[(size_t)k];
                const Batch &B = m->batches[E.b];
                pe.push_back(PatchElem{B.ip_off + pos[E.b][(size_t)E.e] * B.nip, B.nip, kind[E.b][(size_t)E.e]});
            }
            pptr.push_back((int64_t)pe.size());
        }
        r->npatch = (int64_t)pcorner.size();
        // node -> patches: every active corner of every element containing the node, with the groups that element feeds
        std::vector<int64_t> nptr((size_t)nnodes + 1, 0);
        std::vector<std::vector<std::pair<int32_t, uint8_t>>> chunks;
        auto node_list = [&](int64_t n, std::vector<std::pair<int32_t, uint8_t>> &out) {
            out.clear();
            for (int64_t k = aptr[n]; k < aptr[n + 1]; k++) {
                const ERef E = aelem[(size_t)k];
                const Batch &B = m->batches[E.b];
                const int nc = ncorner(B.shape);
                const int kd = kind[E.b][(size_t)E.e];
                const uint8_t bits = (uint8_t)(1 | (kd == AMARU_MAT_VON_MISES ? 2 : 0) | (kd == AMARU_MAT_DRUCKER_PRAGER ? 4 : 0));
                for (int a = 0; a < nc; a++) {
                    const int32_t p = pidx[(size_t)conn[E.b][(size_t)E.e * B.nn + a]];
                    if (p >= 0) out.emplace_back(p, bits);
                }
            }
            std::sort(out.begin(), out.end());
            size_t w = 0;
            for (size_t i = 0; i < out.size(); i++) {
                if (w > 0 && out[w - 1].first == out[i].first) out[w - 1].second |= out[i].second;
                else out[w++] = out[i];
            }
            out.resize(w);
        };
        parallel_chunks(nnodes, [&](int64_t lo, int64_t hi) {
            std::vector<std::pair<int32_t, uint8_t>> tmp;
            for (int64_t n = lo; n < hi; n++) {
                node_list(n, tmp);
                nptr[(size_t)n + 1] = (int64_t)tmp.size();
            }
        });
        for (int64_t i = 0; i < nnodes; i++) nptr[i + 1] += nptr[i];
        std::vector<int32_t> npl((size_t)nptr[nnodes]);
        std::vector<uint8_t> nbits((size_t)nptr[nnodes]);
        parallel_chunks(nnodes, [&](int64_t lo, int64_t hi) {
            std::vector<std::pair<int32_t, uint8_t>> tmp;
            for (int64_t n = lo; n < hi; n++) {
                node_list(n, tmp);
                for (size_t i = 0; i < tmp.size(); i++) {
                    npl[(size_t)nptr[n] + i] = tmp[i].first;
                    nbits[(size_t)nptr[n] + i] = tmp[i].second;
                }
            }
        });
        // upload
        std::vector<int32_t> codes(r->codes, r->codes + r->nfields);
        r->d_codes = upload(codes);
        r->d_pcorner = upload(pcorner);
        r->d_pptr = upload(pptr);
        r->d_pe = upload(pe);
        r->d_nptr = upload(nptr);
        r->d_npatch = upload(npl);
        r->d_nbits = upload(nbits);
        const size_t np1 = (size_t)std::max<int64_t>(r->npatch, 1);
        CUDA_CHECK(cudaMalloc(&r->d_ipx, std::max<size_t>((size_t)3 * m->nip_total, 1) * sizeof(double)));
        CUDA_CHECK(cudaMalloc(&r->d_coef, np1 * MAXT * std::max(r->nfields, 1) * sizeof(double)));
        CUDA_CHECK(cudaMalloc(&r->d_pcs, (size_t)NGROUPS * np1 * 4 * sizeof(double)));
        CUDA_CHECK(cudaMalloc(&r->d_pnt, (size_t)NGROUPS * np1 * sizeof(int32_t)));
        CUDA_CHECK(cudaMemset(r->d_pnt, 0, (size_t)NGROUPS * np1 * sizeof(int32_t)));
        CUDA_CHECK(cudaMalloc(&r->d_V, (size_t)std::max(r->nfields, 1) * nnodes * sizeof(double)));
        // integration-point coordinates (element.jl:160-164), device order
        for (Batch &B : m->batches) {
            const int64_t n = B.nelem * B.nip;
            if (n == 0) continue;
            const int g = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)m->nsm * 16));
#define IPC(NN, ND) k_ip_coords<NN, ND><<<g, 256, 0, m->stream>>>(B.nelem, B.nip, B.d_conn, m->d_coords, B.d_N, B.ip_off, m->nip_total, r->d_ipx)
            switch (B.shape) {
            case AMARU_SHAPE_QUAD4: IPC(4, 2); break;
            case AMARU_SHAPE_QUAD8: IPC(8, 2); break;
            case AMARU_SHAPE_HEX8: IPC(8, 3); break;
            case AMARU_SHAPE_HEX20: IPC(20, 3); break;
            case AMARU_SHAPE_TET10: IPC(10, 3); break;
            }
#undef IPC
            m->launches++;
        }
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaStreamSynchronize(m->stream));
        return AMARU_OK;
    });
}

int amaru_recovery_nfields(const amaru_model *m) {
    if (!m || !m->recovery) return -1;
    return static_cast<const Recovery *>(m->recovery)->nfields;
}

// i-th field in the reference's column order -> its storage row in V, code and UTF-8 name
static int field_slot(const amaru_model *m, int i) {
    const Recovery *r = static_cast<const Recovery *>(m->recovery);
    const int nbase = r->group_f0[1], nvm = r->group_f0[2] - r->group_f0[1], ndp = r->group_f0[3] - r->group_f0[2];
    if (i < nbase) return i;
    if (m->recovery_vm_first || nvm == 0 || ndp == 0) return i;   // storage order == reference order
    i -= nbase;                                                    // DruckerPrager fields come first
    return i < ndp ? r->group_f0[2] + i : r->group_f0[1] + (i - ndp);
}

int amaru_recovery_field(const amaru_model *m, int i, int *code, char *name, int namelen) {
    if (!m || !m->recovery) return AMARU_ERR_ARG;
    const Recovery *r = static_cast<const Recovery *>(m->recovery);
    if (i < 0 || i >= r->nfields) return AMARU_ERR_ARG;
    const int c = r->codes[field_slot(m, i)];
    if (code) *code = c;
    if (name && namelen > 0) std::snprintf(name, (size_t)namelen, "%s", FIELD_NAMES[c]);
    return AMARU_OK;
}

int amaru_recover_nodal(amaru_model *m, double *V, char *msg, int msglen) {
    return guarded(msg, msglen, [&]() {
        AMARU_REQUIRE(m && V, AMARU_ERR_ARG, "amaru_recover_nodal: null argument");
        AMARU_REQUIRE(m->recovery, AMARU_ERR_ARG, "amaru_recover_nodal: call amaru_recovery_create first");
        Recovery *r = static_cast<Recovery *>(m->recovery);
        CUDA_CHECK(cudaSetDevice(m->device));
        const int nf = r->nfields;
        if (nf == 0) return AMARU_OK;
        if (m->grp) amaru_group_refresh_output_state(m);   // multi-GPU handle: gather the owners' IP state first
        if (r->npatch > 0) {
            const int gfit = (int)std::max<int64_t>(1, std::min<int64_t>((r->npatch + 3) / 4, (int64_t)m->nsm * 16));
            for (int g = 0; g < NGROUPS; g++) {
                const int f0 = r->group_f0[g], nfg = r->group_f0[g + 1] - f0;
                if (nfg == 0) continue;
                double *pcs = r->d_pcs + (size_t)g * r->npatch * 4;
                int32_t *pnt = r->d_pnt + (size_t)g * r->npatch;
                if (m->nd == 3)
                    k_patch_fit<3><<<gfit, 128, 0, m->stream>>>(r->npatch, r->d_pptr, r->d_pe, r->d_pcorner, m->d_coords, r->d_ipx,
                                                                 m->d_state, m->nip_total, r->group_mask[g], f0, nfg, nf, r->d_codes,
                                                                 r->d_coef, pcs, pnt);
                else
                    k_patch_fit<2><<<gfit, 128, 0, m->stream>>>(r->npatch, r->d_pptr, r->d_pe, r->d_pcorner, m->d_coords, r->d_ipx,
                                                                 m->d_state, m->nip_total, r->group_mask[g], f0, nfg, nf, r->d_codes,
                                                                 r->d_coef, pcs, pnt);
                m->launches++;
            }
        }
        const int gev = (int)std::max<int64_t>(1, std::min<int64_t>((m->nnodes + 3) / 4, (int64_t)m->nsm * 16));
        if (m->nd == 3)
            k_node_eval<3><<<gev, 128, 0, m->stream>>>(m->nnodes, r->npatch, r->d_nptr, r->d_npatch, r->d_nbits, m->d_coords, r->d_coef,
                                                        r->d_pcs, r->d_pnt, nf, r->group_f0[1], r->group_f0[2], r->d_V);
        else
            k_node_eval<2><<<gev, 128, 0, m->stream>>>(m->nnodes, r->npatch, r->d_nptr, r->d_npatch, r->d_nbits, m->d_coords, r->d_coef,
                                                        r->d_pcs, r->d_pnt, nf, r->group_f0[1], r->group_f0[2], r->d_V);
        m->launches++;
        CUDA_CHECK(cudaGetLastError());
        // storage rows -> the reference's column order
        for (int i = 0; i < nf; i++)
            CUDA_CHECK(cudaMemcpyAsync(V + (size_t)i * m->nnodes, r->d_V + (size_t)field_slot(m, i) * m->nnodes,
                                       (size_t)m->nnodes * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
        CUDA_CHECK(cudaStreamSynchronize(m->stream));
        return AMARU_OK;
    });
}

}  // extern "C"
