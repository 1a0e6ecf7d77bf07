// Host-side preprocessing of libamaru_b200.so: shape tables, node->element adjacency, element colouring and
// the symbolic block-CSR pattern.  Runs once per amaru_create; multi-threaded with std::thread.
//
// Replaces, for the GPU path, what the reference redoes on every mount_K call: the COO triplet lists and the
// sort/merge inside sparse(R,C,V) (reference src/mech/mech-solver.jl:81-102).  The pattern built here is the
// SYMBOLIC pattern (connectivity x dofs): for every node the union of the nodes of all elements touching it.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

#include "amaru_internal.h"

int amaru_host_threads() {
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 1;
    if (n > 64) n = 64;
    return (int)n;
}

template <class F>
static void parallel_chunks(int64_t n, F f) {
    int nt = amaru_host_threads();
    if (n < 4096) nt = 1;
    if (nt == 1) {
        f(0, n, 0);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) {
        int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        th.emplace_back([=] { f(lo, hi, t); });
    }
    for (auto &x : th) x.join();
}

// ---------------------------------------------------------------------------------------------- shape tables
// Natural coordinates and default quadrature of each shape are data of the reference
// (src/shape/solids2d.jl:287-291,359-367; src/shape/solids3d.jl:108-118,459-467,557-579;
//  src/shape/quadrature.jl:62-66,110-114,167-175).  N and dN/dR are evaluated from the node natural coordinates
// (tensor-product / serendipity / barycentric forms) rather than from per-node expanded polynomials.
static const double NAT_QUAD4[] = {-1, -1, 1, -1, 1, 1, -1, 1};
static const double NAT_QUAD8[] = {-1, -1, 1, -1, 1, 1, -1, 1, 0, -1, 1, 0, 0, 1, -1, 0};
static const double NAT_HEX8[] = {-1, -1, -1, 1, -1, -1, 1, 1, -1, -1, 1, -1, -1, -1, 1, 1, -1, 1, 1, 1, 1, -1, 1, 1};
static const double NAT_HEX20[] = {-1, -1, -1, 1, -1, -1, 1, 1, -1, -1, 1, -1, -1, -1, 1, 1, -1, 1, 1, 1, 1, -1, 1, 1,
                                   0, -1, -1, 1, 0, -1, 0, 1, -1, -1, 0, -1, 0, -1, 1, 1, 0, 1, 0, 1, 1, -1, 0, 1,
                                   -1, -1, 0, 1, -1, 0, 1, 1, 0, -1, 1, 0};
static const double NAT_LIN2[] = {-1, 1};
static const double NAT_LIN3[] = {-1, 1, 0};
static const double NAT_TRI6[] = {0, 0, 1, 0, 0, 1, .5, 0, .5, .5, 0, .5};
static const double NAT_TET10[] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, .5, 0, 0, .5, .5, 0, 0, .5, 0, 0, 0, .5, .5, 0, .5, 0, .5, .5};

static void eval_box(int nn, int nd, const double *nat, const double *R, double *N, double *D) {
    // linear (no zero coordinate anywhere) or serendipity (mid-side nodes have one zero coordinate)
    for (int i = 0; i < nn; i++) {
        const double *xi = nat + i * nd;
        int m = -1;
        for (int d = 0; d < nd; d++)
            if (xi[d] == 0.0) m = d;
        bool quadratic = false;
        for (int j = 0; j < nn * nd; j++)
            if (nat[j] == 0.0) quadratic = true;
        double f[3];
        for (int d = 0; d < nd; d++) f[d] = 1.0 + xi[d] * R[d];
        if (m < 0) {
            double P = 1.0;
            for (int d = 0; d < nd; d++) P *= 0.5 * f[d];
            double S = 1.0, dS[3] = {0, 0, 0};
            if (quadratic) {
                S = -(double)(nd - 1);
                for (int d = 0; d < nd; d++) {
                    S += xi[d] * R[d];
                    dS[d] = xi[d];
                }
            }
            N[i] = P * S;
            for (int d = 0; d < nd; d++) {
                double Pd = 0.5 * xi[d];
                for (int e = 0; e < nd; e++)
                    if (e != d) Pd *= 0.5 * f[e];
                D[i * nd + d] = Pd * S + P * dS[d];
            }
        } else {
            const double q = 1.0 - R[m] * R[m];
            double P = 1.0;
            for (int d = 0; d < nd; d++)
                if (d != m) P *= 0.5 * f[d];
            N[i] = q * P;
            for (int d = 0; d < nd; d++) {
                if (d == m) {
                    D[i * nd + d] = -2.0 * R[m] * P;
                } else {
                    double Pd = 0.5 * xi[d];
                    for (int e = 0; e < nd; e++)
                        if (e != d && e != m) Pd *= 0.5 * f[e];
                    D[i * nd + d] = q * Pd;
                }
            }
        }
    }
}

static void eval_simplex2(int nn, int nd, const double *nat, const double *R, double *N, double *D) {
    double L[4], dL[4][3];
    L[0] = 1.0;
    for (int d = 0; d < nd; d++) {
        L[0] -= R[d];
        L[d + 1] = R[d];
    }
    for (int a = 0; a <= nd; a++)
        for (int d = 0; d < nd; d++) dL[a][d] = (a == 0) ? -1.0 : (a == d + 1 ? 1.0 : 0.0);
    for (int i = 0; i < nn; i++) {
        double Li[4];
        Li[0] = 1.0;
        for (int d = 0; d < nd; d++) {
            Li[0] -= nat[i * nd + d];
            Li[d + 1] = nat[i * nd + d];
        }
        int c[2], k = 0;
        for (int a = 0; a <= nd; a++)
            if (Li[a] > 0.25 && k < 2) c[k++] = a;
        if (k == 1) {
            N[i] = L[c[0]] * (2.0 * L[c[0]] - 1.0);
            for (int d = 0; d < nd; d++) D[i * nd + d] = (4.0 * L[c[0]] - 1.0) * dL[c[0]][d];
        } else {
            N[i] = 4.0 * L[c[0]] * L[c[1]];
            for (int d = 0; d < nd; d++) D[i * nd + d] = 4.0 * (L[c[0]] * dL[c[1]][d] + L[c[1]] * dL[c[0]][d]);
        }
    }
}

bool amaru_shape_info(int id, ShapeInfo &s) {
    const double g = 0.577350269189626;  // quadrature.jl:62-66,167-175
    const double *nat = nullptr;
    s.id = id;
    switch (id) {
    case AMARU_SHAPE_QUAD4: s.nn = 4; s.nd = 2; nat = NAT_QUAD4; break;
    case AMARU_SHAPE_QUAD8: s.nn = 8; s.nd = 2; nat = NAT_QUAD8; break;
    case AMARU_SHAPE_HEX8: s.nn = 8; s.nd = 3; nat = NAT_HEX8; break;
    case AMARU_SHAPE_HEX20: s.nn = 20; s.nd = 3; nat = NAT_HEX20; break;
    case AMARU_SHAPE_TET10: s.nn = 10; s.nd = 3; nat = NAT_TET10; break;
    // facet shapes (src/shape/lines.jl:10-108, src/shape/solids2d.jl:99-151): only used by the load integration
    case AMARU_SHAPE_LIN2: s.nn = 2; s.nd = 1; nat = NAT_LIN2; break;
    case AMARU_SHAPE_LIN3: s.nn = 3; s.nd = 1; nat = NAT_LIN3; break;
    case AMARU_SHAPE_TRI6: s.nn = 6; s.nd = 2; nat = NAT_TRI6; break;
    default: return false;
    }
    s.nat.assign(nat, nat + s.nn * s.nd);
    s.ips.clear();
    if (id == AMARU_SHAPE_TET10) {
        const double a = 0.5854101966249685, b = 0.1381966011250105, w = 0.04166666666666667;  // quadrature.jl:110-114
        const double t[16] = {a, b, b, w, b, a, b, w, b, b, a, w, b, b, b, w};
        s.ips.assign(t, t + 16);
    } else if (id == AMARU_SHAPE_TRI6) {
        const double t[12] = {1.0 / 6, 1.0 / 6, 0, 1.0 / 6, 2.0 / 3, 1.0 / 6, 0, 1.0 / 6, 1.0 / 6, 2.0 / 3, 0, 1.0 / 6};  // quadrature.jl:37-40
        s.ips.assign(t, t + 12);
    } else if (s.nd == 1) {
        const double gl = 0.577350269189625764509149;  // quadrature.jl:16-18 (LIN_IP2)
        const double t[8] = {-gl, 0, 0, 1.0, gl, 0, 0, 1.0};
        s.ips.assign(t, t + 8);
    } else if (s.nd == 2) {
        for (int j = -1; j <= 1; j += 2)
            for (int i = -1; i <= 1; i += 2) {
                const double p[4] = {i * g, j * g, 0.0, 1.0};
                s.ips.insert(s.ips.end(), p, p + 4);
            }
    } else {
        for (int k = -1; k <= 1; k += 2)
            for (int j = -1; j <= 1; j += 2)
                for (int i = -1; i <= 1; i += 2) {
                    const double p[4] = {i * g, j * g, k * g, 1.0};
                    s.ips.insert(s.ips.end(), p, p + 4);
                }
    }
    s.nip = (int)s.ips.size() / 4;
    s.N.resize((size_t)s.nip * s.nn);
    s.dNdR.resize((size_t)s.nip * s.nn * s.nd);
    for (int q = 0; q < s.nip; q++) {
        if (id == AMARU_SHAPE_TET10 || id == AMARU_SHAPE_TRI6)
            eval_simplex2(s.nn, s.nd, nat, &s.ips[4 * q], &s.N[(size_t)q * s.nn], &s.dNdR[(size_t)q * s.nn * s.nd]);
        else
            eval_box(s.nn, s.nd, nat, &s.ips[4 * q], &s.N[(size_t)q * s.nn], &s.dNdR[(size_t)q * s.nn * s.nd]);
    }
    return true;
}

// ---------------------------------------------------------------------------------------------- adjacency
void amaru_build_adjacency(int64_t nnodes, int nbatches, const int *nn, const int64_t *nelem,
                           const int32_t *const *conn, std::vector<int64_t> &adj_ptr, std::vector<int64_t> &adj) {
    adj_ptr.assign((size_t)nnodes + 1, 0);
    for (int b = 0; b < nbatches; b++) {
        const int64_t n = nelem[b] * nn[b];
        for (int64_t i = 0; i < n; i++) adj_ptr[(size_t)conn[b][i] + 1]++;
    }
    for (int64_t i = 0; i < nnodes; i++) adj_ptr[i + 1] += adj_ptr[i];
    adj.resize((size_t)adj_ptr[nnodes]);
    std::vector<int64_t> fill(adj_ptr.begin(), adj_ptr.end() - 1);
    int64_t gid = 0;
    for (int b = 0; b < nbatches; b++)
        for (int64_t e = 0; e < nelem[b]; e++, gid++)
            for (int a = 0; a < nn[b]; a++) adj[(size_t)fill[conn[b][e * nn[b] + a]]++] = gid;  // ascending element ids
}

// ---------------------------------------------------------------------------------------------- colouring
int amaru_color_elements(int64_t nnodes, int nbatches, const int *nn, const int64_t *nelem,
                         const int32_t *const *conn, const std::vector<int64_t> &adj_ptr,
                         const std::vector<int64_t> &adj, std::vector<int32_t> &color) {
    (void)nnodes;
    int64_t total = 0;
    for (int b = 0; b < nbatches; b++) total += nelem[b];
    color.assign((size_t)total, -1);
    constexpr int W = 8;  // up to 512 colours
    int ncolors = 0;
    int64_t gid = 0;
    for (int b = 0; b < nbatches; b++)
        for (int64_t e = 0; e < nelem[b]; e++, gid++) {
            uint64_t used[W] = {0};
            for (int a = 0; a < nn[b]; a++) {
                const int64_t n = conn[b][e * nn[b] + a];
                for (int64_t k = adj_ptr[n]; k < adj_ptr[n + 1]; k++) {
                    const int32_t c = color[(size_t)adj[k]];
                    if (c >= 0) used[c >> 6] |= (1ull << (c & 63));
                }
            }
            int c = 0;
            for (int w = 0; w < W; w++) {
                if (~used[w]) {
                    c = w * 64 + __builtin_ctzll(~used[w]);
                    break;
                }
                c = (w + 1) * 64;
            }
            if (c >= W * 64) return -1;
            color[(size_t)gid] = c;
            if (c + 1 > ncolors) ncolors = c + 1;
        }
    return ncolors;
}

// ---------------------------------------------------------------------------------------------- pattern
void amaru_build_pattern(int64_t nrows, int nbatches, const int *nn, const int64_t *nelem,
                         const int32_t *const *conn, const std::vector<int64_t> &adj_ptr,
                         const std::vector<int64_t> &adj, HostPattern &pat) {
    // batch lookup for a global element id
    std::vector<int64_t> boff(nbatches + 1, 0);
    for (int b = 0; b < nbatches; b++) boff[b + 1] = boff[b] + nelem[b];
    auto row_nodes = [&](int64_t node, std::vector<int32_t> &tmp) {
        tmp.clear();
        for (int64_t k = adj_ptr[node]; k < adj_ptr[node + 1]; k++) {
            const int64_t g = adj[k];
            int b = 0;
            while (g >= boff[b + 1]) b++;
            const int32_t *c = conn[b] + (g - boff[b]) * nn[b];
            tmp.insert(tmp.end(), c, c + nn[b]);
        }
        std::sort(tmp.begin(), tmp.end());
        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    };
    std::vector<int32_t> cnt((size_t)nrows, 0);
    parallel_chunks(nrows, [&](int64_t lo, int64_t hi, int) {
        std::vector<int32_t> tmp;
        for (int64_t i = lo; i < hi; i++) {
            row_nodes(i, tmp);
            cnt[(size_t)i] = (int32_t)tmp.size();
        }
    });
    pat.rowptr.assign((size_t)nrows + 1, 0);
    int64_t tot = 0;
    for (int64_t i = 0; i < nrows; i++) {
        tot += cnt[(size_t)i];
        if (tot > 2147483647LL) throw AmaruError{AMARU_ERR_ARG, "pattern has more than 2^31-1 blocks on one device"};
        pat.rowptr[(size_t)i + 1] = (int32_t)tot;
    }
    pat.col.resize((size_t)tot);
    pat.diag.assign((size_t)nrows, -1);
    parallel_chunks(nrows, [&](int64_t lo, int64_t hi, int) {
        std::vector<int32_t> tmp;
        for (int64_t i = lo; i < hi; i++) {
            row_nodes(i, tmp);
            int32_t *dst = pat.col.data() + pat.rowptr[(size_t)i];
            std::memcpy(dst, tmp.data(), tmp.size() * sizeof(int32_t));
            auto it = std::lower_bound(tmp.begin(), tmp.end(), (int32_t)i);
            if (it != tmp.end() && *it == (int32_t)i) pat.diag[(size_t)i] = pat.rowptr[(size_t)i] + (int32_t)(it - tmp.begin());
        }
    });
}

// ---------------------------------------------------------------------------------------------- CPU-only probe
// Runs the host preprocessing alone (no CUDA calls) so that the host logic can be tested without a GPU.
// rowptr_out [nnodes+1], col_out [capacity] (may be NULL to only count), color_out [nelem_total] (may be NULL).
extern "C" int amaru_host_prep_probe(int64_t nnodes, int nbatches, const int32_t *batch_shape, const int64_t *batch_nelem,
                                     const int32_t *conn, int32_t *rowptr_out, int32_t *col_out, int64_t col_capacity,
                                     int32_t *color_out, int64_t *nblk_out, int *ncolors_out) {
    try {
        std::vector<ShapeInfo> info(nbatches);
        std::vector<int> nn(nbatches);
        std::vector<const int32_t *> connp(nbatches);
        int64_t coff = 0;
        for (int b = 0; b < nbatches; b++) {
            if (!amaru_shape_info(batch_shape[b], info[b])) return AMARU_ERR_UNSUPPORTED;
            nn[b] = info[b].nn;
            connp[b] = conn + coff;
            coff += batch_nelem[b] * nn[b];
        }
        std::vector<int64_t> adj_ptr, adj;
        amaru_build_adjacency(nnodes, nbatches, nn.data(), batch_nelem, connp.data(), adj_ptr, adj);
        std::vector<int32_t> color;
        const int nc = amaru_color_elements(nnodes, nbatches, nn.data(), batch_nelem, connp.data(), adj_ptr, adj, color);
        HostPattern pat;
        amaru_build_pattern(nnodes, nbatches, nn.data(), batch_nelem, connp.data(), adj_ptr, adj, pat);
        if (ncolors_out) *ncolors_out = nc;
        if (nblk_out) *nblk_out = (int64_t)pat.col.size();
        if (rowptr_out) std::memcpy(rowptr_out, pat.rowptr.data(), pat.rowptr.size() * sizeof(int32_t));
        if (col_out && (int64_t)pat.col.size() <= col_capacity) std::memcpy(col_out, pat.col.data(), pat.col.size() * sizeof(int32_t));
        if (color_out) std::memcpy(color_out, color.data(), color.size() * sizeof(int32_t));
        return AMARU_OK;
    } catch (const AmaruError &e) {
        return e.code;
    }
}
