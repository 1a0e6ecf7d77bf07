// Host-side structured block mesher and dof numbering, straight into the flat arrays amaru_create takes,
// so that 1M-5M element models never exist as per-node / per-cell objects.
//   Block -> nodes / cells   reference src/mesh/structured.jl:182-231 (2D), :384-552 (3D), box corners src/mesh/block.jl:3-30,
//                            node rounding src/node.jl:57-61, ids = creation order src/mesh/mesh.jl:348-356
//   configure_dofs!          reference src/bc.jl:198-233 (unknown dofs first, stable)
// Same creation order as the reference: grid points k (outer), j, i (inner) with the serendipity points skipped; cells
// k, j, i; local node orders of structured.jl:211-222,416-426,499-542.  Two-corner boxes with uniform spacing only.
// Threads: node and cell loops are split over the host cores by k-planes (ids come from closed-form prefix counts).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "amaru_internal.h"

namespace {

struct Grid {
    int shape, nd, f;            // f = 2 for quadratic cells
    int64_t mx, my, mz;          // grid intervals per direction (points = m+1)
    int nn;
};

bool make_grid(int shape, int nx, int ny, int nz, Grid &g) {
    g.shape = shape;
    switch (shape) {
    case AMARU_SHAPE_QUAD4: g.nd = 2; g.f = 1; g.nn = 4; break;
    case AMARU_SHAPE_QUAD8: g.nd = 2; g.f = 2; g.nn = 8; break;
    case AMARU_SHAPE_HEX8: g.nd = 3; g.f = 1; g.nn = 8; break;
    case AMARU_SHAPE_HEX20: g.nd = 3; g.f = 2; g.nn = 20; break;
    case AMARU_SHAPE_TET10: g.nd = 3; g.f = 2; g.nn = 10; break;
    default: return false;
    }
    if (nx < 1 || ny < 1 || (g.nd == 3 && nz < 1)) return false;
    g.mx = (int64_t)g.f * nx;
    g.my = (int64_t)g.f * ny;
    g.mz = g.nd == 3 ? (int64_t)g.f * nz : 0;
    return true;
}

// is grid point (i,j,k) a node?  QUAD8: not both odd (structured.jl:186); HEX20: fewer than two odd (:466-470)
inline bool kept(const Grid &g, int64_t i, int64_t j, int64_t k) {
    if (g.shape == AMARU_SHAPE_QUAD8) return !((i & 1) && (j & 1));
    if (g.shape == AMARU_SHAPE_HEX20) return ((i & 1) + (j & 1) + (k & 1)) < 2;
    return true;
}
// nodes in one grid row (fixed j,k) and in one grid plane (fixed k)
inline int64_t row_count(const Grid &g, int64_t j, int64_t k) {
    const int64_t full = g.mx + 1, even = g.mx / 2 + 1;
    if (g.shape == AMARU_SHAPE_QUAD8) return (j & 1) ? even : full;
    if (g.shape == AMARU_SHAPE_HEX20) {
        const int odd = (int)(j & 1) + (int)(k & 1);
        return odd == 0 ? full : odd == 1 ? even : 0;
    }
    return full;
}
inline int64_t plane_count(const Grid &g, int64_t k) {
    int64_t s = 0;
    // rows with even j: my/2+1 (quadratic) ; odd j: my/2
    if (g.f == 1) return (g.mx + 1) * (g.my + 1);
    const int64_t jeven = g.my / 2 + 1, jodd = g.my / 2;
    s = jeven * row_count(g, 0, k) + jodd * row_count(g, 1, k);
    return s;
}
inline int64_t node_id(const Grid &g, const std::vector<int64_t> &plane_off, int64_t i, int64_t j, int64_t k) {
    int64_t id = plane_off[(size_t)k];
    if (g.f == 1) return id + j * (g.mx + 1) + i;
    const int64_t jeven = (j + 1) / 2, jodd = j / 2;           // rows before j
    id += jeven * row_count(g, 0, k) + jodd * row_count(g, 1, k);
    const int64_t rc = row_count(g, j, k);
    return id + (rc == g.mx + 1 ? i : i / 2);                  // reduced rows keep the even i only
}

inline double axis_point(double x0, double x1, int64_t m, int64_t i) {
    // block shape-function interpolation r = -1 + 2 i/m (structured.jl:392,475), then round to 8 digits (node.jl:58-60)
    const double r = -1.0 + 2.0 * ((1.0 / (double)m) * (double)i);
    const double x = 0.5 * (1.0 - r) * x0 + 0.5 * (1.0 + r) * x1;
    return std::nearbyint(x * 1e8) / 1e8 + 0.0;
}

template <class F>
void parallel_range(int64_t n, F f) {
    int nt = amaru_host_threads();
    if (n < 4) nt = 1;
    if (nt > n) nt = (int)n;
    if (nt <= 1) {
        f(0, n);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([=] { f(n * t / nt, n * (t + 1) / nt); });
    for (auto &x : th) x.join();
}

}  // namespace

extern "C" {

int amaru_mesh_block_sizes(int shape, int nx, int ny, int nz, int64_t *nnodes, int64_t *nelems, int *nn) {
    Grid g;
    if (!make_grid(shape, nx, ny, nz, g)) return AMARU_ERR_ARG;
    int64_t n = 0;
    for (int64_t k = 0; k <= g.mz; k++) n += plane_count(g, k);
    if (nnodes) *nnodes = n;
    if (nelems) *nelems = (int64_t)nx * ny * (g.nd == 3 ? nz : 1) * (shape == AMARU_SHAPE_TET10 ? 6 : 1);
    if (nn) *nn = g.nn;
    return AMARU_OK;
}

int amaru_mesh_block(int shape, const double *box, int nx, int ny, int nz, double *coords, int32_t *conn, char *msg,
                     int msglen) {
    auto fail = [&](const char *s) {
        if (msg && msglen > 0) std::snprintf(msg, (size_t)msglen, "%s", s);
        return AMARU_ERR_ARG;
    };
    if (msg && msglen > 0) msg[0] = 0;
    Grid g;
    if (!box || !coords || !conn) return fail("amaru_mesh_block: null argument");
    if (!make_grid(shape, nx, ny, nz, g)) return fail("block: cannot discretize using this shape / these divisions");
    std::vector<int64_t> plane_off((size_t)g.mz + 2, 0);
    for (int64_t k = 0; k <= g.mz; k++) plane_off[(size_t)k + 1] = plane_off[(size_t)k] + plane_count(g, k);
    if (plane_off[(size_t)g.mz + 1] > 2147483647LL) return fail("amaru_mesh_block: more than 2^31-1 nodes");
    std::vector<double> xs((size_t)g.mx + 1), ys((size_t)g.my + 1), zs((size_t)g.mz + 1);
    for (int64_t i = 0; i <= g.mx; i++) xs[(size_t)i] = axis_point(box[0], box[3], g.mx, i);
    for (int64_t j = 0; j <= g.my; j++) ys[(size_t)j] = axis_point(box[1], box[4], g.my, j);
    for (int64_t k = 0; k <= g.mz; k++) zs[(size_t)k] = g.nd == 3 ? axis_point(box[2], box[5], g.mz, k) : 0.0;
    // nodes, creation order k, j, i
    parallel_range(g.mz + 1, [&](int64_t k0, int64_t k1) {
        for (int64_t k = k0; k < k1; k++) {
            int64_t id = plane_off[(size_t)k];
            for (int64_t j = 0; j <= g.my; j++)
                for (int64_t i = 0; i <= g.mx; i++) {
                    if (!kept(g, i, j, k)) continue;
                    coords[id * 3 + 0] = xs[(size_t)i];
                    coords[id * 3 + 1] = ys[(size_t)j];
                    coords[id * 3 + 2] = zs[(size_t)k];
                    id++;
                }
        }
    });
    auto P = [&](int64_t i, int64_t j, int64_t k) { return (int32_t)node_id(g, plane_off, i, j, k); };
    const int f = g.f;
    if (g.nd == 2) {
        for (int64_t cj = 0; cj < ny; cj++)
            for (int64_t ci = 0; ci < nx; ci++) {
                const int64_t i = ci * f, j = cj * f;
                int32_t *c = conn + (cj * nx + ci) * g.nn;
                if (shape == AMARU_SHAPE_QUAD4) {
                    c[0] = P(i, j, 0); c[1] = P(i + 1, j, 0); c[2] = P(i + 1, j + 1, 0); c[3] = P(i, j + 1, 0);
                } else {   // structured.jl:211-222
                    c[0] = P(i, j, 0); c[1] = P(i + 2, j, 0); c[2] = P(i + 2, j + 2, 0); c[3] = P(i, j + 2, 0);
                    c[4] = P(i + 1, j, 0); c[5] = P(i + 2, j + 1, 0); c[6] = P(i + 1, j + 2, 0); c[7] = P(i, j + 1, 0);
                }
            }
        return AMARU_OK;
    }
    // 27-point stencil p1..p27 of a quadratic cell (structured.jl:499-533), 0-based here
    static const int OFF[27][3] = {{0, 0, 0}, {2, 0, 0}, {2, 2, 0}, {0, 2, 0}, {0, 0, 2}, {2, 0, 2}, {2, 2, 2}, {0, 2, 2},
                                   {1, 0, 0}, {2, 1, 0}, {1, 2, 0}, {0, 1, 0}, {1, 0, 2}, {2, 1, 2}, {1, 2, 2}, {0, 1, 2},
                                   {0, 0, 1}, {2, 0, 1}, {2, 2, 1}, {0, 2, 1},
                                   {0, 1, 1}, {2, 1, 1}, {1, 0, 1}, {1, 2, 1}, {1, 1, 0}, {1, 1, 2}, {1, 1, 1}};
    // six TET10 per cell (structured.jl:537-542), 1-based stencil ids
    static const int TETS[6][10] = {{2, 4, 1, 8, 25, 12, 9, 27, 20, 21}, {2, 1, 5, 8, 9, 17, 23, 27, 21, 16},
                                    {2, 5, 6, 8, 23, 13, 18, 27, 16, 26}, {2, 6, 7, 8, 18, 14, 22, 27, 26, 15},
                                    {2, 3, 4, 8, 10, 11, 25, 27, 24, 20}, {2, 7, 3, 8, 22, 19, 10, 27, 15, 24}};
    parallel_range(nz, [&](int64_t c0, int64_t c1) {
        for (int64_t ck = c0; ck < c1; ck++)
            for (int64_t cj = 0; cj < ny; cj++)
                for (int64_t ci = 0; ci < nx; ci++) {
                    const int64_t i = ci * f, j = cj * f, k = ck * f;
                    const int64_t cell = (ck * ny + cj) * nx + ci;
                    if (shape == AMARU_SHAPE_HEX8) {   // structured.jl:416-426
                        int32_t *c = conn + cell * 8;
                        c[0] = P(i, j, k); c[1] = P(i + 1, j, k); c[2] = P(i + 1, j + 1, k); c[3] = P(i, j + 1, k);
                        c[4] = P(i, j, k + 1); c[5] = P(i + 1, j, k + 1); c[6] = P(i + 1, j + 1, k + 1); c[7] = P(i, j + 1, k + 1);
                    } else if (shape == AMARU_SHAPE_HEX20) {
                        int32_t *c = conn + cell * 20;
                        for (int a = 0; a < 20; a++) c[a] = P(i + OFF[a][0], j + OFF[a][1], k + OFF[a][2]);
                    } else {
                        int32_t p[27];
                        for (int a = 0; a < 27; a++) p[a] = P(i + OFF[a][0], j + OFF[a][1], k + OFF[a][2]);
                        int32_t *c = conn + cell * 60;
                        for (int t = 0; t < 6; t++)
                            for (int a = 0; a < 10; a++) c[t * 10 + a] = p[TETS[t][a] - 1];
                    }
                }
    });
    return AMARU_OK;
}

// configure_dofs! (bc.jl:198-233): dofs in node order, ux uy [uz] per node, stable split into unknown then prescribed
int amaru_configure_dofs(int64_t nnodes, int nd, const uint8_t *prescribed, int32_t *eqid, int64_t *nu) {
    if (!prescribed || !eqid || nnodes < 0 || (nd != 2 && nd != 3)) return AMARU_ERR_ARG;
    const int64_t n = nnodes * nd;
    if (n > 2147483647LL) return AMARU_ERR_ARG;
    int64_t free_count = 0;
    for (int64_t i = 0; i < n; i++) free_count += prescribed[i] ? 0 : 1;
    int64_t a = 0, b = free_count;
    for (int64_t i = 0; i < n; i++) eqid[i] = (int32_t)(prescribed[i] ? b++ : a++);
    if (nu) *nu = free_count;
    return AMARU_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- outer facets
// get_outer_facets (reference src/mesh/mesh.jl:69-85): the facets seen by exactly one cell, in cell order then local
// facet order (facet_idxs of src/shape/solids2d.jl:327,427; solids3d.jl:206,519,706).  The reference hashes every facet into
// a Dict; here the facets are keyed by their sorted corner nodes, sorted in parallel chunks and merged.
namespace {
struct FacetTable {
    int nf, nfn, nc;          // facets per cell, nodes per facet, corner nodes per facet (the first nc of each row)
    const int *idx;           // nf x nfn, 0-based local node ids
};
const int F_QUAD4[] = {0, 1, 1, 2, 2, 3, 3, 0};
const int F_QUAD8[] = {0, 1, 4, 1, 2, 5, 2, 3, 6, 3, 0, 7};
const int F_HEX8[] = {0, 4, 7, 3, 1, 2, 6, 5, 0, 1, 5, 4, 2, 3, 7, 6, 0, 3, 2, 1, 4, 5, 6, 7};
const int F_HEX20[] = {0, 4, 7, 3, 16, 15, 19, 11, 1, 2, 6, 5, 9, 18, 13, 17, 0, 1, 5, 4, 8, 17, 12, 16,
                       2, 3, 7, 6, 10, 19, 14, 18, 0, 3, 2, 1, 11, 10, 9, 8, 4, 5, 6, 7, 12, 13, 14, 15};
const int F_TET10[] = {0, 3, 2, 7, 9, 6, 0, 1, 3, 4, 8, 7, 0, 2, 1, 6, 5, 4, 1, 2, 3, 5, 9, 8};
bool facet_table(int shape, FacetTable &t) {
    switch (shape) {
    case AMARU_SHAPE_QUAD4: t = {4, 2, 2, F_QUAD4}; return true;
    case AMARU_SHAPE_QUAD8: t = {4, 3, 2, F_QUAD8}; return true;
    case AMARU_SHAPE_HEX8: t = {6, 4, 4, F_HEX8}; return true;
    case AMARU_SHAPE_HEX20: t = {6, 8, 4, F_HEX20}; return true;
    case AMARU_SHAPE_TET10: t = {4, 6, 3, F_TET10}; return true;
    }
    return false;
}
struct FKey {
    int32_t k[4];
    int64_t id;   // cell * nf + local facet
    bool operator<(const FKey &o) const {
        for (int i = 0; i < 4; i++)
            if (k[i] != o.k[i]) return k[i] < o.k[i];
        return id < o.id;
    }
    bool same(const FKey &o) const { return k[0] == o.k[0] && k[1] == o.k[1] && k[2] == o.k[2] && k[3] == o.k[3]; }
};
}  // namespace

// facet_nodes [capacity * nfn] and owner [capacity] may be NULL to count only; returns the number of outer facets (< 0: error)
extern "C" int64_t amaru_outer_facets(int shape, int64_t nelem, const int32_t *conn, int32_t *facet_nodes, int64_t *owner,
                                      int64_t capacity, int *nodes_per_facet) {
    FacetTable T;
    Grid g;
    if (!facet_table(shape, T) || !make_grid(shape, 1, 1, 1, g) || !conn || nelem < 0) return AMARU_ERR_ARG;
    if (nodes_per_facet) *nodes_per_facet = T.nfn;
    const int nn = g.nn;
    const int64_t total = nelem * T.nf;
    std::vector<FKey> keys((size_t)total);
    parallel_range(nelem, [&](int64_t e0, int64_t e1) {
        for (int64_t e = e0; e < e1; e++)
            for (int f = 0; f < T.nf; f++) {
                FKey &K = keys[(size_t)(e * T.nf + f)];
                for (int i = 0; i < 4; i++) K.k[i] = i < T.nc ? conn[e * nn + T.idx[f * T.nfn + i]] : -1;
                std::sort(K.k, K.k + 4);
                K.id = e * T.nf + f;
            }
    });
    // parallel sort: sorted chunks, then pairwise merges
    int nt = amaru_host_threads();
    if (total < 65536) nt = 1;
    std::vector<int64_t> cut((size_t)nt + 1);
    for (int t = 0; t <= nt; t++) cut[(size_t)t] = total * t / nt;
    {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back([&, t] { std::sort(keys.begin() + cut[(size_t)t], keys.begin() + cut[(size_t)t + 1]); });
        for (auto &x : th) x.join();
    }
    for (int step = 1; step < nt; step *= 2) {
        std::vector<std::thread> th;
        for (int t = 0; t + step < nt; t += 2 * step) {
            const int64_t a = cut[(size_t)t], b = cut[(size_t)t + step], c = cut[(size_t)std::min(t + 2 * step, nt)];
            th.emplace_back([&, a, b, c] { std::inplace_merge(keys.begin() + a, keys.begin() + b, keys.begin() + c); });
        }
        for (auto &x : th) x.join();
    }
    // facets whose key occurs once, back in (cell, local facet) order
    std::vector<int64_t> once;
    for (int64_t i = 0; i < total;) {
        int64_t j = i + 1;
        while (j < total && keys[(size_t)j].same(keys[(size_t)i])) j++;
        if (j == i + 1) once.push_back(keys[(size_t)i].id);
        i = j;
    }
    std::sort(once.begin(), once.end());
    const int64_t nfac = (int64_t)once.size();
    if (facet_nodes && owner) {
        if (nfac > capacity) return AMARU_ERR_ARG;
        for (int64_t i = 0; i < nfac; i++) {
            const int64_t e = once[(size_t)i] / T.nf;
            const int f = (int)(once[(size_t)i] - e * T.nf);
            owner[i] = e;
            for (int a = 0; a < T.nfn; a++) facet_nodes[i * T.nfn + a] = conn[e * nn + T.idx[f * T.nfn + a]];
        }
    }
    return nfac;
}
