// Patch plan of the matrix-free tangent operator (ebe.cu, k_ebe_patch): host preprocessing, once per amaru_create.
//
// The colour-ordered operator re-reads x and y of every node once per element that touches it (4.9 times per application
// for HEX20, measured 2.3x the algorithmic DRAM bytes, profiles/ncu_ebe_mma_v1_r2.txt).  Here the elements of a batch are
// grouped into spatially compact PATCHES (a 4x4x4 brick of hexahedra, 8x8 quadrilaterals, 4x4x2 cells of tetrahedra): one
// warp keeps x and y of the patch's nodes in shared memory while it runs over the patch's elements, so a node is read and
// written once per patch.  Inside a patch the elements are processed in PHASES of up to 8 mutually node-disjoint elements
// (consecutive elements of one global colour), so the accumulation order at every node is fixed; patches are coloured too
// (patches sharing a node get different colours) and processed in colour-major order, a patch waiting for the
// lower-coloured patches it shares nodes with: the result is bitwise reproducible and needs no atomics on y.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "amaru_internal.h"

namespace {

inline uint32_t spread3(uint32_t v) {   // 3 bits -> every third bit
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4);
}

}  // namespace

void amaru_patch_shape_params(int shape, int &pe, int &maxpn, int brick[3]) {
    brick[0] = brick[1] = brick[2] = 4;
    pe = 64;
    switch (shape) {
    case AMARU_SHAPE_QUAD4: maxpn = 96; brick[0] = brick[1] = 8; brick[2] = 1; break;     // 8x8 cells: 81 nodes
    case AMARU_SHAPE_QUAD8: maxpn = 256; brick[0] = brick[1] = 8; brick[2] = 1; break;    // 225 nodes
    case AMARU_SHAPE_HEX8: maxpn = 160; break;                                            // 4x4x4 cells: 125 nodes
    case AMARU_SHAPE_HEX20: maxpn = 448; break;                                           // 425 nodes
    case AMARU_SHAPE_TET10: maxpn = 448; pe = 192; brick[2] = 2; break;                   // 4x4x2 cells x 6 tetrahedra: 405 nodes
    default: maxpn = 0;
    }
}

void amaru_build_patches(int nn, int nd, int pe, int maxpn, const int brick[3], int64_t nelem, const int32_t *sconn,
                         const std::vector<int64_t> &color_off, int64_t nnodes, int64_t nowned, const double *coords,
                         const uint8_t *fixed, std::vector<uint8_t> &touched, PatchPlan &out) {
    out = PatchPlan();
    out.pe = pe;
    out.maxpn = maxpn;
    out.nn = nn;
    if (nelem == 0) return;
    AMARU_REQUIRE(nnodes <= (int64_t)PN_NODE, AMARU_ERR_UNSUPPORTED, "ebe: more than 2^27 nodes per GPU");
    AMARU_REQUIRE(nelem < (int64_t)1 << 31, AMARU_ERR_UNSUPPORTED, "ebe: more than 2^31 elements per batch");
    // ---- (a) centroids, mean bounding-box extents of an element, brick key
    std::vector<double> cen((size_t)nelem * 3, 0.0);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, ext[3] = {0.0, 0.0, 0.0};
    for (int64_t e = 0; e < nelem; e++) {
        double bl[3] = {1e300, 1e300, 1e300}, bh[3] = {-1e300, -1e300, -1e300};
        for (int a = 0; a < nn; a++) {
            const double *X = coords + (size_t)sconn[e * nn + a] * 3;
            for (int d = 0; d < nd; d++) {
                cen[(size_t)e * 3 + d] += X[d] / nn;
                bl[d] = std::min(bl[d], X[d]);
                bh[d] = std::max(bh[d], X[d]);
            }
        }
        for (int d = 0; d < nd; d++) {
            ext[d] += bh[d] - bl[d];
            lo[d] = std::min(lo[d], bl[d]);
            hi[d] = std::max(hi[d], bh[d]);
        }
    }
    // cell size per axis = mean extent of an element's bounding box (exact for structured blocks, also anisotropic ones);
    // at most 2^17 bricks per axis
    double h[3] = {1.0, 1.0, 1.0};
    for (int d = 0; d < nd; d++) h[d] = std::max(std::max(ext[d] / (double)nelem, 1e-300), (hi[d] - lo[d]) / (brick[d] * 131000.0));
    std::vector<std::pair<uint64_t, int32_t>> key((size_t)nelem);
    for (int64_t e = 0; e < nelem; e++) {
        uint64_t b[3] = {0, 0, 0};
        uint32_t s[3] = {0, 0, 0};
        for (int d = 0; d < nd; d++) {
            const double t = std::max(0.0, (cen[(size_t)e * 3 + d] - lo[d]) / (h[d] * brick[d]));
            b[d] = (uint64_t)t;
            s[d] = (uint32_t)std::min(7.0, (t - (double)b[d]) * 8.0);
        }
        const uint64_t sub = spread3(s[0]) | (spread3(s[1]) << 1) | (spread3(s[2]) << 2);
        key[(size_t)e] = {(((b[2] << 17 | b[1]) << 17 | b[0]) << 9) | sub, (int32_t)e};
    }
    std::sort(key.begin(), key.end());
    auto colour_of = [&](int64_t s) { return (int)(std::upper_bound(color_off.begin(), color_off.end(), s) - color_off.begin()) - 1; };

    // ---- (b) greedy fill in key order: a patch closes at a brick boundary, at pe elements or at maxpn nodes
    std::vector<int32_t> pelem_ptr(1, 0), pelem;        // patch -> elements (sorted-element indices)
    pelem.reserve((size_t)nelem);
    std::vector<int32_t> stamp((size_t)nnodes, -1);
    {
        int cur = 0, cnt = 0, nnd = 0;
        uint64_t curbrick = key[0].first >> 9;
        for (int64_t i = 0; i < nelem; i++) {
            const int32_t e = key[(size_t)i].second;
            const uint64_t bk = key[(size_t)i].first >> 9;
            int fresh = 0;
            for (int a = 0; a < nn; a++) fresh += stamp[(size_t)sconn[(int64_t)e * nn + a]] != cur;
            // (an element listing a node twice over-counts `fresh`: harmless, the patch only closes a little early)
            if (cnt > 0 && (bk != curbrick || cnt == pe || nnd + fresh > maxpn)) {
                pelem_ptr.push_back((int32_t)pelem.size());
                cur++;
                cnt = 0;
                nnd = 0;
            }
            curbrick = bk;
            for (int a = 0; a < nn; a++) {
                int32_t &st = stamp[(size_t)sconn[(int64_t)e * nn + a]];
                if (st != cur) {
                    st = cur;
                    nnd++;
                }
            }
            AMARU_REQUIRE(nnd <= maxpn, AMARU_ERR_UNSUPPORTED, "ebe: one element has more nodes than a patch can hold");
            pelem.push_back(e);
            cnt++;
        }
        pelem_ptr.push_back((int32_t)pelem.size());
    }
    const int np = (int)pelem_ptr.size() - 1;

    // ---- (c) node lists (ascending) and node -> patch incidence
    std::vector<int64_t> pnode_ptr((size_t)np + 1, 0);
    std::vector<int32_t> pnode;                          // global node ids, ascending inside a patch
    pnode.reserve((size_t)np * (size_t)std::min<int64_t>(maxpn, (int64_t)pe * nn) / 2);
    std::fill(stamp.begin(), stamp.end(), -1);
    std::vector<int32_t> tmp;
    for (int p = 0; p < np; p++) {
        tmp.clear();
        for (int32_t k = pelem_ptr[(size_t)p]; k < pelem_ptr[(size_t)p + 1]; k++)
            for (int a = 0; a < nn; a++) {
                const int32_t n = sconn[(int64_t)pelem[(size_t)k] * nn + a];
                if (stamp[(size_t)n] != p) {
                    stamp[(size_t)n] = p;
                    tmp.push_back(n);
                }
            }
        std::sort(tmp.begin(), tmp.end());
        pnode.insert(pnode.end(), tmp.begin(), tmp.end());
        pnode_ptr[(size_t)p + 1] = (int64_t)pnode.size();
    }
    std::vector<int64_t> inc_ptr((size_t)nnodes + 1, 0);
    for (int32_t n : pnode) inc_ptr[(size_t)n + 1]++;
    for (int64_t n = 0; n < nnodes; n++) inc_ptr[(size_t)n + 1] += inc_ptr[(size_t)n];
    std::vector<int32_t> inc((size_t)pnode.size());
    {
        std::vector<int64_t> fill(inc_ptr.begin(), inc_ptr.end() - 1);
        for (int p = 0; p < np; p++)
            for (int64_t k = pnode_ptr[(size_t)p]; k < pnode_ptr[(size_t)p + 1]; k++) inc[(size_t)fill[(size_t)pnode[(size_t)k]]++] = p;
    }

    // ---- (d) greedy patch colouring in creation (spatial) order
    std::vector<int32_t> pcol((size_t)np, -1);
    int npc = 0;
    for (int p = 0; p < np; p++) {
        uint64_t used = 0;
        for (int64_t k = pnode_ptr[(size_t)p]; k < pnode_ptr[(size_t)p + 1]; k++) {
            const int32_t n = pnode[(size_t)k];
            for (int64_t q = inc_ptr[(size_t)n]; q < inc_ptr[(size_t)n + 1]; q++) {
                const int32_t c = pcol[(size_t)inc[(size_t)q]];
                if (c >= 0) used |= 1ull << c;
            }
        }
        AMARU_REQUIRE(~used != 0ull, AMARU_ERR_UNSUPPORTED, "ebe: patch colouring needs more than 64 colours");
        pcol[(size_t)p] = __builtin_ctzll(~used);
        npc = std::max(npc, pcol[(size_t)p] + 1);
    }
    std::vector<int32_t> order((size_t)np), pos((size_t)np);
    for (int p = 0; p < np; p++) order[(size_t)p] = p;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return pcol[(size_t)a] < pcol[(size_t)b]; });
    for (int i = 0; i < np; i++) pos[(size_t)order[(size_t)i]] = i;

    // ---- (e) final arrays in colour-major patch order
    const int ks2 = (nn + 3) / 4, nw = amaru_patch_id_words(nn);
    out.npatch = np;
    out.ncolors = npc;
    out.desc.assign((size_t)np * 8, 0);
    out.pnodes.reserve(pnode.size());
    std::vector<int32_t> lid((size_t)nnodes, 0), seen((size_t)np, -1);
    std::vector<std::pair<int32_t, int32_t>> ce;         // (colour, element) of one patch
    int64_t ngroups = 0;
    for (int i = 0; i < np; i++) {
        const int p = order[(size_t)i];
        int32_t *d = &out.desc[(size_t)i * 8];
        d[0] = (int32_t)ngroups;
        d[2] = (int32_t)out.pnodes.size();
        d[3] = (int32_t)(pnode_ptr[(size_t)p + 1] - pnode_ptr[(size_t)p]);
        d[4] = (int32_t)out.deps.size();
        AMARU_REQUIRE(out.pnodes.size() + (size_t)d[3] < ((size_t)1 << 31), AMARU_ERR_UNSUPPORTED, "ebe: patch node lists exceed 2^31 entries");
        for (int64_t k = pnode_ptr[(size_t)p]; k < pnode_ptr[(size_t)p + 1]; k++) {
            const int32_t n = pnode[(size_t)k];
            lid[(size_t)n] = (int32_t)(k - pnode_ptr[(size_t)p]);
            uint32_t ent = (uint32_t)n;
            for (int c = 0; c < nd; c++) ent |= (uint32_t)(fixed[(size_t)n * nd + c] ? 1u : 0u) << (28 + c);
            if (n >= nowned) ent |= PN_GHOST;
            if (!touched[(size_t)n]) {
                touched[(size_t)n] = 1;
                ent |= PN_FIRST;
            }
            out.pnodes.push_back(ent);
            for (int64_t q = inc_ptr[(size_t)n]; q < inc_ptr[(size_t)n + 1]; q++) {   // lower-coloured patches sharing the node
                const int32_t o = inc[(size_t)q];
                if (o != p && pcol[(size_t)o] < pcol[(size_t)p] && seen[(size_t)o] != p) {
                    seen[(size_t)o] = p;
                    out.deps.push_back(pos[(size_t)o]);
                }
            }
        }
        d[5] = (int32_t)out.deps.size() - d[4];
        ce.clear();
        for (int32_t k = pelem_ptr[(size_t)p]; k < pelem_ptr[(size_t)p + 1]; k++) ce.push_back({colour_of(pelem[(size_t)k]), pelem[(size_t)k]});
        std::sort(ce.begin(), ce.end());
        size_t k = 0;
        int ng = 0;
        while (k < ce.size()) {                          // phases: up to 8 elements of one colour
            size_t k1 = k;
            while (k1 < ce.size() && ce[k1].first == ce[k].first && k1 - k < 8) k1++;
            const size_t g0 = out.lane_ids.size();
            out.lane_ids.resize(g0 + (size_t)nw * 128, (uint16_t)maxpn);   // padding / empty slots: the dummy row behind the brick
            for (size_t s = 0; s < 8; s++) {
                const int32_t e = k + s < k1 ? ce[k + s].second : -1;
                out.slot_elem.push_back(e);
                if (e < 0) continue;
                for (int a = 0; a < nn; a++) {           // node a of slot s: fragment positions of both contractions
                    const uint16_t l = (uint16_t)lid[(size_t)sconn[(int64_t)e * nn + a]];
                    const int t1 = a / 4, lane1 = (int)s * 4 + a % 4;
                    const int t2 = ks2 + 2 * (a / 8) + (a % 2), lane2 = (int)s * 4 + (a % 8) / 2;
                    out.lane_ids[g0 + ((size_t)(t1 / 4) * 32 + lane1) * 4 + t1 % 4] = l;
                    out.lane_ids[g0 + ((size_t)(t2 / 4) * 32 + lane2) * 4 + t2 % 4] = l;
                }
            }
            ng++;
            k = k1;
        }
        d[1] = ng;
        ngroups += ng;
    }
    out.nslots = ngroups * 8;
    out.fill = (double)nelem / (double)out.nslots;
    out.maxgroups = 0;
    for (int i = 0; i < np; i++) out.maxgroups = std::max(out.maxgroups, out.desc[(size_t)i * 8 + 1]);
}

// Host-only check of a plan (tests, -m "not gpu"): every element in exactly one slot, phases node-disjoint, local ids
// consistent, dependencies = every lower-coloured patch sharing a node, colour-major order.  Returns 0 or the failed rule.
extern "C" int amaru_patch_plan_check(int shape, int64_t nnodes, int64_t nowned, const double *coords, int64_t nelem,
                                      const int32_t *conn, double *stats /*[6]: patches, colours, fill, slots, max groups, deps*/) {
    try {
        ShapeInfo si;
        if (!amaru_shape_info(shape, si)) return -1;
        const int nn = si.nn, nd = si.nd;
        // colour + colour-sort like amaru_create
        std::vector<int64_t> adj_ptr, adj;
        const int32_t *cp = conn;
        amaru_build_adjacency(nnodes, 1, &nn, &nelem, &cp, adj_ptr, adj);
        std::vector<int32_t> color;
        const int nc = amaru_color_elements(nnodes, 1, &nn, &nelem, &cp, adj_ptr, adj, color);
        if (nc <= 0) return -2;
        std::vector<int64_t> cnt((size_t)nc + 1, 0);
        for (int64_t e = 0; e < nelem; e++) cnt[(size_t)color[(size_t)e] + 1]++;
        for (int c = 0; c < nc; c++) cnt[(size_t)c + 1] += cnt[(size_t)c];
        std::vector<int64_t> fill(cnt.begin(), cnt.end() - 1);
        std::vector<int32_t> sconn((size_t)nelem * nn);
        for (int64_t e = 0; e < nelem; e++) std::memcpy(&sconn[(size_t)(fill[(size_t)color[(size_t)e]]++) * nn], conn + e * nn, sizeof(int32_t) * nn);
        std::vector<uint8_t> fixed((size_t)nnodes * nd, 0), touched((size_t)nnodes, 0);
        PatchPlan P;
        int pe, maxpn, brick[3];
        amaru_patch_shape_params(shape, pe, maxpn, brick);
        amaru_build_patches(nn, nd, pe, maxpn, brick, nelem, sconn.data(), cnt, nnodes, nowned, coords, fixed.data(), touched, P);
        if (stats) {
            stats[0] = P.npatch; stats[1] = P.ncolors; stats[2] = P.fill; stats[3] = (double)P.nslots; stats[4] = P.maxgroups;
            stats[5] = (double)P.deps.size();
        }
        // rule 1: every element exactly once
        std::vector<uint8_t> cov((size_t)nelem, 0);
        for (int32_t e : P.slot_elem)
            if (e >= 0) {
                if (e >= nelem || cov[(size_t)e]) return 1;
                cov[(size_t)e] = 1;
            }
        for (uint8_t c : cov) if (!c) return 1;
        std::vector<int32_t> owner_patch((size_t)nnodes, -1);
        std::vector<uint8_t> first_seen((size_t)nnodes, 0);
        for (int p = 0; p < P.npatch; p++) {
            const int32_t *d = &P.desc[(size_t)p * 8];
            if (d[3] > maxpn) return 2;
            // rule 3: local ids map to the element's nodes; phases are node-disjoint
            for (int g = 0; g < d[1]; g++) {
                std::vector<int32_t> used;
                for (int s = 0; s < 8; s++) {
                    const int64_t slot = ((int64_t)d[0] + g) * 8 + s;
                    const int32_t e = P.slot_elem[(size_t)slot];
                    if (e < 0) continue;
                    const int ks2 = (nn + 3) / 4, nw = amaru_patch_id_words(nn);
                    const size_t g0 = ((size_t)d[0] + g) * nw * 128;
                    for (int a = 0; a < nn; a++) {
                        const int t1 = a / 4, lane1 = s * 4 + a % 4, t2 = ks2 + 2 * (a / 8) + (a % 2), lane2 = s * 4 + (a % 8) / 2;
                        const int l = P.lane_ids[g0 + ((size_t)(t1 / 4) * 32 + lane1) * 4 + t1 % 4];
                        if (l != P.lane_ids[g0 + ((size_t)(t2 / 4) * 32 + lane2) * 4 + t2 % 4]) return 3;
                        if (l >= d[3]) return 3;
                        const uint32_t ent = P.pnodes[(size_t)d[2] + l];
                        if ((int32_t)(ent & PN_NODE) != sconn[(size_t)e * nn + a]) return 3;
                        used.push_back(l);
                    }
                }
                std::sort(used.begin(), used.end());
                if (std::adjacent_find(used.begin(), used.end()) != used.end()) return 4;
            }
            // rule 5: FIRST exactly at the first patch (in order) that lists the node; ghost flag
            for (int k = 0; k < d[3]; k++) {
                const uint32_t ent = P.pnodes[(size_t)d[2] + k];
                const int32_t n = (int32_t)(ent & PN_NODE);
                if (((ent & PN_FIRST) != 0) != (first_seen[(size_t)n] == 0)) return 5;
                first_seen[(size_t)n] = 1;
                if (((ent & PN_GHOST) != 0) != (n >= nowned)) return 5;
            }
        }
        // rule 6: two patches sharing a node: the later one lists the earlier one as a dependency (colour-major order makes
        // "earlier" = lower colour); patches of one colour share no node
        std::vector<std::vector<int32_t>> node_p((size_t)nnodes);
        for (int p = 0; p < P.npatch; p++) {
            const int32_t *d = &P.desc[(size_t)p * 8];
            for (int k = 0; k < d[3]; k++) node_p[(size_t)(P.pnodes[(size_t)d[2] + k] & PN_NODE)].push_back(p);
        }
        for (int p = 0; p < P.npatch; p++) {
            const int32_t *d = &P.desc[(size_t)p * 8];
            std::vector<int32_t> dep(P.deps.begin() + d[4], P.deps.begin() + d[4] + d[5]);
            std::sort(dep.begin(), dep.end());
            for (int32_t q : dep)
                if (q >= p) return 6;
            for (int k = 0; k < d[3]; k++)
                for (int32_t q : node_p[(size_t)(P.pnodes[(size_t)d[2] + k] & PN_NODE)])
                    if (q < p && !std::binary_search(dep.begin(), dep.end(), q)) return 6;
        }
        return 0;
    } catch (const AmaruError &e) {
        return -e.code - 100;
    }
}
