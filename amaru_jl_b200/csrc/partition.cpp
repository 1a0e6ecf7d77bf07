// Host-side domain decomposition for multi-GPU handles: element partition + node ownership + rank-local
// views with duplicated halo elements.  The reference has no distributed path (its src/mesh/partition.jl:4-15 is a spatial
// bin index for point location); this is what `amaru_create(..., ngpus > 1)` runs before it hands each GPU its row block.
//
//   * elements -> parts: recursive coordinate bisection of the centroids (deterministic, compact boxes on structured
//     meshes) or METIS k-way on the element dual graph (METIS_PartMeshDual from the toolkit's libmetis_static.a; elements
//     are adjacent when they share a facet's worth of nodes);
//   * a node is owned by the lowest part touching it;
//   * rank p's local view: owned nodes first (ascending global id), then ghosts grouped by owner (ascending global id in a
//     group); local elements = every element touching an owned node (own + halo elements), batch structure preserved, so
//     the rows and internal forces of the owned nodes are complete without communication;
//   * halo lists: p sends to q exactly the nodes q sees as ghosts owned by p, ascending global id — the order q stores
//     them in — so pushed data lands in place.
// Same contract as amaru_jl_b200/partition.py (the numpy version used by the one-process-per-GPU launcher and its tests).
#include <algorithm>
#include <cstring>
#include <numeric>
#include <thread>

#include "partition.h"

extern "C" {
// libmetis_static.a of the CUDA toolkit: idx_t = int64_t, real_t = float (probed: METIS_PartGraphKway on a 4x4 grid)
int METIS_SetDefaultOptions(int64_t *options);
int METIS_PartMeshDual(int64_t *ne, int64_t *nn, int64_t *eptr, int64_t *eind, int64_t *vwgt, int64_t *vsize, int64_t *ncommon,
                       int64_t *nparts, float *tpwgts, int64_t *options, int64_t *objval, int64_t *epart, int64_t *npart);
}

namespace {

void rcb(const std::vector<double> &cent, std::vector<int64_t> &idx, int64_t lo, int64_t hi, int p0, int k, int32_t *part) {
    if (k == 1) {
        for (int64_t i = lo; i < hi; i++) part[idx[(size_t)i]] = p0;
        return;
    }
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = lo; i < hi; i++)
        for (int d = 0; d < 3; d++) {
            const double v = cent[(size_t)idx[(size_t)i] * 3 + d];
            mn[d] = std::min(mn[d], v);
            mx[d] = std::max(mx[d], v);
        }
    int axis = 0;
    for (int d = 1; d < 3; d++)
        if (mx[d] - mn[d] > mx[axis] - mn[axis]) axis = d;
    const int kl = k / 2;
    const int64_t nl = ((hi - lo) * kl) / k;
    // stable order along the axis (ties keep the order of the previous level): the same split as partition.py's argsort
    std::stable_sort(idx.begin() + lo, idx.begin() + hi, [&](int64_t a, int64_t b) { return cent[(size_t)a * 3 + axis] < cent[(size_t)b * 3 + axis]; });
    rcb(cent, idx, lo, lo + nl, p0, kl, part);
    rcb(cent, idx, lo + nl, hi, p0 + kl, k - kl, part);
}

int facet_nodes(int shape) {   // nodes two neighbouring cells share across a facet (METIS ncommon)
    switch (shape) {
    case AMARU_SHAPE_QUAD4: return 2;
    case AMARU_SHAPE_QUAD8: return 3;
    case AMARU_SHAPE_HEX8: return 4;
    case AMARU_SHAPE_HEX20: return 8;
    case AMARU_SHAPE_TET10: return 6;
    }
    return 1;
}

}  // namespace

void amaru_partition_elements(int method, int nparts, int64_t nnodes, const double *coords, int nbatches, const int *nn,
                              const int32_t *batch_shape, const int64_t *nelem, const int32_t *const *conn,
                              std::vector<int32_t> &part) {
    int64_t total = 0;
    for (int b = 0; b < nbatches; b++) total += nelem[b];
    part.assign((size_t)total, 0);
    if (nparts <= 1 || total == 0) return;
    if (method == AMARU_PARTITION_METIS) {
        std::vector<int64_t> eptr((size_t)total + 1, 0), eind;
        int64_t g = 0, ncommon = 1 << 30;
        for (int b = 0; b < nbatches; b++) {
            ncommon = std::min<int64_t>(ncommon, facet_nodes(batch_shape[b]));
            for (int64_t e = 0; e < nelem[b]; e++, g++) {
                eptr[(size_t)g + 1] = eptr[(size_t)g] + nn[b];
                for (int a = 0; a < nn[b]; a++) eind.push_back(conn[b][e * nn[b] + a]);
            }
        }
        int64_t ne = total, nnod = nnodes, np = nparts, objval = 0;
        std::vector<int64_t> epart((size_t)total), npart((size_t)nnodes);
        int64_t opt[40];
        METIS_SetDefaultOptions(opt);
        const int rc = METIS_PartMeshDual(&ne, &nnod, eptr.data(), eind.data(), nullptr, nullptr, &ncommon, &np, nullptr, opt,
                                          &objval, epart.data(), npart.data());
        AMARU_REQUIRE(rc == 1, AMARU_ERR_ARG, "METIS_PartMeshDual failed");
        for (int64_t e = 0; e < total; e++) part[(size_t)e] = (int32_t)epart[(size_t)e];
        // METIS may leave a part empty on tiny meshes: fall through to RCB in that case
        std::vector<int64_t> cnt((size_t)nparts, 0);
        for (int64_t e = 0; e < total; e++) cnt[(size_t)part[(size_t)e]]++;
        if (*std::min_element(cnt.begin(), cnt.end()) > 0) return;
    }
    std::vector<double> cent((size_t)total * 3, 0.0);
    int64_t g = 0;
    for (int b = 0; b < nbatches; b++)
        for (int64_t e = 0; e < nelem[b]; e++, g++) {
            double c[3] = {0, 0, 0};
            for (int a = 0; a < nn[b]; a++) {
                const double *x = coords + (size_t)conn[b][e * nn[b] + a] * 3;
                c[0] += x[0]; c[1] += x[1]; c[2] += x[2];
            }
            for (int d = 0; d < 3; d++) cent[(size_t)g * 3 + d] = c[d] / nn[b];
        }
    std::vector<int64_t> idx((size_t)total);
    std::iota(idx.begin(), idx.end(), 0);
    rcb(cent, idx, 0, total, 0, nparts, part.data());
}

void amaru_node_owners(int64_t nnodes, int nbatches, const int *nn, const int64_t *nelem, const int32_t *const *conn,
                       const std::vector<int32_t> &part, std::vector<int32_t> &owner) {
    owner.assign((size_t)nnodes, INT32_MAX);
    int64_t g = 0;
    for (int b = 0; b < nbatches; b++)
        for (int64_t e = 0; e < nelem[b]; e++, g++) {
            const int32_t p = part[(size_t)g];
            for (int a = 0; a < nn[b]; a++) {
                int32_t &o = owner[(size_t)conn[b][e * nn[b] + a]];
                o = std::min(o, p);
            }
        }
}

void amaru_local_view(int rank, int nranks, int64_t nnodes, int nbatches, const int *nn, const int64_t *nelem,
                      const int32_t *const *conn, const std::vector<int32_t> &owner, AmaruLocalView &v) {
    v.rank = rank;
    v.nranks = nranks;
    v.elem_gid.assign((size_t)nbatches, {});
    v.elem_owned.assign((size_t)nbatches, {});
    v.conn.assign((size_t)nbatches, {});
    std::vector<uint8_t> mark((size_t)nnodes, 0);
    for (int b = 0; b < nbatches; b++) {
        for (int64_t e = 0; e < nelem[b]; e++) {
            const int32_t *c = conn[b] + e * nn[b];
            bool mine = false;
            int32_t lo = INT32_MAX;
            for (int a = 0; a < nn[b]; a++) {
                const int32_t o = owner[(size_t)c[a]];
                mine |= o == rank;
                lo = std::min(lo, o);
            }
            if (!mine) continue;
            v.elem_gid[(size_t)b].push_back(e);
            v.elem_owned[(size_t)b].push_back((uint8_t)(lo == rank));   // authoritative copy of the IP state / counted in p.Ap
            for (int a = 0; a < nn[b]; a++) mark[(size_t)c[a]] = 1;
        }
    }
    std::vector<int64_t> owned, ghosts;
    for (int64_t n = 0; n < nnodes; n++)
        if (mark[(size_t)n]) (owner[(size_t)n] == rank ? owned : ghosts).push_back(n);
    std::stable_sort(ghosts.begin(), ghosts.end(), [&](int64_t a, int64_t b) { return owner[(size_t)a] < owner[(size_t)b]; });
    v.nowned = (int64_t)owned.size();
    v.node_gid = owned;
    v.node_gid.insert(v.node_gid.end(), ghosts.begin(), ghosts.end());
    std::vector<int32_t> g2l((size_t)nnodes, -1);
    for (size_t i = 0; i < v.node_gid.size(); i++) g2l[(size_t)v.node_gid[i]] = (int32_t)i;
    for (int b = 0; b < nbatches; b++) {
        auto &lc = v.conn[(size_t)b];
        lc.resize(v.elem_gid[(size_t)b].size() * (size_t)nn[b]);
        for (size_t le = 0; le < v.elem_gid[(size_t)b].size(); le++) {
            const int32_t *c = conn[b] + v.elem_gid[(size_t)b][le] * nn[b];
            for (int a = 0; a < nn[b]; a++) lc[le * nn[b] + a] = g2l[(size_t)c[a]];
        }
    }
    // receive side: one contiguous ghost range per owner
    v.neigh.clear();
    std::vector<int64_t> rstart, rcount;
    for (size_t i = 0; i < ghosts.size();) {
        const int32_t q = owner[(size_t)ghosts[i]];
        size_t j = i;
        while (j < ghosts.size() && owner[(size_t)ghosts[j]] == q) j++;
        v.neigh.push_back(q);
        rstart.push_back(v.nowned + (int64_t)i);
        rcount.push_back((int64_t)(j - i));
        i = j;
    }
    // send side: my nodes of the local elements that also hold a node owned by q
    std::vector<std::vector<int64_t>> send((size_t)nranks);
    for (int b = 0; b < nbatches; b++)
        for (int64_t e : v.elem_gid[(size_t)b]) {
            const int32_t *c = conn[b] + e * nn[b];
            int32_t qs[AMARU_MAXNN];
            int nq = 0;
            for (int a = 0; a < nn[b]; a++) {
                const int32_t o = owner[(size_t)c[a]];
                if (o == rank) continue;
                bool seen = false;
                for (int k = 0; k < nq; k++) seen |= qs[k] == o;
                if (!seen) qs[nq++] = o;
            }
            for (int k = 0; k < nq; k++)
                for (int a = 0; a < nn[b]; a++)
                    if (owner[(size_t)c[a]] == rank) send[(size_t)qs[k]].push_back(c[a]);
        }
    std::vector<int32_t> all = v.neigh;
    for (int q = 0; q < nranks; q++)
        if (!send[(size_t)q].empty()) all.push_back(q);
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    std::vector<int32_t> neigh_recv = v.neigh;
    v.neigh = all;
    v.send_ptr.assign(1, 0);
    v.send_nodes.clear();
    v.recv_start.clear();
    v.recv_count.clear();
    for (int32_t q : all) {
        auto &s = send[(size_t)q];
        std::sort(s.begin(), s.end());
        s.erase(std::unique(s.begin(), s.end()), s.end());
        for (int64_t n : s) v.send_nodes.push_back(g2l[(size_t)n]);
        v.send_ptr.push_back((int64_t)v.send_nodes.size());
        const auto it = std::find(neigh_recv.begin(), neigh_recv.end(), q);
        if (it != neigh_recv.end()) {
            v.recv_start.push_back(rstart[(size_t)(it - neigh_recv.begin())]);
            v.recv_count.push_back(rcount[(size_t)(it - neigh_recv.begin())]);
        } else {
            v.recv_start.push_back(v.nowned);
            v.recv_count.push_back(0);
        }
    }
}
