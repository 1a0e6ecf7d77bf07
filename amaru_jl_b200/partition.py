"""Host-side domain decomposition for the multi-GPU path (one process per GPU) — SURVEY.md §8(e).

The reference has no distributed path (its only parallelism is the shared-memory `@withthreads`, src/tools/threads.jl);
its `src/mesh/partition.jl` is a spatial bin index, not a partitioner.  Here:

* elements are split into `nranks` parts by recursive coordinate bisection of the element centroids (METIS-style k-way
  balance without the graph library: balanced counts, compact parts; deterministic, so every rank computes the same
  partition with no setup communication);
* a node is owned by the lowest-numbered part touching it;
* rank p works on its LOCAL view: owned nodes first (ascending global id), then ghosts grouped by owner (ascending
  global id inside a group); local elements = every element touching an owned node (own + duplicated halo elements), so
  that the CSR rows and internal forces of the owned nodes are complete without communication;
* halo lists: rank p sends to q exactly the nodes q sees as ghosts owned by p, in ascending global id — the order q
  stores them in — so received data lands in place.

Everything is numpy; the output feeds `amaru_create_partitioned` (include/amaru_b200.h).
"""
from __future__ import annotations

import numpy as np


def rcb_partition(points: np.ndarray, nparts: int) -> np.ndarray:
    """Recursive coordinate bisection of `points` (n, d) into `nparts` balanced parts -> part id per point."""
    part = np.zeros(points.shape[0], dtype=np.int32)

    def split(idx, p0, k):
        if k == 1:
            part[idx] = p0
            return
        pts = points[idx]
        axis = int(np.argmax(pts.max(axis=0) - pts.min(axis=0)))
        kl = k // 2
        nl = (idx.size * kl) // k
        order = np.argsort(pts[:, axis], kind="stable")
        split(idx[order[:nl]], p0, kl)
        split(idx[order[nl:]], p0 + kl, k - kl)

    split(np.arange(points.shape[0]), 0, int(nparts))
    return part


class LocalView:
    """Rank-local mesh view + halo lists (all ids local unless named *_gid)."""

    def __init__(self):
        self.rank = self.nranks = 0
        self.node_gid = self.elem_gid = None        # local -> global
        self.nowned = 0
        self.elem_owned = None                      # bool per local element
        self.conn = None                            # local node ids
        self.neigh = self.send_ptr = self.send_nodes = self.recv_start = self.recv_count = None


def partition_mesh(coords: np.ndarray, conn: np.ndarray, nranks: int, rank: int, elem_part: np.ndarray | None = None) -> LocalView:
    nn = conn.shape[1]
    nnodes = coords.shape[0]
    if elem_part is None:
        cent = coords[conn].mean(axis=1)
        elem_part = rcb_partition(cent, nranks)
    elem_part = np.asarray(elem_part, dtype=np.int32)
    owner = np.full(nnodes, np.iinfo(np.int32).max, dtype=np.int32)
    for p in range(int(nranks) - 1, -1, -1):                       # lowest part touching a node wins (assigned last)
        owner[conn[elem_part == p].reshape(-1)] = p
    eown = owner[conn]                                              # (nelem, nn) owners of the element's nodes
    v = LocalView()
    v.rank, v.nranks = rank, nranks
    v.node_owner_global = owner
    v.elem_part = elem_part
    # local elements: any node owned by this rank
    mine = (eown == rank).any(axis=1)
    v.elem_gid = np.nonzero(mine)[0]
    # the IP state of an element is authoritative on the lowest rank owning one of its nodes (that rank always holds it)
    v.elem_owned = eown[v.elem_gid].min(axis=1) == rank
    lnodes = np.unique(conn[v.elem_gid])
    lown = owner[lnodes]
    owned = lnodes[lown == rank]
    ghosts = lnodes[lown != rank]
    gorder = np.lexsort((ghosts, owner[ghosts]))                    # by owner, then global id
    ghosts = ghosts[gorder]
    v.node_gid = np.concatenate((owned, ghosts)).astype(np.int64)
    v.nowned = int(owned.size)
    g2l = np.full(nnodes, -1, dtype=np.int64)
    g2l[v.node_gid] = np.arange(v.node_gid.size)
    v.conn = g2l[conn[v.elem_gid]].astype(np.int32)
    # receive side: contiguous ghost ranges per owner
    gown = owner[ghosts]
    neigh_r = np.unique(gown)
    # send side: pairs (q, b): node b owned by me, sharing an element with a node owned by q != me
    iface = np.nonzero((eown.min(axis=1) != eown.max(axis=1)) & mine)[0]
    send = {}
    if iface.size:
        eo = eown[iface]                                            # (m, nn)
        en = conn[iface]
        for q in np.unique(eo):
            if q == rank:
                continue
            has_q = (eo == q).any(axis=1)
            nodes_q = np.unique(en[has_q][eo[has_q] == rank])      # my nodes in elements that q also touches
            if nodes_q.size:
                send[int(q)] = nodes_q
    neigh = sorted(set(int(q) for q in neigh_r) | set(send.keys()))
    v.neigh = np.array(neigh, dtype=np.int32)
    send_ptr, send_nodes, recv_start, recv_count = [0], [], [], []
    for q in neigh:
        s = send.get(q, np.zeros(0, dtype=np.int64))
        send_nodes.append(g2l[s])
        send_ptr.append(send_ptr[-1] + s.size)
        sel = np.nonzero(gown == q)[0]
        recv_start.append(v.nowned + (int(sel[0]) if sel.size else 0))
        recv_count.append(int(sel.size))
    v.send_ptr = np.array(send_ptr, dtype=np.int64)
    v.send_nodes = (np.concatenate(send_nodes) if send_nodes else np.zeros(0)).astype(np.int32)
    v.recv_start = np.array(recv_start, dtype=np.int64)
    v.recv_count = np.array(recv_count, dtype=np.int64)
    return v


def local_flat(flat: dict, eqid: np.ndarray, view: LocalView):
    """Restrict the flattened model (FEModel.flatten()) and the global eq ids to a rank's local view."""
    out = dict(flat)
    out["coords"] = np.ascontiguousarray(flat["coords"][view.node_gid])
    out["conn"] = np.ascontiguousarray(view.conn)
    out["batch_nelem"] = np.array([view.conn.shape[0]], dtype=np.int64)
    out["elem_mat"] = np.ascontiguousarray(flat["elem_mat"][view.elem_gid])
    eq_local = np.ascontiguousarray(np.asarray(eqid).reshape(flat["coords"].shape[0], -1)[view.node_gid], dtype=np.int32)
    return out, eq_local


def peer_recv_starts(rank: int, neigh, everyone) -> np.ndarray:
    """Addressing of the peer-memory halo push (include/amaru_b200.h, amaru_p2p_connect): for every neighbour q of `rank`,
    the first local node id — in q's numbering — of the ghost range q keeps for `rank`.  `everyone[r]` holds rank r's
    ``neigh`` and ``recv_start`` lists (what the host all-gathers)."""
    return np.array([everyone[int(q)]["recv_start"][list(everyone[int(q)]["neigh"]).index(int(rank))] for q in neigh],
                    dtype=np.int64)
