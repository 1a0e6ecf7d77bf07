// Host-side domain decomposition (partition.cpp) — internal to libamaru_b200.so.
#pragma once
#include "amaru_internal.h"

struct AmaruLocalView {
    int rank = 0, nranks = 1;
    std::vector<int64_t> node_gid;                 // local node -> global node (owned first, then ghosts grouped by owner)
    int64_t nowned = 0;
    std::vector<std::vector<int64_t>> elem_gid;    // per batch: local element -> element index inside the global batch
    std::vector<std::vector<uint8_t>> elem_owned;  // per batch: lowest rank owning one of the element's nodes == this rank
    std::vector<std::vector<int32_t>> conn;        // per batch: local node ids
    std::vector<int32_t> neigh;
    std::vector<int64_t> send_ptr, recv_start, recv_count;
    std::vector<int32_t> send_nodes;
};

// element -> part (AMARU_PARTITION_RCB | AMARU_PARTITION_METIS); `conn[b]` = connectivity of batch b
void amaru_partition_elements(int method, int nparts, int64_t nnodes, const double *coords, int nbatches, const int *nn,
                              const int32_t *batch_shape, const int64_t *nelem, const int32_t *const *conn,
                              std::vector<int32_t> &part);
// owner[node] = lowest part touching the node
void amaru_node_owners(int64_t nnodes, int nbatches, const int *nn, const int64_t *nelem, const int32_t *const *conn,
                       const std::vector<int32_t> &part, std::vector<int32_t> &owner);
void amaru_local_view(int rank, int nranks, int64_t nnodes, int nbatches, const int *nn, const int64_t *nelem,
                      const int32_t *const *conn, const std::vector<int32_t> &owner, AmaruLocalView &v);
