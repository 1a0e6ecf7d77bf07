// K10: multi-GPU plumbing (halo exchange of ghost entries, scalar all-reduce) — see halo design in DESIGN.md.
// Single-GPU handles never call into here.
#include "amaru_internal.h"

void amaru_halo_exchange(amaru_model *m, double *d_v) {
    (void)d_v;
    if (m->nranks > 1) throw AmaruError{AMARU_ERR_COMM, "halo exchange: communicator not initialised"};
}
void amaru_allreduce_sum(amaru_model *m, double *d_vals, int n) {
    (void)d_vals; (void)n;
    if (m->nranks > 1) throw AmaruError{AMARU_ERR_COMM, "all-reduce: communicator not initialised"};
}
void amaru_allreduce_max_int(amaru_model *m, int *d_val) {
    (void)d_val;
    if (m->nranks > 1) throw AmaruError{AMARU_ERR_COMM, "all-reduce: communicator not initialised"};
}

extern "C" int amaru_nccl_unique_id(void *uid128, char *msg, int msglen) {
    (void)uid128;
    if (msg && msglen > 0) msg[0] = 0;
    return AMARU_ERR_UNSUPPORTED;
}
extern "C" int amaru_create_partitioned(int, int, double, int64_t, int64_t, const double *, const int64_t *, const int32_t *,
                                        int, const int32_t *, const int64_t *, const int32_t *, const int32_t *,
                                        const uint8_t *, int, const int32_t *, const double *, const int32_t *,
                                        const uint8_t *, int64_t, int64_t, int, int, const void *, int,
                                        amaru_model **out, char *msg, int msglen) {
    if (out) *out = nullptr;
    if (msg && msglen > 0) msg[0] = 0;
    return AMARU_ERR_UNSUPPORTED;
}
