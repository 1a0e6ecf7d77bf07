// Peer-memory exchange primitives shared by halo.cu (stand-alone exchange kernels), pcg.cu and ebe_patch.cuh (the fused CG
// loop: exchanges folded into the kernels that produce / consume the data).  See halo.cu for the protocol.
#pragma once
#include <stdint.h>

constexpr int P2P_MAXR = 16;
struct P2PWin {
    double slot[2][P2P_MAXR][4];
    unsigned long long sflag[2][P2P_MAXR];
    unsigned long long hflag[P2P_MAXR];
    unsigned long long halo_epoch, scal_epoch;   // completed epochs of this rank
    int abort;
};
struct P2PDev {
    int rank, nranks;
    unsigned long long timeout_ns;
    P2PWin *win[P2P_MAXR];
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// wait until *flag >= epoch; false (and the abort flag raised everywhere) when the budget ran out or a peer aborted
__device__ __forceinline__ bool p2p_wait(const P2PDev &pd, const unsigned long long *flag, unsigned long long epoch) {
    P2PWin *me = pd.win[pd.rank];
    const unsigned long long t0 = global_ns();
    unsigned int it = 0;
    while (ld_acquire_sys(flag) < epoch) {
        if ((++it & 255u) == 0) {
            const bool late = global_ns() - t0 > pd.timeout_ns;
            if (late || *reinterpret_cast<volatile int *>(&me->abort)) {
                for (int r = 0; r < pd.nranks; r++) *reinterpret_cast<volatile int *>(&pd.win[r]->abort) = 1;
                __threadfence_system();
                return false;
            }
        }
        __nanosleep(20);
    }
    return true;
}


// ---- fused CG loop --------------------------------------------------------------------------------------------------------
// Scalar all-reduce split over two kernels: the LAST CTA of the producing kernel stores this rank's partial values into every
// rank's window (p2p_push_scalars), EVERY CTA of the consuming kernel waits for all ranks' flags in its own (local) window and
// sums the slots in rank order (p2p_collect) — bitwise the same result on every rank and in every CTA.  The consuming kernel's
// last CTA advances scal_epoch.  Halo: the kernel that writes p pushes the boundary entries into the neighbours' ghost slots
// (per-node table) and its last CTA raises the epoch flags; the operator kernel's CTAs wait for the neighbours' flags before
// they read ghost entries, its last CTA advances halo_epoch.
struct P2PFused {
    P2PDev pd;
    int nneigh;
    const int *neigh;               // [nneigh] neighbour ranks
    double *const *peer_p;          // [nneigh] neighbours' p vectors
    const int32_t *bidx;            // [nowned] index of the node in the boundary table or -1
    const int32_t *bent_ptr;        // [nboundary + 1]
    const int32_t *bent_q;          // [entries] neighbour index
    const int64_t *bent_remote;     // [entries] node index in that neighbour's numbering
    unsigned int *counter;          // last-CTA detection of the vector kernels
};

struct amaru_model;
P2PFused amaru_comm_fused_args(amaru_model *m);   // halo.cu

// lanes 0..nranks-1 of one warp
__device__ __forceinline__ void p2p_push_scalars(const P2PDev &pd, unsigned long long epoch, const double *vals, int n, int lane) {
    if (lane < pd.nranks) {
        P2PWin *w = pd.win[lane];
        const int par = (int)(epoch & 1ull);
        for (int k = 0; k < n; k++) w->slot[par][pd.rank][k] = vals[k];
        __threadfence_system();
        st_release_sys(&w->sflag[par][pd.rank], epoch);
    }
}
// first warp of a CTA (all 32 lanes): lane r waits for rank r's flag (the waits overlap) and reads rank r's slot itself — the
// lane that acquired the flag is the one that reads the data behind it —, then the values are summed in rank order through
// shuffles; out[0..n) valid in every lane; false on time-out / abort
__device__ __forceinline__ bool p2p_collect(const P2PDev &pd, unsigned long long epoch, double *out, int n) {
    P2PWin *me = pd.win[pd.rank];
    const int par = (int)(epoch & 1ull);
    const int lane = threadIdx.x & 31;
    bool ok = true;
    double mine[4] = {0.0, 0.0, 0.0, 0.0};
    if (lane < pd.nranks) {
        ok = p2p_wait(pd, &me->sflag[par][lane], epoch);
        if (ok)
            for (int k = 0; k < n; k++) mine[k] = *reinterpret_cast<volatile double *>(&me->slot[par][lane][k]);
    }
    ok = __all_sync(0xffffffffu, ok);
    if (!ok) return false;
    for (int k = 0; k < n; k++) {
        double s = __shfl_sync(0xffffffffu, mine[k], 0);
        for (int r = 1; r < pd.nranks; r++) s += __shfl_sync(0xffffffffu, mine[k], r);
        out[k] = s;
    }
    return true;
}
