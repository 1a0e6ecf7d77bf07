// K13 (patch form): the DMMA matrix-free tangent operator of ebe_mma.cuh with x and y of a PATCH of elements held in
// shared memory.  Included by ebe.cu only (after ebe_mma.cuh: same fragment layouts, same per-IP algebra).
//
// Why (profiles/ncu_ebe_mma_v1_r2.txt, one colour launch of k_ebe_mma at 1 M HEX20): 46 % of the warp-stall samples are
// long-scoreboard waits on the element record and on the read-modify-write of y in global memory, and the launch moves 264 MB
// through DRAM for 115 MB of algorithmic bytes because x and y of a node are re-read once per element touching it.
//
// One WARP owns one patch at a time (patches.cpp: 64 hexahedra / quadrilaterals or 192 tetrahedra, spatially compact):
//   1. ticket = atomicAdd: patches are handed out in colour-major order, so a patch only ever waits for patches with
//      smaller tickets, which are running or finished (no deadlock, whatever the number of resident CTAs);
//   2. wait until the lower-coloured patches sharing a node with this one have published their rows (done[patch] == epoch,
//      ld.acquire.gpu; bounded spin: on a time-out the solve is flagged as broken down instead of hanging the GPU);
//   3. x of the patch's nodes -> shared memory (cp.async), y of the nodes -> shared memory (zero where this patch is the
//      first one to touch the node, else the value the earlier patches left in global memory, ld.global.cg);
//   4. the patch's element groups (8 node-disjoint elements each), exactly k_ebe_mma's group body, with the A fragments of
//      contraction 1 read from the x brick and the result of contraction 2 added into the y brick (patch-local node ids
//      arrive lane-major, one group ahead, as are J⁻¹, coef, w and the element record of the next group);
//   5. y brick -> global memory (plain stores, prescribed rows as zeros), __threadfence, done[patch] = epoch (st.release).
// The accumulation order at a node is: patches in colour order, inside a patch the groups in order — fixed, no atomics on y.
// p·Ap: one partial per PATCH (fixed lane/shuffle order), summed in patch order by the last CTA.
// There is no memset of y and one launch per batch instead of one per element colour.
#pragma once
// (included inside the unnamed namespace of ebe.cu)

struct PatchArgs {
    const int32_t *desc;        // [npatch][8]
    const uint32_t *pnodes;
    const unsigned long long *lane_ids;   // [group][NW][32]
    const int32_t *deps;
    const int32_t *einfo;       // [nslots] material | plastic << 29 | empty << 30 | not-owned << 31
    const int32_t *slot_elem;   // [nslots] (mass term: density lookup)
    const double *geo;          // [(nd*nd+1)][nipp]
    const double *w;            // [6][nipp]
    int64_t nipp;
    const double *dog;
    int nmats;
    const double *dNdR, *Nf, *rho;
    double sa, sb;
    const double *x;
    double *y;
    int mask, npatch;
    unsigned int *ticket, *epoch, *done, *counter;
    double *epatch;             // [npatch]
    CgScalars *scal;
    int dot, first, last, finalize, check_done;
    int fused;                  // multi-GPU fused CG loop (p2p.cuh): wait here for the neighbours' halo of x, push p.Ap to the peers
    P2PFused fz;
};

constexpr int EP_PLASTIC = 1 << 29;
constexpr int EP_EMPTY = 1 << 30;
constexpr int EP_MAT = (1 << 29) - 1;

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int NN, int ND, int NIP, bool MASS, int MAXPN>
struct PatchLayout {
    using M = MmaLayout<NN, ND, NIP, MASS>;
    static constexpr int WARPS = 4;
    static constexpr int NID = M::KS2 + 2 * M::NT3;              // patch-local ids per lane and group
    static constexpr int NW = (NID + 3) / 4;                     // 64-bit words of packed ids per lane and group
    // x and y bricks hold MAXPN rows + one dummy row (zero; never written) that padded fragment positions and empty slots address
    static constexpr size_t warp_bytes = (size_t)(MAXPN + 1) * ND * 8 * 2 + (size_t)MAXPN * 4;
    static constexpr size_t bytes = M::tab_doubles * 8 + 3 * EBE_SMATS * 8 + WARPS * warp_bytes;
};

template <int NN, int ND, int NIP, bool MASS, int MAXPN>
__global__ void __launch_bounds__(128, 2) k_ebe_patch(PatchArgs p) {
    if (p.check_done && p.scal->done) return;
    using L = PatchLayout<NN, ND, NIP, MASS, MAXPN>;
    using M = typename L::M;
    constexpr int IPL = M::IPL, KS2 = M::KS2, NT2 = M::NT2, NG2 = M::NG2, KS3 = M::KS3, NT3 = M::NT3, NW = L::NW, NT = 128;
    extern __shared__ __align__(16) double psm[];
    double *sB2 = psm;                                   // [KS2][NT2][32]
    double *sB3 = sB2 + KS2 * NT2 * 32;                  // [KS3][NT3][32]
    double *sDog = sB3 + KS3 * NT3 * 32;                 // [EBE_SMATS][3]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    char *wbase = reinterpret_cast<char *>(sDog + 3 * EBE_SMATS) + (size_t)wid * L::warp_bytes;
    double *sX = reinterpret_cast<double *>(wbase);      // [MAXPN + 1][ND]
    double *sY = sX + (MAXPN + 1) * ND;                  // [MAXPN + 1][ND]
    uint32_t *sP = reinterpret_cast<uint32_t *>(sY + (MAXPN + 1) * ND);   // [MAXPN] node entries
    if (lane < ND) sX[MAXPN * ND + lane] = sY[MAXPN * ND + lane] = 0.0;   // the dummy row
    const int er = lane >> 2, j = lane & 3;
    __shared__ int s_last;

    // ---- constant operand fragments (same construction as k_ebe_mma)
    for (int f = tid; f < KS2 * NT2 * 32; f += NT) {
        const int l = f & 31, t = f >> 5, s = t / NT2, n = t - s * NT2;
        const int a = 4 * s + (l & 3), c = l >> 2;
        double v = 0.0;
        if (a < NN) {
            if (n < NG2) {
                const int q = NIP == 8 ? c : c >> 1, k = NIP == 8 ? n : 2 * n + (c & 1);
                if (k < ND) v = p.dNdR[(q * NN + a) * ND + k];
            } else if (MASS) {
                const int q = NIP == 8 ? c : c >> 1;
                if (NIP == 8 || !(c & 1)) v = p.Nf[q * NN + a];
            }
        }
        sB2[f] = v;
    }
    for (int f = tid; f < KS3 * NT3 * 32; f += NT) {
        const int l = f & 31, t = f >> 5, s = t / NT3, n = t - s * NT3;
        const int a = 8 * n + (l >> 2);
        const int k = s / IPL, h = s - k * IPL;          // k == ND: mass step
        const int q = IPL == 2 ? 2 * (l & 3) + h : (l & 3);
        double v = 0.0;
        if (a < NN) v = k < ND ? p.dNdR[(q * NN + a) * ND + k] : (MASS ? p.Nf[q * NN + a] : 0.0);
        sB3[f] = v;
    }
    const bool smats = p.nmats <= EBE_SMATS;
    if (smats)
        for (int i = tid; i < 3 * p.nmats; i += NT) sDog[i] = p.dog[i];
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int *>(p.epoch) + 1u;
    unsigned long long halo_epoch = 0;
    if (p.fused) {   // ghost entries of x: every neighbour's push of this exchange has landed (flags in the local window)
        halo_epoch = *reinterpret_cast<volatile unsigned long long *>(&p.fz.pd.win[p.fz.pd.rank]->halo_epoch) + 1ull;
        if (tid < p.fz.nneigh) p2p_wait(p.fz.pd, &p.fz.pd.win[p.fz.pd.rank]->hflag[p.fz.neigh[tid]], halo_epoch);
    }
    __syncthreads();

    double Ji[IPL][ND * ND], coef[IPL], wv[IPL][6];
    unsigned long long ids[NW];
    // J⁻¹, coef and the packed patch-local ids of group `grp` (global group index) -> registers
    auto load_geo = [&](int64_t grp) {
        const int64_t ipl = (grp * 8 + er) * NIP + j * IPL;
#pragma unroll
        for (int t = 0; t < NW; t++) ids[t] = __ldg(p.lane_ids + (grp * NW + t) * 32 + lane);
#pragma unroll
        for (int k = 0; k < ND * ND + 1; k++) {
            const double *src = p.geo + (int64_t)k * p.nipp + ipl;
            if constexpr (IPL == 2) {
                const double2 v = __ldcs(reinterpret_cast<const double2 *>(src));
                if (k < ND * ND) { Ji[0][k] = v.x; Ji[1][k] = v.y; } else { coef[0] = v.x; coef[1] = v.y; }
            } else {
                const double v = __ldcs(src);
                if (k < ND * ND) Ji[0][k] = v; else coef[0] = v;
            }
        }
    };
    // w of the group's integration points, only where the element record says the tangent has a rank-one part
    auto load_w = [&](int64_t grp, int rec) {
        if (rec & EP_PLASTIC) {
            const int64_t ipl = (grp * 8 + er) * NIP + j * IPL;
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const double *src = p.w + (int64_t)c * p.nipp + ipl;
                if constexpr (IPL == 2) {
                    const double2 v = __ldcs(reinterpret_cast<const double2 *>(src));
                    wv[0][c] = v.x; wv[1][c] = v.y;
                } else {
                    wv[0][c] = __ldcs(src);
                }
            }
        }
    };
    auto lid = [&](int t) -> int { return (int)((ids[t >> 2] >> ((t & 3) * 16)) & 0xffffull); };

    // Tickets are taken two patches ahead (the atomic's round trip and the descriptor / node-list loads of the next patch
    // hide behind the current one).  A warp works through its tickets in increasing order, so the holder of the smallest
    // unfinished ticket is always working on it: still no deadlock.
    auto take_raw = [&]() -> int {                       // the atomic's result stays in lane 0 until it is needed
        int t = 0;
        if (lane == 0) t = (int)atomicAdd(p.ticket, 1u);
        return t;
    };
    constexpr int NPL = (MAXPN + 31) / 32;               // node entries per lane
    uint32_t pn[NPL], pq_[NPL];                          // entries of the current / of the next patch (registers)
    auto load_entries = [&](uint32_t (&e)[NPL], int first, int count) {
#pragma unroll
        for (int u = 0; u < NPL; u++) {
            const int i = u * 32 + lane;
            e[u] = i < count ? __ldg(p.pnodes + first + i) : (PN_GHOST | PN_FIRST);
        }
    };
    auto load_desc = [&](int t, int4 &a, int4 &b) {
        if (t < p.npatch) {
            a = __ldg(reinterpret_cast<const int4 *>(p.desc) + 2 * t);
            b = __ldg(reinterpret_cast<const int4 *>(p.desc) + 2 * t + 1);
        }
    };
    auto prefetch_next = [&]() {
#pragma unroll
        for (int u = 0; u < NPL; u++) {
            if (!(pq_[u] & PN_GHOST) || !(pq_[u] & PN_FIRST)) {   // (PN_GHOST | PN_FIRST) marks the padding past the list
                const int64_t o = (int64_t)(pq_[u] & PN_NODE) * ND;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + o));
                if (!(pq_[u] & (PN_FIRST | PN_GHOST))) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.y + o));
            }
        }
    };
    int pt = __shfl_sync(0xffffffffu, take_raw(), 0), pt1 = __shfl_sync(0xffffffffu, take_raw(), 0);
    int4 d0 = make_int4(0, 0, 0, 0), d1 = d0, n0 = d0, n1 = d0;
    load_desc(pt, d0, d1);
    if (pt < p.npatch) load_entries(pn, d0.z, d0.w);
    while (pt < p.npatch) {
        const int grp0 = d0.x, ngrp = d0.y, nnode = d0.w, dep0 = d1.x, ndep = d1.y;
        const int pt2_raw = take_raw();                  // consumed at the end of this patch
        load_desc(pt1, n0, n1);                          // consumed after the first group
        int stage_next = pt1 < p.npatch ? 0 : 2;         // next patch: 0 entries not loaded, 1 loaded, 2 prefetched
        int ei = __ldg(p.einfo + (int64_t)grp0 * 8 + er);
        int ei1 = ngrp > 1 ? __ldg(p.einfo + (int64_t)(grp0 + 1) * 8 + er) : EP_EMPTY;
        load_geo(grp0);                                  // in flight while the bricks are being filled
        // ---- x brick (does not depend on other patches)
#pragma unroll
        for (int u = 0; u < NPL; u++) {                  // the entries are in registers since the previous patch
            const int i = u * 32 + lane;
            if (i < nnode) {
                sP[i] = pn[u];
                const double *src = p.x + (int64_t)(pn[u] & PN_NODE) * ND;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sX + i * ND);
#pragma unroll
                for (int c = 0; c < ND; c++)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + c * 8u), "l"(src + c) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        // ---- wait for the lower-coloured neighbours, then the y brick
        {
            bool ok = true;
            for (int i = lane; i < ndep; i += 32) {
                const unsigned int *f = p.done + __ldg(p.deps + dep0 + i);
                unsigned int spins = 0;
                while (ld_acquire_gpu_u32(f) != epoch) {
                    __nanosleep(64);
                    if (++spins > (1u << 20)) { ok = false; break; }   // ~1 s
                }
            }
            if (!__all_sync(0xffffffffu, ok) && lane == 0) p.scal->done = 3;   // a neighbour never finished: report, do not hang
        }
        __syncwarp();
#pragma unroll
        for (int u0 = 0; u0 < NPL; u0 += 4) {            // 4 nodes per lane in flight
            double v[4][ND];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t ent = u0 + u < NPL ? pn[u0 + u < NPL ? u0 + u : 0] : (PN_FIRST | PN_GHOST);
                const bool ld = !(ent & (PN_FIRST | PN_GHOST));
                const double *src = p.y + (int64_t)(ent & PN_NODE) * ND;
#pragma unroll
                for (int c = 0; c < ND; c++) v[u][c] = ld ? __ldcg(src + c) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = (u0 + u) * 32 + lane;
                if (u0 + u < NPL && i < nnode) {
#pragma unroll
                    for (int c = 0; c < ND; c++) sY[i * ND + c] = v[u][c];
                }
            }
        }
        load_w(grp0, ei);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();

        double en = 0.0;
        for (int g = 0; g < ngrp; g++) {
            const int eig = ei;                          // record of this group's element (registers hold group g)
            const bool act = !(eig & EP_EMPTY);
            const bool plastic = (eig & EP_PLASTIC) != 0;
            int l1[KS2], l3[NT3][2];
#pragma unroll
            for (int s = 0; s < KS2; s++) l1[s] = lid(s);
#pragma unroll
            for (int n = 0; n < NT3; n++) { l3[n][0] = lid(KS2 + 2 * n); l3[n][1] = lid(KS2 + 2 * n + 1); }
            // ---- contraction 1: C[i][n] += A(x; m-tile i, step s) · B2[s][n]
            double C[ND][NT2][2];
#pragma unroll
            for (int i = 0; i < ND; i++)
#pragma unroll
                for (int n = 0; n < NT2; n++) C[i][n][0] = C[i][n][1] = 0.0;
#pragma unroll
            for (int s = 0; s < KS2; s++) {
                double b[NT2];
#pragma unroll
                for (int n = 0; n < NT2; n++) b[n] = sB2[(s * NT2 + n) * 32 + lane];
                // no lane-dependent condition here: mma.sync needs every lane, and a select on `act` invites the compiler to
                // branch around the loads with the DMMAs inside (observed: the 2D instantiations hung).  Empty slots and
                // padded node columns carry the id of the dummy row behind the brick (zero, never written).
#pragma unroll
                for (int i = 0; i < ND; i++) {
                    const double a = sX[l1[s] * ND + i];
#pragma unroll
                    for (int n = 0; n < NT2; n++) dmma(C[i][n][0], C[i][n][1], a, b[n]);
                }
            }
            // ---- per-IP algebra in registers
            const double *dg = (smats ? sDog : p.dog) + 3 * (eig & EP_MAT);
            const double dd = dg[0], oo = dg[1], gg = dg[2];
            double S[IPL][ND * ND], mv[IPL][ND];
            double rho = 0.0;
            if constexpr (MASS) {
                if (act) rho = p.rho[p.slot_elem[(int64_t)(grp0 + g) * 8 + er]];
            }
#pragma unroll
            for (int h = 0; h < IPL; h++) {
                double G[ND * ND], ub[ND];
#pragma unroll
                for (int i = 0; i < ND; i++) {
#pragma unroll
                    for (int k = 0; k < ND; k++) G[i * ND + k] = NIP == 8 ? C[i][k][h] : C[i][k >> 1][k & 1];
                    if constexpr (MASS) ub[i] = C[i][NG2][NIP == 8 ? h : 0];
                    else ub[i] = 0.0;
                }
                if (!plastic) {
#pragma unroll
                    for (int c = 0; c < 6; c++) wv[h][c] = 0.0;
                }
                const double e1 = ebe_ip_algebra<ND, MASS>(G, Ji[h], coef[h], plastic, wv[h], dd, oo, gg, p.sa, coef[h] * p.sb * rho, ub, S[h], mv[h]);
                en += (p.dot && act && eig >= 0) ? e1 : 0.0;   // bit 31 of the record: the element belongs to another rank
            }
            if (g + 1 < ngrp) {                          // the registers of group g are dead: fetch group g + 1 behind contraction 2
                load_geo(grp0 + g + 1);
                load_w(grp0 + g + 1, ei1);               // its record arrived a group ago
                ei = ei1;
                ei1 = g + 2 < ngrp ? __ldg(p.einfo + (int64_t)(grp0 + g + 2) * 8 + er) : EP_EMPTY;
            }
            if (stage_next == 1) {                       // next patch: x and y towards L2 (entries arrived a group ago)
                prefetch_next();
                stage_next = 2;
            } else if (stage_next == 0) {                // next patch: node entries -> registers (its descriptor arrived)
                load_entries(pq_, n0.z, n0.w);
                stage_next = 1;
            }
            // ---- contraction 2: F[i][n] += A(S; step (k,h), m-tile i) · B3[step][n]; the accumulators start from the y brick
            double F[ND][NT3][2];
#pragma unroll
            for (int n = 0; n < NT3; n++)
#pragma unroll
                for (int c = 0; c < 2; c++) {
#pragma unroll
                    for (int i = 0; i < ND; i++) F[i][n][c] = sY[l3[n][c] * ND + i];   // unconditional (see above); only real nodes are stored back
                }
#pragma unroll
            for (int s = 0; s < KS3; s++) {
                const int k = s / IPL, h = s - k * IPL;
                double b[NT3];
#pragma unroll
                for (int n = 0; n < NT3; n++) b[n] = sB3[(s * NT3 + n) * 32 + lane];
#pragma unroll
                for (int i = 0; i < ND; i++) {
                    const double a = k < ND ? S[h][(k < ND ? k : 0) * ND + i] : mv[h][i];
#pragma unroll
                    for (int n = 0; n < NT3; n++) dmma(F[i][n][0], F[i][n][1], a, b[n]);
                }
            }
#pragma unroll
            for (int n = 0; n < NT3; n++)
#pragma unroll
                for (int c = 0; c < 2; c++)
                    if (act && 8 * n + 2 * j + c < NN) {
#pragma unroll
                        for (int i = 0; i < ND; i++) sY[l3[n][c] * ND + i] = F[i][n][c];
                    }
            __syncwarp();                                // the next group's elements touch nodes this group wrote
        }
        // ---- publish the patch's rows
        for (int i = lane; i < nnode; i += 32) {
            const uint32_t ent = sP[i];
            if (!(ent & PN_GHOST)) {
                double *dst = p.y + (int64_t)(ent & PN_NODE) * ND;
#pragma unroll
                for (int c = 0; c < ND; c++) {
                    const bool fx = p.mask && ((ent >> (28 + c)) & 1u);
                    __stcg(dst + c, fx ? 0.0 : sY[i * ND + c]);
                }
            }
        }
        if (p.dot) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) en += __shfl_xor_sync(0xffffffffu, en, o);
            if (lane == 0) __stcg(p.epatch + pt, en);
        }
        __syncwarp();                                    // every lane's stores precede (fence cumulativity) the release below
        if (lane == 0) {
            __threadfence();
            st_release_gpu_u32(p.done + pt, epoch);
        }
        if (stage_next == 0) load_entries(pq_, n0.z, n0.w);   // patches of one group
        pt = pt1;
        pt1 = __shfl_sync(0xffffffffu, pt2_raw, 0);
        d0 = n0; d1 = n1;
#pragma unroll
        for (int u = 0; u < NPL; u++) pn[u] = pq_[u];
    }

    // ---- last CTA: p·Ap in patch order, CG scalars, reset of the ticket counter, epoch of this application
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicInc(p.counter, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (p.dot) {
        double s[1] = {0.0};
        for (int i = tid; i < p.npatch; i += NT) s[0] += __ldcg(p.epatch + i);
        block_sum<1, NT>(s);
        if (tid == 0) {
            const double acc = p.first ? s[0] : p.scal->pq + s[0];   // batches are summed in launch order
            p.scal->pq = acc;
            p.scal->acc[0] = acc;   // multi-GPU: all-reduced in place after the last batch
            if (p.last && p.finalize) {
                if (!(acc > 0.0)) p.scal->done = 3;   // not SPD / breakdown
                p.scal->alpha = p.scal->rz_old / acc;
            }
        }
    }
    if (p.fused && p.last) {
        P2PWin *me = p.fz.pd.win[p.fz.pd.rank];
        if (p.dot) {   // this rank's p.Ap -> every rank's window; the vector-update kernel sums the slots in rank order
            __syncthreads();
            const unsigned long long se = *reinterpret_cast<volatile unsigned long long *>(&me->scal_epoch) + 1ull;
            if (tid < 32) p2p_push_scalars(p.fz.pd, se, p.scal->acc, 1, tid);
        }
        if (tid == 0) me->halo_epoch = halo_epoch;
    }
    if (tid == 0) {
        *p.ticket = 0u;
        *p.epoch = epoch;
    }
}
