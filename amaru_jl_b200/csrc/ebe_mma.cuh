// K13 (tensor-core form): the matrix-free tangent operator of ebe.cu with both element contractions on the FP64 tensor
// cores (mma.sync m8n8k4, DMMA).  Included by ebe.cu only.
//
// Why DMMA (BASELINE north_star: "only if ncu shows the batched BᵀDB contraction benefits"): the DFMA version
// (k_ebe_apply) is bound by the shared-memory pipe, not by FP64 issue — ncu at 1 M HEX20 (profiles/ncu_ebe_v3_r2.txt):
// l1tex data-pipe wavefronts 69 % of peak, FP64 pipe 29 %, 438 warp instructions per element.  Every DFMA of the two
// contractions  G = Σ_a x_a ⊗ ∂N_a/∂R  and  f_a = Σ_q,k ∂N_a/∂R(q)[k]·S(q)[k][:]  needs two shared-memory operands per 1.5
// FMAs.  DMMA has the same peak as DFMA on B200 (37.0 TFLOP/s measured, profiles/fp64_peak_r2.json) but takes its
// operands as register fragments, and the shape-function table is the SAME for every element: its fragments are loaded
// once per warp and group, the element-side operand of the second contraction never leaves the registers.
//
// One WARP owns a group of 8 elements (no CTA-level barrier inside the loop):
//   rows of both products     R = i·8 + e           (i = displacement component, e = element of the group): m-tile i
//   contraction 1  G = X·B2   K = node a (padded to 4), N = (k, q) arranged so that lane (e = l/4, j = l%4) receives, in
//                             its accumulator registers, the complete ND×ND gradient G[i][k] of ITS integration points:
//                             NIP = 8: n-tile k, column c = q      -> lane holds q = 2j, 2j+1 (two IPs per lane)
//                             NIP = 4: n-tile t, column c = 2q + (k − 2t) -> lane holds q = j   (one IP per lane)
//   per-IP algebra in registers: H = G·J⁻¹, ε = sym H (Mandel), σ = coef·a·(De − w wᵀ)ε, S = J⁻¹·T(σ), energy ε·σ
//   contraction 2  f = S·B3   K = (k, h) steps whose four K-slots are the lanes' own IPs (q = 2j + h, or q = j), so the
//                             A fragment of step (k, h), m-tile i is simply S[h][k][i] of the lane: no data movement;
//                             N = node a (padded to 8): lane (e, j) receives f of nodes 8n + 2j, 8n + 2j + 1 for all i
//   y[node] += f              read-modify-write of 3 consecutive doubles per node; elements of a colour share no node
// x of the group's nodes is gathered by cp.async into a double-buffered, conflict-free stage (row stride ≡ 4 mod 8 doubles);
// J⁻¹, coef and w are read straight from global memory as 16-byte lane-contiguous loads (a warp reads 512 contiguous bytes
// per plane), issued before contraction 1 so that their latency hides behind it; so are the y entries the lane will
// store (ld.global.cg: a colour reads rows the previous colour wrote, and the colours run inside ONE persistent
// cooperative launch separated by a grid barrier — 8 launches of ~20 µs each were the floor of small batches).
#pragma once
// (included inside the unnamed namespace of ebe.cu, after EbeArgs)

// 1/√2: the Mandel shear factors are applied as multiplications here (a correctly rounded FP64 division costs ~15
// instructions; the operator only has to agree with the assembled tangent to CG accuracy, K itself keeps the reference's
// divisions by SR2, assemble.cu / materials.cuh)
#define MMA_ISR2 0.70710678118654752440

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NN, int ND, int NIP, bool MASS>
struct MmaLayout {
    static constexpr int EW = 8;                                  // elements per warp group
    static constexpr int IPL = NIP / 4;                           // integration points per lane (8 elements x NIP / 32)
    static constexpr int KS2 = (NN + 3) / 4;                      // k-steps of contraction 1 (nodes)
    static constexpr int NT2 = (NIP == 8 ? ND : (ND + 1) / 2) + (MASS ? 1 : 0);   // n-tiles of contraction 1
    static constexpr int NG2 = NIP == 8 ? ND : (ND + 1) / 2;      // of which gradient tiles
    static constexpr int KS3 = ND * IPL + (MASS ? IPL : 0);       // k-steps of contraction 2
    static constexpr int NT3 = (NN + 7) / 8;                      // n-tiles of contraction 2 (nodes)
    static constexpr int K2 = KS2 * 4;
    static constexpr int RSX = (K2 % 8 == 4) ? K2 : K2 + 4;       // row stride of the x stage: ≡ 4 (mod 8) doubles
    static constexpr int XW = ND * EW * RSX;                      // doubles of one x stage
    static constexpr int NW = EW * NN;                            // econn entries of one stage
    static constexpr int WARPS = 4;
    static constexpr size_t tab_doubles = (size_t)(KS2 * NT2 + KS3 * NT3) * 32;
    static constexpr size_t bytes = tab_doubles * 8 + (size_t)WARPS * 2 * XW * 8 + (size_t)WARPS * 2 * NW * 4 + 3 * EBE_SMATS * 8;
    static_assert(NIP == 4 || NIP == 8, "lane <-> integration point mapping needs 4 or 8 integration points");
};

// Per-IP algebra shared by both lane mappings: G (row-major [i][k]) -> S ([k][i]) and the energy; ub -> mv (mass term)
template <int ND, bool MASS>
__device__ __forceinline__ double ebe_ip_algebra(const double (&G)[ND * ND], const double (&Ji)[ND * ND], double coef, bool plastic,
                                                 const double (&w)[6], double dd, double oo, double gg, double sa, double cm,
                                                 const double (&ub)[ND], double (&S)[ND * ND], double (&mv)[ND]) {
    double H[ND * ND];
#pragma unroll
    for (int i = 0; i < ND; i++)
#pragma unroll
        for (int j = 0; j < ND; j++) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < ND; k++) v += G[i * ND + k] * Ji[k * ND + j];
            H[i * ND + j] = v;
        }
    double ep[6], sg[6];
    if constexpr (ND == 3) {
        ep[0] = H[0]; ep[1] = H[4]; ep[2] = H[8];
        ep[3] = (H[5] + H[7]) * MMA_ISR2; ep[4] = (H[2] + H[6]) * MMA_ISR2; ep[5] = (H[1] + H[3]) * MMA_ISR2;
    } else {
        ep[0] = H[0]; ep[1] = H[3]; ep[2] = 0.0; ep[3] = 0.0; ep[4] = 0.0; ep[5] = (H[1] + H[2]) * MMA_ISR2;
    }
    sg[0] = dd * ep[0] + oo * ep[1] + oo * ep[2];
    sg[1] = oo * ep[0] + dd * ep[1] + oo * ep[2];
    sg[2] = oo * ep[0] + oo * ep[1] + dd * ep[2];
    sg[3] = gg * ep[3]; sg[4] = gg * ep[4]; sg[5] = gg * ep[5];
    if (plastic) {
        double t = 0.0;
#pragma unroll
        for (int c = 0; c < 6; c++) t += w[c] * ep[c];
#pragma unroll
        for (int c = 0; c < 6; c++) sg[c] -= w[c] * t;
    }
    const double ca = coef * sa;
    double en = 0.0;
#pragma unroll
    for (int c = 0; c < 6; c++) {
        sg[c] *= ca;
        en += ep[c] * sg[c];
    }
    if constexpr (ND == 3) {
        const double T[9] = {sg[0], sg[5] * MMA_ISR2, sg[4] * MMA_ISR2, sg[5] * MMA_ISR2, sg[1], sg[3] * MMA_ISR2,
                             sg[4] * MMA_ISR2, sg[3] * MMA_ISR2, sg[2]};
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int i = 0; i < 3; i++) S[k * 3 + i] = Ji[k * 3] * T[i * 3] + Ji[k * 3 + 1] * T[i * 3 + 1] + Ji[k * 3 + 2] * T[i * 3 + 2];
    } else {
        const double T[4] = {sg[0], sg[5] * MMA_ISR2, sg[5] * MMA_ISR2, sg[1]};
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
            for (int i = 0; i < 2; i++) S[k * 2 + i] = Ji[k * 2] * T[i * 2] + Ji[k * 2 + 1] * T[i * 2 + 1];
    }
    if (MASS) {
#pragma unroll
        for (int i = 0; i < ND; i++) {
            mv[i] = cm * ub[i];
            en += cm * ub[i] * ub[i];
        }
    }
    return en;
}

// Grid-wide barrier of the persistent colour loop (all CTAs are resident: cooperative launch).  bar[0] counts arrivals
// (monotonic inside one kernel), bar[1] publishes the completed generation.  Bounded: on a time-out the solve is flagged as broken
// down instead of hanging the GPU.
__device__ __forceinline__ void ebe_grid_barrier(unsigned int *bar, unsigned int gen, CgScalars *scal) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&bar[0], 1u) == gen * gridDim.x - 1u) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen) : "memory");
        } else {
            unsigned int v, spins = 0;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar + 1) : "memory");
                if (v < gen) __nanosleep(32);
            } while (v < gen && ++spins < (1u << 22));
            if (v < gen) scal->done = 3;
        }
    }
    __syncthreads();
}

template <int NN, int ND, int NIP, bool MASS>
__global__ void __launch_bounds__(128, 2) k_ebe_mma(EbeArgs p) {
    if (p.check_done && p.scal->done) return;
    using L = MmaLayout<NN, ND, NIP, MASS>;
    constexpr int EW = L::EW, IPL = L::IPL, KS2 = L::KS2, NT2 = L::NT2, NG2 = L::NG2, KS3 = L::KS3, NT3 = L::NT3, RSX = L::RSX,
                  XW = L::XW, NW = L::NW, NT = 128;
    constexpr int NLD = (NW + 31) / 32;                  // econn entries per lane and group
    extern __shared__ __align__(16) double msm[];
    double *sB2 = msm;                                   // [KS2][NT2][32]
    double *sB3 = sB2 + KS2 * NT2 * 32;                  // [KS3][NT3][32]
    double *sDog = sB3 + KS3 * NT3 * 32;                 // [EBE_SMATS][3]
    double *sXall = sDog + 3 * EBE_SMATS;                // [WARPS][2][ND][EW][RSX]
    int32_t *sNall = reinterpret_cast<int32_t *>(sXall + L::WARPS * 2 * XW);   // [WARPS][2][EW][NN]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int er = lane >> 2, j = lane & 3;              // this lane's element of the group / position in the quad
    __shared__ int s_last;

    // ---- constant operand fragments (same for every element): B2[a][col], B3[ip slot][a]
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
    double *sX = sXall + wid * 2 * XW;
    int32_t *sN = sNall + wid * 2 * NW;
    for (int i = lane; i < 2 * XW; i += 32) sX[i] = 0.0;  // padded node columns stay zero for the whole kernel
    unsigned long long halo_epoch = 0;
    if (p.fused) {   // ghost entries of x: every neighbour's push of this exchange has landed (flags in the local window)
        halo_epoch = *reinterpret_cast<volatile unsigned long long *>(&p.fz.pd.win[p.fz.pd.rank]->halo_epoch) + 1ull;
        if (tid < p.fz.nneigh) p2p_wait(p.fz.pd, &p.fz.pd.win[p.fz.pd.rank]->hflag[p.fz.neigh[tid]], halo_epoch);
    }
    __syncthreads();

    double dsum[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * L::WARPS;
    // ---- the element colours of the batch, one after the other (elements of a colour share no node); a grid-wide barrier
    // separates two colours: the rows a colour wrote are read by the next one (through L2: ld.global.cg / st.global.cg)
    int64_t e_begin = p.cb[0], e_end = p.cb[1], ngroups = (e_end - e_begin + EW - 1) / EW;
    int ei_next = 0;
    for (int col = 0; col < p.ncol; col++) {
        int32_t nreg[NLD];
        auto load_conn = [&](int64_t g) {
            const int64_t ge = e_begin + g * EW;
#pragma unroll
            for (int t = 0; t < NLD; t++) {
                const int i = lane + t * 32;
                nreg[t] = (g < ngroups && i < NW && ge * NN + i < e_end * NN) ? p.econn[ge * NN + i] : -1;
            }
        };
        auto issue_copies = [&](int buf) {               // x of the nodes in nreg -> stage buf (rows past the end: zero)
#pragma unroll
            for (int t = 0; t < NLD; t++) {
                const int i = lane + t * 32;
                if (i < NW) {
                    const int e = i / NN, a = i - e * NN;
                    const int32_t ent = nreg[t];
                    sN[buf * NW + i] = ent;
                    double *dst = sX + buf * XW + e * RSX + a;
                    if (ent != -1) {
                        const double *src = p.x + (int64_t)((uint32_t)ent & EC_NODE) * ND;
                        const uint32_t d32 = (uint32_t)__cvta_generic_to_shared(dst);
#pragma unroll
                        for (int d = 0; d < ND; d++)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d32 + (uint32_t)(d * EW * RSX) * 8u), "l"(src + d) : "memory");
                    } else {
#pragma unroll
                        for (int d = 0; d < ND; d++) dst[d * EW * RSX] = 0.0;
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto load_einfo = [&](int64_t gg) -> int {       // record of this lane's element in group gg (0 past the end)
            const int64_t e = e_begin + gg * EW + er;
            return (gg < ngroups && e < e_end) ? p.einfo[e] : 0;
        };
        int64_t g = (int64_t)blockIdx.x * L::WARPS + wid;
        // prologue of a colour: connectivity and x of the warp's first group (x is read-only: the prologue of colour c + 1 runs
        // before the grid barrier that ends colour c, so its latency hides behind the wait)
        auto prologue = [&]() {
            __syncwarp();                                // the stages of the previous colour are free
            load_conn(g);
            issue_copies(0);
            ei_next = load_einfo(g);
        };
        if (col == 0) prologue();
        load_conn(g + stride);
        int buf = 0;
        for (; g < ngroups; g += stride, buf ^= 1) {
            const int64_t e0 = e_begin + g * EW;
            const int ne = (int)min((int64_t)EW, e_end - e0);
            const bool act = er < ne;                    // this lane's element exists
            // ---- geometry / tangent of this lane's integration points: issued first, consumed after contraction 1
            const int64_t ipl = (e0 + er) * NIP + j * IPL;   // first IP of the lane inside the batch
            double Ji[IPL][ND * ND], coef[IPL], wv[IPL][6];
            const int ei = ei_next;                      // loaded one group ago
            ei_next = load_einfo(g + stride);
            if (act) {
#pragma unroll
                for (int k = 0; k < ND * ND + 1; k++) {
                    const double *src = p.geo + (int64_t)k * p.nipb + ipl;
                    if constexpr (IPL == 2) {
                        const double2 v = __ldcs(reinterpret_cast<const double2 *>(src));
                        if (k < ND * ND) { Ji[0][k] = v.x; Ji[1][k] = v.y; } else { coef[0] = v.x; coef[1] = v.y; }
                    } else {
                        const double v = __ldcs(src);
                        if (k < ND * ND) Ji[0][k] = v; else coef[0] = v;
                    }
                }
            } else {
#pragma unroll
                for (int h = 0; h < IPL; h++) {
                    coef[h] = 0.0;
#pragma unroll
                    for (int k = 0; k < ND * ND; k++) Ji[h][k] = 0.0;
                }
            }
            const bool plastic = act && (ei & EI_PLASTIC) != 0;
            if (plastic) {
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    const double *src = p.w + (int64_t)c * p.nip_total + p.ip_off + ipl;
                    if constexpr (IPL == 2) {
                        const double2 v = __ldcs(reinterpret_cast<const double2 *>(src));
                        wv[0][c] = v.x; wv[1][c] = v.y;
                    } else {
                        wv[0][c] = __ldcs(src);
                    }
                }
            } else {
#pragma unroll
                for (int h = 0; h < IPL; h++)
#pragma unroll
                    for (int c = 0; c < 6; c++) wv[h][c] = 0.0;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();                                // stage `buf` is complete; everybody left the previous group
            // ---- y of the nodes this lane will store: issued now, consumed by contraction 2 (the latency hides behind
            // contraction 1); unconditional loads (row 0 for lanes that store nothing): no select in front of mma.sync
            double F[ND][NT3][2];
            int64_t yk[NT3][2];
            uint32_t skip = 0;
#pragma unroll
            for (int n = 0; n < NT3; n++)
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const int a = 8 * n + 2 * j + c;
                    const uint32_t ent = (act && a < NN) ? (uint32_t)sN[buf * NW + er * NN + a] : 0x80000000u;
                    const bool st = !(ent >> 31);        // ghost rows belong to the neighbour rank
                    yk[n][c] = st ? (int64_t)(ent & EC_NODE) * ND : -1;
                    if (p.mask) skip |= ((ent >> 28) & 7u) << ((n * 2 + c) * ND);
                    const double *ysrc = p.y + (st ? yk[n][c] : 0);
#pragma unroll
                    for (int i = 0; i < ND; i++) F[i][n][c] = __ldcg(ysrc + i);
                }
            issue_copies(buf ^ 1);                       // nreg holds the entries of group g + stride
            load_conn(g + 2 * stride);

            // ---- contraction 1: C[i][n] += A(x; m-tile i, step s) · B2[s][n]
            double C[ND][NT2][2];
#pragma unroll
            for (int i = 0; i < ND; i++)
#pragma unroll
                for (int n = 0; n < NT2; n++) C[i][n][0] = C[i][n][1] = 0.0;
            const double *xa = sX + buf * XW + er * RSX + j;
#pragma unroll
            for (int s = 0; s < KS2; s++) {
                double b[NT2];
#pragma unroll
                for (int n = 0; n < NT2; n++) b[n] = sB2[(s * NT2 + n) * 32 + lane];
#pragma unroll
                for (int i = 0; i < ND; i++) {
                    const double a = xa[i * EW * RSX + 4 * s];
#pragma unroll
                    for (int n = 0; n < NT2; n++) dmma(C[i][n][0], C[i][n][1], a, b[n]);
                }
            }
            // ---- per-IP algebra in registers
            const double *dg = (smats ? sDog : p.dog) + 3 * (ei & EI_MAT);
            const double dd = dg[0], oo = dg[1], gg = dg[2];
            double S[IPL][ND * ND], mv[IPL][ND];
            const double rho = (MASS && act) ? p.rho[e0 + er] : 0.0;
            double en = 0.0;
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
                en += ebe_ip_algebra<ND, MASS>(G, Ji[h], coef[h], plastic, wv[h], dd, oo, gg, p.sa, coef[h] * p.sb * rho, ub, S[h], mv[h]);
            }
            dsum[0] += (p.dot && act && ei >= 0) ? en : 0.0;   // bit 31 of the record: the element belongs to another rank
            // ---- contraction 2: F[i][n] += A(S; step (k,h), m-tile i) · B3[step][n]; the accumulators started from y
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
                    if (yk[n][c] >= 0) {
#pragma unroll
                        for (int i = 0; i < ND; i++)
                            if (!((skip >> ((n * 2 + c) * ND + i)) & 1u)) __stcg(p.y + yk[n][c] + i, F[i][n][c]);
                    }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (col + 1 < p.ncol) {
            e_begin = p.cb[col + 1];
            e_end = p.cb[col + 2];
            ngroups = (e_end - e_begin + EW - 1) / EW;
            g = (int64_t)blockIdx.x * L::WARPS + wid;
            prologue();
            ebe_grid_barrier(p.bar, (unsigned int)(col + 1), p.scal);
        }
    }

    // ---- last CTA: p·Ap (per-CTA partials summed in CTA order), CG scalars, barrier reset, fused exchanges
    if (p.dot) {
        block_sum<1, NT>(dsum);
        if (tid == 0) p.partial[blockIdx.x] = dsum[0];
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicInc(&p.scal->counter[0], gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (p.dot) {
        sum_partials<1, NT>(dsum, p.partial);
        if (tid == 0) {
            const double acc = p.first ? dsum[0] : p.scal->pq + dsum[0];   // batches are summed in launch order
            p.scal->pq = acc;
            p.scal->acc[0] = acc;   // multi-GPU: all-reduced after the last batch
            if (p.last && p.finalize) {
                if (!(acc > 0.0)) p.scal->done = 3;   // not SPD / breakdown
                p.scal->alpha = p.scal->rz_old / acc;
            }
        }
    }
    if (p.fused && p.last) {   // this rank's p.Ap -> every rank's window (summed in rank order by the vector update)
        __syncthreads();
        P2PWin *me = p.fz.pd.win[p.fz.pd.rank];
        if (p.dot) {
            const unsigned long long se = *reinterpret_cast<volatile unsigned long long *>(&me->scal_epoch) + 1ull;
            if (tid < 32) p2p_push_scalars(p.fz.pd, se, p.scal->acc, 1, tid);
        }
        if (tid == 0) me->halo_epoch = halo_epoch;
    }
    if (tid == 0) {
        p.bar[0] = 0u;
        p.bar[1] = 0u;
    }
}
