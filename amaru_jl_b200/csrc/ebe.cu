// K13: matrix-free (element-by-element) tangent operator  y = (a·K + b·M) x  (+ fused p·Ap) for the PCG that replaces
// lu(K11) (reference src/solver.jl:38-43).  K is the matrix mount_K builds (src/mech/mech-solver.jl:78-110 with
// elem_stiffness, src/mech/elem/mech-solid.jl:124-166); instead of streaming its 18 KB per HEX20 element every CG
// iteration (block-CSR SpMV, spmv.cu) the product is re-integrated from 80-128 bytes per integration point:
//
//   per IP (constant over the analysis, small strains): J⁻¹ (nd² doubles) and coef = detJ·w·th
//   per IP (refreshed by amaru_assemble_K):             w (6 doubles) with  D = De − w wᵀ
//
// All three materials of the path have a tangent of that form (associated flow): von Mises  w = De·n/√(n·De·n + √1.5·H)
// (von-mises.jl:112-125), Drucker–Prager  w = De·V/√(|V|·(V·De·V/|V| + H)) (drucker-prager.jl:86-109), elastic w = 0.
// With G = Σ_a x_a ⊗ ∂N_a/∂R the displacement gradient is H = G·J⁻¹, ε = sym(H) in Mandel order (setB,
// mech-solid.jl:82-121), σ = coef·a·D·ε, and the nodal force is f_a = Σ_q ∂N_a/∂R(q) · S(q) with S = J⁻¹·T(σ): the shape
// table is shared by all elements, so both contractions are small GEMMs against a constant operand held in shared memory.
//
// Both contractions run on the FP64 tensor cores (DMMA, ebe_mma.cuh: one warp = 8 elements, accumulator registers = the
// lane's own integration points).  Two forms, chosen per batch at amaru_create:
//   k_ebe_patch (ebe_patch.cuh)  x / y of a 64-element patch in shared memory, ticketed colour-major patch order with epoch
//                                flags; for batches with enough patches to hide the 8-deep chain of neighbouring patches;
//   k_ebe_mma   (ebe_mma.cuh)    element groups straight from global memory, ONE persistent cooperative launch that walks
//                                through the element colours with a grid barrier between two colours (elements of a colour
//                                share no node: plain read-modify-write of y, fixed order, no atomics — same determinism
//                                argument as assemble.cu); for small batches (multi-GPU strong scaling) and poor patch fills.
// The earlier DFMA form (one thread per integration point, shape table in shared memory) was retired after the A/B of
// it was bound by the shared-memory pipe (0.97 ms per application at 1 M HEX20 against 0.605 ms).
// p·Ap = Σ_e Σ_q coef·εᵀDε over the elements this rank owns (p vanishes on prescribed dofs, so this is the masked dot);
// partial sums are combined in a fixed order by the last CTA of a launch.
//
// Algorithmic bytes per application: 8·(nd²+1)·nip + 48·nip(plastic elements) + (2|4)·nn·nelem + 8n + 8n.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "materials.cuh"
#include "p2p.cuh"
#include "reduce.cuh"

namespace {

constexpr uint32_t EC_NODE = 0x0fffffffu;   // econn entry: node | prescribed-dof mask << 28 | ghost << 31
constexpr int EI_PLASTIC = 1 << 30;          // einfo entry: material | plastic << 30 | not-owned << 31
constexpr int EI_MAT = (1 << 30) - 1;
constexpr int EBE_MAXCOL = 96;               // element colours one persistent launch walks through (more: several launches)
constexpr int EBE_SMATS = 32;                // material moduli staged in shared memory up to this many materials

struct EbeBatch {          // per element batch, colour-sorted element order (same order as the IP state planes)
    double *d_geo = nullptr;      // [(nd*nd+1)][nipb]: J⁻¹ row-major planes, then coef
    int32_t *d_econn = nullptr;   // [nelem*nn] node | prescribed-dof mask << 28 | ghost << 31
    int32_t *d_einfo = nullptr;   // [nelem] material | (some IP has w != 0) << 30 | (element owned by another rank) << 31
    int grid_mma = 1;             // persistent (co-resident) grid of the colour-ordered DMMA kernel (k_ebe_mma)
    unsigned int *d_bar = nullptr;   // its grid barrier: arrivals, completed generation
    // patch form (k_ebe_patch, ebe_patch.cuh): arrays in SLOT order (8 element slots per group, groups patch by patch)
    bool patch = false;
    int npatch = 0, grid_patch = 1;
    int64_t nslots = 0;
    double fill = 0.0;
    int64_t pnode_total = 0, pnode_loaded = 0;   // node entries of all patches / those that read y back (not first-touch, not ghost)
    int32_t *d_desc = nullptr, *d_deps = nullptr, *d_pinfo = nullptr, *d_slot_elem = nullptr, *d_elem_slot = nullptr;
    uint32_t *d_pnodes = nullptr;
    unsigned long long *d_lane_ids = nullptr;
    double *d_pgeo = nullptr;     // [(nd*nd+1)][nslots*nip]
    double *d_pw = nullptr;       // [6][nslots*nip]
    unsigned int *d_sync = nullptr;   // [0] ticket, [1] epoch, [2..2+npatch) done flags
    double *d_epatch = nullptr;   // [npatch] p.Ap partial of every patch
    std::vector<int32_t> h_elem_slot;   // colour-sorted element -> slot (host copy for amaru_ebe_set_owned)
};

struct Ebe {
    std::vector<EbeBatch> b;
    double *d_w = nullptr;        // [6][nip_total]
    double *d_dog = nullptr;      // [nmats][3]: c(1-ν), cν, c(1-2ν)
    int64_t *d_nplastic = nullptr;   // device counter of IPs in flagged elements (for the byte count)
    int64_t nplastic_ip = 0;
    bool mma = true;              // contractions on the FP64 tensor cores (the DFMA form was retired after the A/B)
    bool want_patch = true;       // patch form where the plan fills its element slots well enough (AMARU_EBE_PATCH=0: never)
    bool memset_y = false;        // some owned node belongs to no element: y is cleared before the patch launches
    unsigned int epoch_base = 0;  // every solve starts a fresh range of application epochs (amaru_ebe_begin)
};

struct EbeArgs {
    const int32_t *econn;
    const int32_t *einfo;
    const double *dog;
    int nmats;
    const double *geo;
    int64_t nipb;
    const double *w;
    int64_t nip_total, ip_off;
    const double *dNdR;     // [NIP][NN][ND] (Batch::d_dNdR)
    const double *Nf;       // [NIP][NN]
    const double *rho;      // [nelem] (mass term) or nullptr
    double sa, sb;
    const double *x;
    double *y;
    int mask;
    int ncol;               // non-empty element colours of the batch
    int64_t cb[EBE_MAXCOL + 1];   // their element ranges [cb[c], cb[c+1]) are consecutive
    unsigned int *bar;      // grid barrier between two colours
    double *partial;
    CgScalars *scal;
    int dot, first, last, finalize, check_done;
    int fused;              // multi-GPU fused CG loop (p2p.cuh): first launch waits for the halo of x, last launch pushes p.Ap
    P2PFused fz;
};

// econn / einfo of a batch (once per handle)
__global__ void k_ebe_econn(int64_t n, int nd, int64_t nowned, const int32_t *__restrict__ conn, const uint8_t *__restrict__ fixed,
                            int32_t *econn) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t node = conn[i];
        uint32_t v = (uint32_t)node;
        for (int d = 0; d < nd; d++) v |= (uint32_t)(fixed[(int64_t)node * nd + d] ? 1u : 0u) << (28 + d);
        if (node >= nowned) v |= 1u << 31;
        econn[i] = (int32_t)v;
    }
}
__global__ void k_ebe_einfo(int64_t n, const int32_t *__restrict__ emat, const uint8_t *__restrict__ owned, int32_t *einfo) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        einfo[i] = (einfo[i] & EI_PLASTIC) | emat[i] | ((owned && !owned[i]) ? (int)0x80000000 : 0);
}

// J⁻¹ and coef = detJ·w·th of every integration point (once per handle: the geometry does not change)
template <int NN, int ND, int NIP>
__global__ void k_ebe_geometry(int64_t nelem, const int32_t *__restrict__ conn, const double *__restrict__ coords,
                               const double *__restrict__ dNdR, const double *__restrict__ wq, double th, double *geo,
                               int64_t nipb) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nelem * NIP; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i / NIP;
        const int q = (int)(i - e * NIP);
        double X[NN * ND];
        for (int a = 0; a < NN; a++) {
            const int64_t node = conn[e * NN + a];
#pragma unroll
            for (int d = 0; d < ND; d++) X[a * ND + d] = coords[node * 3 + d];
        }
        double Ji[ND * ND];
        const double det = am_jacobian<NN, ND>(X, dNdR + q * NN * ND, Ji);
#pragma unroll
        for (int k = 0; k < ND * ND; k++) geo[(int64_t)k * nipb + i] = Ji[k];
        geo[(int64_t)(ND * ND) * nipb + i] = det * wq[q] * th;
    }
}

// w of every integration point from the current IP state (calcD of the three materials in rank-one form)
__global__ void k_ebe_tangent(int nip, int64_t nelem, int64_t ip_off, int64_t nip_total, const int32_t *__restrict__ emat,
                              const int32_t *__restrict__ mat_kind, const double *__restrict__ mat_par,
                              const double *__restrict__ state, double *w, int32_t *einfo) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nelem; e += (int64_t)gridDim.x * blockDim.x) {
        const MatPar mp = load_mat(mat_kind, mat_par, emat[e]);
        int any = 0;
        for (int q = 0; q < nip; q++) {
            const int64_t ip = ip_off + e * nip + q;
            double wv[6] = {0, 0, 0, 0, 0, 0};
            const double dlam = state[(int64_t)13 * nip_total + ip];
            if (mp.kind != AMARU_MAT_LINEAR_ELASTIC && dlam != 0.0) {
                double sig[6];
#pragma unroll
                for (int c = 0; c < 6; c++) sig[c] = state[(int64_t)c * nip_total + ip];
                if (mp.kind == AMARU_MAT_VON_MISES) {
                    if (am_J2(sig) > 0.0) {   // the assembly kernel reports the failing case (von-mises.jl:117)
                        double s[6], n[6], Dn[6];
                        am_dev(sig, s);
                        const double ns = am_norm(s);
#pragma unroll
                        for (int i = 0; i < 6; i++) n[i] = sqrt(1.5) * s[i] / ns;
                        am_De_mul(mp.E, mp.nu, n, Dn);
                        double den = 0.0;
#pragma unroll
                        for (int i = 0; i < 6; i++) den += n[i] * Dn[i];
                        den -= sqrt(1.5) * (-mp.p3);
                        const double sc = 1.0 / sqrt(den);
#pragma unroll
                        for (int i = 0; i < 6; i++) wv[i] = Dn[i] * sc;
                        any = 1;
                    }
                } else {
                    const double alpha = mp.p2, H = mp.p4;
                    double V[6];
                    if (am_J2(sig) != 0.0) {
                        double s[6];
                        am_dev(sig, s);
                        const double ns = am_norm(s);
#pragma unroll
                        for (int i = 0; i < 6; i++) V[i] = alpha * (i < 3 ? 1.0 : 0.0) + (s[i] / ns) / sqrt(2.0);
                    } else {
#pragma unroll
                        for (int i = 0; i < 6; i++) V[i] = (i < 3 ? 1.0 / sqrt(3.0) : 0.0);
                    }
                    // D = De − (De·Nu)(De·V)ᵀ/(H + V·De·Nu) with Nu = V/|V| (j2 != 0) or Nu = V (j2 == 0, |V| = 1)
                    const double nv = am_norm(V);
                    double VD[6];
                    am_De_mul(mp.E, mp.nu, V, VD);
                    double vdv = 0.0;
#pragma unroll
                    for (int i = 0; i < 6; i++) vdv += VD[i] * V[i];
                    const double den = nv * (H + vdv / nv);
                    const double sc = 1.0 / sqrt(den);
#pragma unroll
                    for (int i = 0; i < 6; i++) wv[i] = VD[i] * sc;
                    any = 1;
                }
            }
#pragma unroll
            for (int c = 0; c < 6; c++) w[(int64_t)c * nip_total + ip] = wv[c];
        }
        einfo[e] = (einfo[e] & ~EI_PLASTIC) | (any ? EI_PLASTIC : 0);
    }
}

__global__ void k_count_flags(int64_t n, int nip, const int32_t *__restrict__ f, unsigned long long *out) {
    unsigned long long c = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c += (f[i] & EI_PLASTIC) ? nip : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);   // integer count: order-independent
}

#include "ebe_mma.cuh"
#include "ebe_patch.cuh"

template <int NN, int ND, int NIP>
int ebe_mma_configure(amaru_model *m) {
    int occ0 = 0, occ1 = 0;
    const size_t s0 = MmaLayout<NN, ND, NIP, false>::bytes, s1 = MmaLayout<NN, ND, NIP, true>::bytes;
    auto k0 = k_ebe_mma<NN, ND, NIP, false>;
    auto k1 = k_ebe_mma<NN, ND, NIP, true>;
    CUDA_CHECK(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s0));
    CUDA_CHECK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, k0, 128, s0));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, k1, 128, s1));
    return m->nsm * std::max(1, std::min(std::min(occ0, occ1), 8));
}

// cooperative launch: the colour loop of k_ebe_mma synchronises the whole grid, every CTA must be resident
template <int NN, int ND, int NIP>
void ebe_mma_launch(amaru_model *m, const EbeArgs &a, int grid, bool mass) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(128);
    cfg.stream = m->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (mass) {
        cfg.dynamicSmemBytes = MmaLayout<NN, ND, NIP, true>::bytes;
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_ebe_mma<NN, ND, NIP, true>, a));
    } else {
        cfg.dynamicSmemBytes = MmaLayout<NN, ND, NIP, false>::bytes;
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_ebe_mma<NN, ND, NIP, false>, a));
    }
}

// ---- patch form: geometry / tangent planes in slot order
template <int NN, int ND, int NIP>
__global__ void k_patch_geometry(int64_t nslots, const int32_t *__restrict__ slot_elem, const int32_t *__restrict__ conn,
                                 const double *__restrict__ coords, const double *__restrict__ dNdR, const double *__restrict__ wq,
                                 double th, double *geo) {
    const int64_t nipp = nslots * NIP;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nipp; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t slot = i / NIP;
        const int q = (int)(i - slot * NIP);
        const int32_t e = slot_elem[slot];
        double Ji[ND * ND], c = 0.0;
#pragma unroll
        for (int k = 0; k < ND * ND; k++) Ji[k] = 0.0;
        if (e >= 0) {
            double X[NN * ND];
            for (int a = 0; a < NN; a++) {
                const int64_t node = conn[(int64_t)e * NN + a];
#pragma unroll
                for (int d = 0; d < ND; d++) X[a * ND + d] = coords[node * 3 + d];
            }
            c = am_jacobian<NN, ND>(X, dNdR + q * NN * ND, Ji) * wq[q] * th;
        }
#pragma unroll
        for (int k = 0; k < ND * ND; k++) geo[(int64_t)k * nipp + i] = Ji[k];
        geo[(int64_t)(ND * ND) * nipp + i] = c;
    }
}
// element records of the slots: material | empty | not-owned (plastic bit kept)
__global__ void k_patch_info(int64_t nslots, const int32_t *__restrict__ slot_elem, const int32_t *__restrict__ emat,
                             const uint8_t *__restrict__ owned, int32_t *pinfo) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nslots; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t e = slot_elem[i];
        pinfo[i] = e < 0 ? EP_EMPTY : ((pinfo[i] & EP_PLASTIC) | emat[e] | ((owned && !owned[e]) ? (int)0x80000000 : 0));
    }
}
// w planes in element order (k_ebe_tangent) -> slot order, plastic flag of the slot
__global__ void k_patch_tangent(int nip, int64_t nelem, int64_t ip_off, int64_t nip_total, const int32_t *__restrict__ elem_slot,
                                const int32_t *__restrict__ einfo, const double *__restrict__ w, int64_t nipp, double *pw, int32_t *pinfo) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nelem; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t slot = elem_slot[e];
        const bool pl = (einfo[e] & EI_PLASTIC) != 0;
        pinfo[slot] = (pinfo[slot] & ~EP_PLASTIC) | (pl ? EP_PLASTIC : 0);
        if (pl)
            for (int q = 0; q < nip; q++)
                for (int c = 0; c < 6; c++) pw[(int64_t)c * nipp + slot * nip + q] = w[(int64_t)c * nip_total + ip_off + e * nip + q];
    }
}
__global__ void k_patch_begin(unsigned int *sync, unsigned int epoch, CgScalars *scal) {
    sync[0] = 0u;
    sync[1] = epoch;
    scal->counter[0] = 0u;
}

template <int NN, int ND, int NIP, int MAXPN>
int ebe_patch_configure(amaru_model *m) {
    int occ0 = 0, occ1 = 0;
    const size_t s0 = PatchLayout<NN, ND, NIP, false, MAXPN>::bytes, s1 = PatchLayout<NN, ND, NIP, true, MAXPN>::bytes;
    auto k0 = k_ebe_patch<NN, ND, NIP, false, MAXPN>;
    auto k1 = k_ebe_patch<NN, ND, NIP, true, MAXPN>;
    CUDA_CHECK(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s0));
    CUDA_CHECK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1));
    CUDA_CHECK(cudaFuncSetAttribute(k0, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_CHECK(cudaFuncSetAttribute(k1, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, k0, 128, s0));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, k1, 128, s1));
    return m->nsm * std::max(1, std::min(occ0, occ1));
}

template <int NN, int ND, int NIP, int MAXPN>
void ebe_patch_launch(amaru_model *m, const PatchArgs &a, int grid, bool mass) {
    if (mass) k_ebe_patch<NN, ND, NIP, true, MAXPN><<<grid, 128, PatchLayout<NN, ND, NIP, true, MAXPN>::bytes, m->stream>>>(a);
    else k_ebe_patch<NN, ND, NIP, false, MAXPN><<<grid, 128, PatchLayout<NN, ND, NIP, false, MAXPN>::bytes, m->stream>>>(a);
}

// node capacities of the shared-memory bricks: must match amaru_patch_shape_params (patches.cpp)
constexpr int PN_QUAD4 = 96, PN_QUAD8 = 256, PN_HEX8 = 160, PN_HEX20 = 448, PN_TET10 = 448;

Ebe *ebe_of(amaru_model *m) { return static_cast<Ebe *>(m->ebe); }

}  // namespace


namespace {
template <class T>
T *ebe_upload(const std::vector<T> &v) {
    T *d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) CUDA_CHECK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

// plan + slot-ordered device arrays of one batch; leaves eb.patch == false when the plan fills its slots poorly
void ebe_patch_setup(amaru_model *m, Ebe *E, Batch &b, EbeBatch &eb, const std::vector<double> &h_coords, std::vector<uint8_t> &touched) {
    (void)E;
    int pe = 0, maxpn = 0, brick[3];
    amaru_patch_shape_params(b.shape, pe, maxpn, brick);
    if (maxpn == 0) return;
    std::vector<int32_t> sconn((size_t)b.nelem * b.nn);
    CUDA_CHECK(cudaMemcpy(sconn.data(), b.d_conn, sconn.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    PatchPlan P;
    std::vector<uint8_t> t2 = touched;
    amaru_build_patches(b.nn, b.nd, pe, maxpn, brick, b.nelem, sconn.data(), b.color_off, m->nnodes, m->nowned, h_coords.data(),
                        m->h_fixed.data(), t2, P);
    double minfill = 0.6;
    if (const char *e = getenv("AMARU_EBE_PATCH_MINFILL")) minfill = std::atof(e);
    if (P.fill < minfill) return;
    // Patches of one 2x2x2 neighbourhood run one after the other (8 patch colours), so an application costs at least 8 patch
    // times however many warps are idle: below ~6 patches per resident warp the colour-ordered form (one 8-element group per
    // warp step, no serial chain) is faster — the strong-scaling regime of the multi-GPU runs (measured on config 3: 2 GPUs,
    // 8 200 patches per GPU: patch form 0.31 ms, colour form 0.34 ms; 250 k elements: 0.233 vs 0.199 ms).
    int64_t minpatch = (int64_t)6 * m->nsm * 8;
    if (const char *e = getenv("AMARU_EBE_PATCH_MINPATCH")) minpatch = std::atoll(e);
    if (P.npatch < minpatch) return;
    touched.swap(t2);
    eb.patch = true;
    eb.npatch = P.npatch;
    eb.nslots = P.nslots;
    eb.fill = P.fill;
    eb.pnode_total = (int64_t)P.pnodes.size();
    for (uint32_t v : P.pnodes) eb.pnode_loaded += !(v & (PN_FIRST | PN_GHOST));
    eb.h_elem_slot.assign((size_t)b.nelem, -1);
    for (int64_t s = 0; s < P.nslots; s++)
        if (P.slot_elem[(size_t)s] >= 0) eb.h_elem_slot[(size_t)P.slot_elem[(size_t)s]] = (int32_t)s;
    AMARU_REQUIRE(P.nslots < ((int64_t)1 << 31) / b.nip, AMARU_ERR_UNSUPPORTED, "ebe: too many element slots in one batch");
    eb.d_desc = ebe_upload(P.desc);
    eb.d_deps = ebe_upload(P.deps);
    eb.d_pnodes = ebe_upload(P.pnodes);
    eb.d_slot_elem = ebe_upload(P.slot_elem);
    eb.d_elem_slot = ebe_upload(eb.h_elem_slot);
    {
        static_assert(sizeof(unsigned long long) == 4 * sizeof(uint16_t), "packed ids");
        unsigned long long *d = nullptr;
        CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(P.lane_ids.size(), 4) * sizeof(uint16_t)));
        CUDA_CHECK(cudaMemcpy(d, P.lane_ids.data(), P.lane_ids.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        eb.d_lane_ids = d;
    }
    const int64_t nipp = P.nslots * b.nip;
    const int np = b.nd * b.nd + 1;
    CUDA_CHECK(cudaMalloc(&eb.d_pgeo, (size_t)np * nipp * sizeof(double)));
    CUDA_CHECK(cudaMalloc(&eb.d_pw, (size_t)6 * nipp * sizeof(double)));
    CUDA_CHECK(cudaMemsetAsync(eb.d_pw, 0, (size_t)6 * nipp * sizeof(double), m->stream));
    CUDA_CHECK(cudaMalloc(&eb.d_pinfo, (size_t)P.nslots * sizeof(int32_t)));
    CUDA_CHECK(cudaMemsetAsync(eb.d_pinfo, 0, (size_t)P.nslots * sizeof(int32_t), m->stream));
    CUDA_CHECK(cudaMalloc(&eb.d_sync, ((size_t)P.npatch + 2) * sizeof(unsigned int)));
    CUDA_CHECK(cudaMemsetAsync(eb.d_sync, 0, ((size_t)P.npatch + 2) * sizeof(unsigned int), m->stream));
    CUDA_CHECK(cudaMalloc(&eb.d_epatch, (size_t)P.npatch * sizeof(double)));
    const int g = (int)std::min<int64_t>((nipp + 127) / 128, (int64_t)m->nsm * 16);
    const int gs = (int)std::min<int64_t>((P.nslots + 255) / 256, (int64_t)m->nsm * 16);
    k_patch_info<<<gs, 256, 0, m->stream>>>(P.nslots, eb.d_slot_elem, b.d_emat, nullptr, eb.d_pinfo);
#define PGEO(NN, ND, NIP, MAXPN)                                                                                                           \
    k_patch_geometry<NN, ND, NIP><<<g, 128, 0, m->stream>>>(P.nslots, eb.d_slot_elem, b.d_conn, m->d_coords, b.d_dNdR, b.d_w, m->th, eb.d_pgeo); \
    eb.grid_patch = ebe_patch_configure<NN, ND, NIP, MAXPN>(m)
    switch (b.shape) {
    case AMARU_SHAPE_QUAD4: PGEO(4, 2, 4, PN_QUAD4); break;
    case AMARU_SHAPE_QUAD8: PGEO(8, 2, 4, PN_QUAD8); break;
    case AMARU_SHAPE_HEX8: PGEO(8, 3, 8, PN_HEX8); break;
    case AMARU_SHAPE_HEX20: PGEO(20, 3, 8, PN_HEX20); break;
    case AMARU_SHAPE_TET10: PGEO(10, 3, 4, PN_TET10); break;
    default: throw AmaruError{AMARU_ERR_UNSUPPORTED, "ebe: unsupported shape"};
    }
#undef PGEO
    m->launches += 2;
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(m->stream));   // the plan's host vectors go out of scope
}
}  // namespace

// geometry planes + tangent planes; called from create_impl after the batches are on the device
void amaru_ebe_setup(amaru_model *m) {
    const char *op = getenv("AMARU_OPERATOR");
    m->op_ebe = !(op && std::strcmp(op, "csr") == 0);
    // Tiny systems (BASELINE configs[0]: 200 QUAD8 elements) are pure latency: the colour-ordered kernel walks through its
    // colours one after the other (~4 us each with a handful of CTAs) while the block-CSR SpMV streams a few hundred KB in one
    // step; unless the environment says otherwise such handles start with the CSR operator (amaru_set_operator still switches).
    if (m->stressmodel == AMARU_STRESS_AXISYMMETRIC) m->op_ebe = false;   // no hoop term in the matrix-free operator
    if (!op && m->nelem_total <= 1024 && 5e-6 + (double)m->nblk * m->nd * m->nd * 8.0 / 6.0e12 < 4e-6 * m->ncolors) m->op_ebe = false;
    Ebe *E = new Ebe();
    m->ebe = E;
    if (const char *e = getenv("AMARU_EBE_PATCH")) E->want_patch = std::atoi(e) != 0;
    E->b.resize(m->batches.size());
    std::vector<uint8_t> touched((size_t)m->nnodes, 0);   // first-touch bookkeeping across the batches (launch order)
    std::vector<double> h_coords;
    if (E->mma && E->want_patch) {
        h_coords.resize((size_t)m->nnodes * 3);
        CUDA_CHECK(cudaMemcpy(h_coords.data(), m->d_coords, h_coords.size() * sizeof(double), cudaMemcpyDeviceToHost));
    }
    std::vector<double> dog((size_t)m->nmats * 3);
    {
        std::vector<double> par((size_t)m->nmats * AMARU_MAT_NPARAMS);
        CUDA_CHECK(cudaMemcpy(par.data(), m->d_mat_par, par.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < m->nmats; i++) {
            const double Em = par[(size_t)i * AMARU_MAT_NPARAMS], nu = par[(size_t)i * AMARU_MAT_NPARAMS + 1];
            const double c = Em / ((1.0 + nu) * (1.0 - 2.0 * nu));
            dog[3 * i] = c * (1.0 - nu);
            dog[3 * i + 1] = c * nu;
            dog[3 * i + 2] = c * (1.0 - 2.0 * nu);
        }
    }
    CUDA_CHECK(cudaMalloc(&E->d_dog, dog.size() * sizeof(double)));
    CUDA_CHECK(cudaMemcpy(E->d_dog, dog.data(), dog.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&E->d_w, std::max<size_t>((size_t)6 * m->nip_total, 1) * sizeof(double)));
    CUDA_CHECK(cudaMemsetAsync(E->d_w, 0, std::max<size_t>((size_t)6 * m->nip_total, 1) * sizeof(double), m->stream));
    CUDA_CHECK(cudaMalloc(&E->d_nplastic, sizeof(int64_t)));
    for (size_t i = 0; i < m->batches.size(); i++) {
        Batch &b = m->batches[i];
        EbeBatch &eb = E->b[i];
        const int64_t nipb = b.nelem * b.nip;
        const int np = b.nd * b.nd + 1;
        AMARU_REQUIRE(m->nnodes < (int64_t)EC_NODE && m->nmats < EP_MAT, AMARU_ERR_UNSUPPORTED, "ebe: more than 2^28 nodes per GPU");
        if (E->mma && E->want_patch && b.nelem > 0) ebe_patch_setup(m, E, b, eb, h_coords, touched);
        if (eb.patch) {   // the patch form keeps its own slot-ordered planes; only the element records are shared
            CUDA_CHECK(cudaMalloc(&eb.d_einfo, (size_t)b.nelem * sizeof(int32_t)));
            CUDA_CHECK(cudaMemsetAsync(eb.d_einfo, 0, (size_t)b.nelem * sizeof(int32_t), m->stream));
            continue;
        }
        CUDA_CHECK(cudaMalloc(&eb.d_geo, std::max<size_t>((size_t)np * nipb, 1) * sizeof(double)));
        CUDA_CHECK(cudaMalloc(&eb.d_bar, 2 * sizeof(unsigned int)));
        CUDA_CHECK(cudaMemsetAsync(eb.d_bar, 0, 2 * sizeof(unsigned int), m->stream));
        CUDA_CHECK(cudaMalloc(&eb.d_econn, std::max<size_t>((size_t)b.nelem * b.nn, 1) * sizeof(int32_t)));
        CUDA_CHECK(cudaMalloc(&eb.d_einfo, std::max<size_t>((size_t)b.nelem, 1) * sizeof(int32_t)));
        CUDA_CHECK(cudaMemsetAsync(eb.d_einfo, 0, std::max<size_t>((size_t)b.nelem, 1) * sizeof(int32_t), m->stream));
        if (b.nelem > 0) {
            const int ge = (int)std::min<int64_t>((b.nelem * b.nn + 255) / 256, (int64_t)m->nsm * 16);
            k_ebe_econn<<<ge, 256, 0, m->stream>>>(b.nelem * b.nn, b.nd, m->nowned, b.d_conn, m->d_fixed, eb.d_econn);
            k_ebe_einfo<<<ge, 256, 0, m->stream>>>(b.nelem, b.d_emat, nullptr, eb.d_einfo);
            m->launches += 2;
        }
        if (nipb == 0) continue;
        const int g = (int)std::min<int64_t>((nipb + 127) / 128, (int64_t)m->nsm * 16);
#define GEO(NN, ND, NIP) k_ebe_geometry<NN, ND, NIP><<<g, 128, 0, m->stream>>>(b.nelem, b.d_conn, m->d_coords, b.d_dNdR, b.d_w, m->th, eb.d_geo, nipb)
        switch (b.shape) {
        case AMARU_SHAPE_QUAD4: GEO(4, 2, 4); eb.grid_mma = ebe_mma_configure<4, 2, 4>(m); break;
        case AMARU_SHAPE_QUAD8: GEO(8, 2, 4); eb.grid_mma = ebe_mma_configure<8, 2, 4>(m); break;
        case AMARU_SHAPE_HEX8: GEO(8, 3, 8); eb.grid_mma = ebe_mma_configure<8, 3, 8>(m); break;
        case AMARU_SHAPE_HEX20: GEO(20, 3, 8); eb.grid_mma = ebe_mma_configure<20, 3, 8>(m); break;
        case AMARU_SHAPE_TET10: GEO(10, 3, 4); eb.grid_mma = ebe_mma_configure<10, 3, 4>(m); break;
        default: throw AmaruError{AMARU_ERR_UNSUPPORTED, "ebe: unsupported shape"};
        }
#undef GEO
        m->launches++;
        CUDA_CHECK(cudaGetLastError());
    }
    bool all_patch = true;
    for (size_t i = 0; i < E->b.size(); i++) all_patch = all_patch && (E->b[i].patch || m->batches[i].nelem == 0);
    if (all_patch)
        for (int64_t n = 0; n < m->nowned && !E->memset_y; n++) E->memset_y = !touched[(size_t)n];
    else   // some batch runs the colour-ordered kernel (read-modify-write of y from the start): y is cleared first and the
        E->memset_y = true;   // patch launches load every row (first-touch flags dropped below)
    if (!all_patch)
        for (EbeBatch &eb : E->b)
            if (eb.patch && eb.pnode_total > 0) {
                std::vector<uint32_t> pn((size_t)eb.pnode_total);
                CUDA_CHECK(cudaMemcpy(pn.data(), eb.d_pnodes, pn.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
                eb.pnode_loaded = 0;
                for (uint32_t &v : pn) {
                    v &= ~PN_FIRST;
                    eb.pnode_loaded += !(v & PN_GHOST);
                }
                CUDA_CHECK(cudaMemcpy(eb.d_pnodes, pn.data(), pn.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            }
}

void amaru_ebe_destroy(amaru_model *m) {
    Ebe *E = ebe_of(m);
    if (!E) return;
    for (EbeBatch &eb : E->b) {
        cudaFree(eb.d_geo);
        cudaFree(eb.d_econn);
        cudaFree(eb.d_einfo);
        cudaFree(eb.d_bar);
        for (void *q : {(void *)eb.d_desc, (void *)eb.d_deps, (void *)eb.d_pinfo, (void *)eb.d_slot_elem, (void *)eb.d_elem_slot,
                        (void *)eb.d_pnodes, (void *)eb.d_lane_ids, (void *)eb.d_pgeo, (void *)eb.d_pw, (void *)eb.d_sync,
                        (void *)eb.d_epatch})
            cudaFree(q);
    }
    cudaFree(E->d_w);
    cudaFree(E->d_dog);
    cudaFree(E->d_nplastic);
    delete E;
    m->ebe = nullptr;
}

// element ownership flags of a partitioned handle (abi.cu derives them from the halo lists): batch-local, colour-sorted
void amaru_ebe_set_owned(amaru_model *m, int batch, const uint8_t *h_owned_sorted) {
    Ebe *E = ebe_of(m);
    Batch &b = m->batches[(size_t)batch];
    EbeBatch &eb = E->b[(size_t)batch];
    if (b.nelem == 0) return;
    uint8_t *d_own = nullptr;
    CUDA_CHECK(cudaMalloc(&d_own, (size_t)b.nelem));
    CUDA_CHECK(cudaMemcpy(d_own, h_owned_sorted, (size_t)b.nelem, cudaMemcpyHostToDevice));
    const int ge = (int)std::min<int64_t>((b.nelem + 255) / 256, (int64_t)m->nsm * 16);
    if (eb.patch) {
        const int gs = (int)std::min<int64_t>((eb.nslots + 255) / 256, (int64_t)m->nsm * 16);
        k_patch_info<<<gs, 256, 0, m->stream>>>(eb.nslots, eb.d_slot_elem, b.d_emat, d_own, eb.d_pinfo);
    } else {
        k_ebe_einfo<<<ge, 256, 0, m->stream>>>(b.nelem, b.d_emat, d_own, eb.d_einfo);
    }
    m->launches++;
    CUDA_CHECK(cudaStreamSynchronize(m->stream));
    cudaFree(d_own);
}

// w planes from the current IP state (called by amaru_assemble_K: the operator then equals the assembled tangent)
void amaru_ebe_refresh(amaru_model *m) {
    Ebe *E = ebe_of(m);
    if (!E) return;
    CUDA_CHECK(cudaMemsetAsync(E->d_nplastic, 0, sizeof(int64_t), m->stream));
    for (size_t i = 0; i < m->batches.size(); i++) {
        Batch &b = m->batches[i];
        if (b.nelem == 0) continue;
        const int g = (int)std::min<int64_t>((b.nelem + 127) / 128, (int64_t)m->nsm * 16);
        k_ebe_tangent<<<g, 128, 0, m->stream>>>(b.nip, b.nelem, b.ip_off, m->nip_total, b.d_emat, m->d_mat_kind, m->d_mat_par,
                                                m->d_state, E->d_w, E->b[i].d_einfo);
        k_count_flags<<<g, 128, 0, m->stream>>>(b.nelem, b.nip, E->b[i].d_einfo, reinterpret_cast<unsigned long long *>(E->d_nplastic));
        m->launches += 2;
        if (E->b[i].patch) {
            EbeBatch &eb = E->b[i];
            k_patch_tangent<<<g, 128, 0, m->stream>>>(b.nip, b.nelem, b.ip_off, m->nip_total, eb.d_elem_slot, eb.d_einfo, E->d_w,
                                                      eb.nslots * b.nip, eb.d_pw, eb.d_pinfo);
            m->launches++;
        }
    }
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(&E->nplastic_ip, E->d_nplastic, sizeof(int64_t), cudaMemcpyDeviceToHost, m->stream));
}

// y = (sysA·K + sysB·M) x on the owned rows (+ p·Ap and CG scalar finalisation when dot != 0); x needs valid ghost entries
void amaru_ebe_apply(amaru_model *m, const double *x, double *y, int mask, int dot, int check_done, int finalize, int fused) {
    Ebe *E = ebe_of(m);
    AMARU_REQUIRE(E != nullptr, AMARU_ERR_ARG, "ebe: operator not set up");
    const bool mass = m->sysB != 0.0;
    if (E->memset_y) CUDA_CHECK(cudaMemsetAsync(y, 0, (size_t)m->nowned * m->nd * sizeof(double), m->stream));
    // number of non-empty launches (one per patch batch, one per element colour otherwise), to flag the last one
    int nl = 0, il = 0;
    for (size_t i = 0; i < m->batches.size(); i++) {
        Batch &b = m->batches[i];
        if (E->b[i].patch) {
            nl++;
            continue;
        }
        int ncol = 0;
        for (size_t c = 0; c + 1 < b.color_off.size(); c++) ncol += b.color_off[c + 1] > b.color_off[c];
        nl += (ncol + EBE_MAXCOL - 1) / EBE_MAXCOL;
    }
    EbeArgs a;
    a.dog = E->d_dog; a.w = E->d_w; a.nip_total = m->nip_total; a.sa = m->sysA; a.sb = m->sysB;
    a.x = x; a.y = y; a.mask = mask; a.nmats = m->nmats;
    a.partial = m->d_partial; a.scal = m->d_scal; a.dot = dot; a.finalize = finalize; a.check_done = check_done;
    a.first = 1;
    a.fused = fused;
    if (a.fused) a.fz = amaru_comm_fused_args(m);
    else std::memset(&a.fz, 0, sizeof(a.fz));
    for (size_t i = 0; i < m->batches.size(); i++) {
        Batch &b = m->batches[i];
        EbeBatch &eb = E->b[i];
        AMARU_REQUIRE(!mass || b.d_rho != nullptr, AMARU_ERR_ARG, "ebe: mass term without densities (call amaru_assemble_M)");
        if (eb.patch) {
            PatchArgs pa;
            pa.desc = eb.d_desc; pa.pnodes = eb.d_pnodes; pa.lane_ids = eb.d_lane_ids; pa.deps = eb.d_deps;
            pa.einfo = eb.d_pinfo; pa.slot_elem = eb.d_slot_elem; pa.geo = eb.d_pgeo; pa.w = eb.d_pw; pa.nipp = eb.nslots * b.nip;
            pa.dog = E->d_dog; pa.nmats = m->nmats; pa.dNdR = b.d_dNdR; pa.Nf = b.d_N; pa.rho = b.d_rho;
            pa.sa = m->sysA; pa.sb = m->sysB; pa.x = x; pa.y = y; pa.mask = mask; pa.npatch = eb.npatch;
            pa.ticket = eb.d_sync; pa.epoch = eb.d_sync + 1; pa.done = eb.d_sync + 2; pa.counter = &m->d_scal->counter[0];
            pa.epatch = eb.d_epatch; pa.scal = m->d_scal; pa.dot = dot; pa.finalize = finalize; pa.check_done = check_done;
            pa.fused = fused;
            if (fused) pa.fz = amaru_comm_fused_args(m);
            else std::memset(&pa.fz, 0, sizeof(pa.fz));
            il++;
            pa.first = a.first;
            pa.last = il == nl;
            const int grid = (int)std::min<int64_t>(((int64_t)eb.npatch + 3) / 4, eb.grid_patch);
            switch (b.shape) {
            case AMARU_SHAPE_QUAD4: ebe_patch_launch<4, 2, 4, PN_QUAD4>(m, pa, grid, mass); break;
            case AMARU_SHAPE_QUAD8: ebe_patch_launch<8, 2, 4, PN_QUAD8>(m, pa, grid, mass); break;
            case AMARU_SHAPE_HEX8: ebe_patch_launch<8, 3, 8, PN_HEX8>(m, pa, grid, mass); break;
            case AMARU_SHAPE_HEX20: ebe_patch_launch<20, 3, 8, PN_HEX20>(m, pa, grid, mass); break;
            case AMARU_SHAPE_TET10: ebe_patch_launch<10, 3, 4, PN_TET10>(m, pa, grid, mass); break;
            default: throw AmaruError{AMARU_ERR_UNSUPPORTED, "ebe: unsupported shape"};
            }
            m->launches++;
            a.first = 0;
            continue;
        }
        a.econn = eb.d_econn; a.einfo = eb.d_einfo; a.geo = eb.d_geo; a.bar = eb.d_bar;
        a.nipb = b.nelem * b.nip; a.ip_off = b.ip_off; a.dNdR = b.d_dNdR; a.Nf = b.d_N; a.rho = b.d_rho;
        // the batch's non-empty colours, EBE_MAXCOL per persistent launch (one launch for every mesh seen so far)
        size_t c = 0;
        const size_t nc = b.color_off.size() - 1;
        while (c < nc) {
            a.ncol = 0;
            int64_t biggest = 0;
            while (c < nc && a.ncol < EBE_MAXCOL) {
                const int64_t n = b.color_off[c + 1] - b.color_off[c];
                if (n > 0) {
                    a.cb[a.ncol] = b.color_off[c];       // colour ranges are consecutive: cb[k + 1] of one is cb[k] of the next
                    a.cb[a.ncol + 1] = b.color_off[c + 1];
                    a.ncol++;
                    biggest = std::max(biggest, n);
                }
                c++;
            }
            if (a.ncol == 0) break;
            il++;
            a.last = il == nl;
            const int grid = (int)std::min<int64_t>((biggest + 31) / 32, eb.grid_mma);   // 4 warps per CTA, 8 elements per warp step
            switch (b.shape) {
            case AMARU_SHAPE_QUAD4: ebe_mma_launch<4, 2, 4>(m, a, grid, mass); break;
            case AMARU_SHAPE_QUAD8: ebe_mma_launch<8, 2, 4>(m, a, grid, mass); break;
            case AMARU_SHAPE_HEX8: ebe_mma_launch<8, 3, 8>(m, a, grid, mass); break;
            case AMARU_SHAPE_HEX20: ebe_mma_launch<20, 3, 8>(m, a, grid, mass); break;
            case AMARU_SHAPE_TET10: ebe_mma_launch<10, 3, 4>(m, a, grid, mass); break;
            default: throw AmaruError{AMARU_ERR_UNSUPPORTED, "ebe: unsupported shape"};
            }
            m->launches++;
            a.first = 0;
        }
    }
    CUDA_CHECK(cudaGetLastError());
}

// the DMMA forms of the operator (patch and colour-ordered) carry the hooks of the fused multi-GPU loop
bool amaru_ebe_fusable(const amaru_model *m) {
    const Ebe *E = static_cast<const Ebe *>(m->ebe);
    return E && E->mma;
}

// Start of a solve / measurement: ticket counters cleared and a fresh range of application epochs, so that a kernel that gave
// up waiting in an earlier solve (reported as a breakdown) cannot leave flags that look current
void amaru_ebe_begin(amaru_model *m) {
    Ebe *E = ebe_of(m);
    if (!E) return;
    bool any = false;
    for (EbeBatch &eb : E->b) any = any || eb.patch;
    if (!any) return;
    E->epoch_base += 1u << 20;
    for (EbeBatch &eb : E->b)
        if (eb.patch) {
            k_patch_begin<<<1, 1, 0, m->stream>>>(eb.d_sync, E->epoch_base, m->d_scal);
            m->launches++;
        }
    CUDA_CHECK(cudaGetLastError());
}

// DRAM-side statistics of the patch form: node entries of all patches and those whose y is read back
void amaru_ebe_patch_stats(const amaru_model *m, int64_t *npatch, int64_t *nslots, int64_t *pnodes, int64_t *pnodes_loaded) {
    const Ebe *E = static_cast<const Ebe *>(m->ebe);
    *npatch = *nslots = *pnodes = *pnodes_loaded = 0;
    if (!E) return;
    for (const EbeBatch &eb : E->b)
        if (eb.patch) {
            *npatch += eb.npatch;
            *nslots += eb.nslots;
            *pnodes += eb.pnode_total;
            *pnodes_loaded += eb.pnode_loaded;
        }
}

// algorithmic bytes of one application
int64_t amaru_ebe_bytes(const amaru_model *m) {
    const Ebe *E = static_cast<const Ebe *>(m->ebe);
    int64_t bytes = 0;
    for (size_t i = 0; i < m->batches.size(); i++) {
        const Batch &b = m->batches[i];
        bytes += (int64_t)8 * (b.nd * b.nd + 1) * b.nelem * b.nip;
        if (E && E->b[i].patch)   // stored format: lane-major packed 16-bit ids (one 64-bit word per lane, word and group), element
            // records, node entries of the patches (4 B each)
            bytes += (E->b[i].nslots / 8) * (int64_t)amaru_patch_id_words(b.nn) * 256 + 4 * E->b[i].nslots + 4 * E->b[i].pnode_total;
        else bytes += (int64_t)4 * b.nn * b.nelem + 5 * b.nelem;
    }
    bytes += 48 * (E ? E->nplastic_ip : 0);
    bytes += 16 * m->nowned * m->nd;
    return bytes;
}

const char *amaru_ebe_kernel(const amaru_model *m) {
    if (m->batches.empty()) return "k_ebe_mma";
    const Ebe *E = static_cast<const Ebe *>(m->ebe);
    if (E && E->mma && !E->b.empty() && E->b[0].patch) {
        switch (m->batches[0].shape) {
        case AMARU_SHAPE_QUAD4: return "k_ebe_patch<4,2,4>";
        case AMARU_SHAPE_QUAD8: return "k_ebe_patch<8,2,4>";
        case AMARU_SHAPE_HEX8: return "k_ebe_patch<8,3,8>";
        case AMARU_SHAPE_HEX20: return "k_ebe_patch<20,3,8>";
        case AMARU_SHAPE_TET10: return "k_ebe_patch<10,3,4>";
        }
    }
    if (E && E->mma) {
        switch (m->batches[0].shape) {
        case AMARU_SHAPE_QUAD4: return "k_ebe_mma<4,2,4>";
        case AMARU_SHAPE_QUAD8: return "k_ebe_mma<8,2,4>";
        case AMARU_SHAPE_HEX8: return "k_ebe_mma<8,3,8>";
        case AMARU_SHAPE_HEX20: return "k_ebe_mma<20,3,8>";
        case AMARU_SHAPE_TET10: return "k_ebe_mma<10,3,4>";
        }
    }
    return "k_ebe_mma";
}
