// K2: stress update + internal forces.  Replaces update_elem! (reference src/mech/elem/mech-solid.jl:243-279) and the
// element loop of update_state! (src/mech/mech-solver.jl:124-144) with the material update_state! functions
// (linear-elastic.jl:130-137, von-mises.jl:128-156, drucker-prager.jl:112-149).
//
// One thread per integration point: the NIP consecutive lanes of an element share its gathered coordinates and ΔU in
// shared memory, each lane rebuilds J⁻¹/∇N at its own IP, forms Δε = B·ΔU, runs the return mapping on the IP state it
// loads from / stores to the SoA planes (fully coalesced: consecutive lanes = consecutive IPs), and the element force
// coef·BᵀΔσ is reduced over the NIP lanes with xor-shuffles.  Elements of one colour share no node, so the nodal
// accumulation is a plain read-modify-write with a fixed order (colour by colour): deterministic, no atomics.
// mode 1 computes Σ coef·Bᵀσ (elem_internal_forces, mech-solid.jl:208-240) without touching the state.
//
// Also here: the permutation kernels between the ABI's IP order (element-major in the caller's element order) and the
// device order (colour-sorted), used by amaru_set_state / amaru_get_state (K9 of SURVEY §2.2 is a plain D2D copy).
#include "materials.cuh"

namespace {

struct UpdArgs {
    const double *coords;
    const int32_t *conn;
    const int32_t *emat;
    const uint8_t *owned;
    const int32_t *mat_kind;
    const double *mat_par;
    const double *dNdR;
    const double *w;
    double *state;
    int64_t nip_total, ip_off;
    double th;
    const double *dU;  // node-major
    double *f;         // node-major
    int *status;
    int64_t e_begin, e_end;
    int mode;          // 0 update_state!, 1 internal forces of the current stress
    int axi;           // stressmodel = :axisymmetric (2D): hoop row of B, th = 2*pi*r (mech-solid.jl:94-108, 260-262)
    const double *Nf;  // [NIP][NN] shape functions at the integration points (axisymmetric only)
};

template <int NN, int ND, int NIP, int NT>
__global__ void __launch_bounds__(NT) k_update(UpdArgs p) {
    constexpr int EPB = NT / NIP;
    __shared__ double sdN[NIP * NN * ND];
    __shared__ double sw[NIP];
    __shared__ double sX[EPB * NN * ND];
    __shared__ double sU[EPB * NN * ND];
    __shared__ int32_t sNode[EPB * NN];
    const int tid = threadIdx.x;
    const int64_t e0 = p.e_begin + (int64_t)blockIdx.x * EPB;
    const int ne = (int)min((int64_t)EPB, p.e_end - e0);
    for (int i = tid; i < NIP * NN * ND; i += NT) sdN[i] = p.dNdR[i];
    if (tid < NIP) sw[tid] = p.w[tid];
    for (int i = tid; i < ne * NN; i += NT) {
        const int32_t node = p.conn[e0 * NN + i];
        sNode[i] = node;
#pragma unroll
        for (int d = 0; d < ND; d++) {
            sX[i * ND + d] = p.coords[(int64_t)node * 3 + d];
            sU[i * ND + d] = p.mode == 0 ? p.dU[(int64_t)node * ND + d] : 0.0;
        }
    }
    __syncthreads();
    const int e = tid / NIP, q = tid - e * NIP;
    const bool active = e < ne;          // whole NIP-lane groups are active or not, so the shuffles below are safe
    double Ji[ND * ND], coef = 0.0, ds[6], rad = 1.0;
#pragma unroll
    for (int c = 0; c < 6; c++) ds[c] = 0.0;
    if (active) {
        const double *X = sX + e * NN * ND, *dN = sdN + q * NN * ND, *U = sU + e * NN * ND;
        const double det = am_jacobian<NN, ND>(X, dN, Ji);
        // no detJ > 0 test here: the reference's update_elem! has none (mech-solid.jl:243-279; only elem_stiffness :150 and
        // elem_mass :194 raise), and with fixed coordinates the assembly that precedes every update has already tested it
        coef = det * sw[q] * p.th;
        if constexpr (ND == 2) {
            if (p.axi) {
                rad = 0.0;
                for (int a = 0; a < NN; a++) rad += p.Nf[q * NN + a] * X[a * 2];   // ip.coord.x
                coef = det * sw[q] * 2.0 * 3.14159265358979323846 * rad;
            }
        }
        const int64_t ip = p.ip_off + (e0 + e) * NIP + q;
        if (p.mode == 0) {
            double de[6];
#pragma unroll
            for (int c = 0; c < 6; c++) de[c] = 0.0;
            for (int a = 0; a < NN; a++) {
                double g[ND];
#pragma unroll
                for (int j = 0; j < ND; j++) {
                    double v = 0.0;
#pragma unroll
                    for (int k = 0; k < ND; k++) v += dN[a * ND + k] * Ji[k * ND + j];
                    g[j] = v;
                }
                if constexpr (ND == 3) {
                    const double ux = U[a * 3], uy = U[a * 3 + 1], uz = U[a * 3 + 2];
                    const double hx = g[0] / AM_SR2, hy = g[1] / AM_SR2, hz = g[2] / AM_SR2;
                    de[0] += g[0] * ux;
                    de[1] += g[1] * uy;
                    de[2] += g[2] * uz;
                    de[3] += hz * uy + hy * uz;
                    de[4] += hz * ux + hx * uz;
                    de[5] += hy * ux + hx * uy;
                } else {
                    const double ux = U[a * 2], uy = U[a * 2 + 1];
                    de[0] += g[0] * ux;
                    de[1] += g[1] * uy;
                    if (p.axi) de[2] += (p.Nf[q * NN + a] / rad) * ux;
                    de[5] += (g[1] / AM_SR2) * ux + (g[0] / AM_SR2) * uy;
                }
            }
            double sig[6], eps[6];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                sig[c] = p.state[(int64_t)c * p.nip_total + ip];
                eps[c] = p.state[(int64_t)(6 + c) * p.nip_total + ip];
            }
            double epa = p.state[(int64_t)12 * p.nip_total + ip];
            double dlam = p.state[(int64_t)13 * p.nip_total + ip];
            const MatPar mp = load_mat(p.mat_kind, p.mat_par, p.emat[e0 + e]);
            const int st = am_update(mp, sig, eps, epa, dlam, de, ds);
            if (st) atomicMax(p.status, st);
#pragma unroll
            for (int c = 0; c < 6; c++) {
                p.state[(int64_t)c * p.nip_total + ip] = sig[c];
                p.state[(int64_t)(6 + c) * p.nip_total + ip] = eps[c];
            }
            p.state[(int64_t)12 * p.nip_total + ip] = epa;
            p.state[(int64_t)13 * p.nip_total + ip] = dlam;
        } else {
#pragma unroll
            for (int c = 0; c < 6; c++) ds[c] = p.state[(int64_t)c * p.nip_total + ip];
        }
    }
    // element force: reduce coef·B_aᵀ·Δσ over the NIP lanes of the element, lane q stores nodes a ≡ q (mod NIP)
    const double *dN = sdN + q * NN * ND;
    for (int a = 0; a < NN; a++) {
        double f[ND];
        {
            double g[ND];
#pragma unroll
            for (int j = 0; j < ND; j++) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < ND; k++) v += dN[a * ND + k] * (active ? Ji[k * ND + j] : 0.0);
                g[j] = v;
            }
            if constexpr (ND == 3) {
                const double hx = g[0] / AM_SR2, hy = g[1] / AM_SR2, hz = g[2] / AM_SR2;
                f[0] = coef * (g[0] * ds[0] + hz * ds[4] + hy * ds[5]);
                f[1] = coef * (g[1] * ds[1] + hz * ds[3] + hx * ds[5]);
                f[2] = coef * (g[2] * ds[2] + hy * ds[3] + hx * ds[4]);
            } else {
                f[0] = coef * (g[0] * ds[0] + (g[1] / AM_SR2) * ds[5]);
                if (p.axi && active) f[0] += coef * (p.Nf[q * NN + a] / rad) * ds[2];
                f[1] = coef * (g[1] * ds[1] + (g[0] / AM_SR2) * ds[5]);
            }
        }
#pragma unroll
        for (int d = 0; d < ND; d++) {
#pragma unroll
            for (int o = 1; o < NIP; o <<= 1) f[d] += __shfl_xor_sync(0xffffffffu, f[d], o);
        }
        if (active && (a % NIP) == q) {
            // halo elements (multi-GPU) also add to the rows this rank owns; ghost rows are never read
            const int64_t node = sNode[e * NN + a];
#pragma unroll
            for (int d = 0; d < ND; d++) p.f[node * ND + d] += f[d];
        }
    }
}

template <int NN, int ND, int NIP, int NT>
void launch_upd(amaru_model *m, Batch &b, const double *dU, double *f, int mode) {
    UpdArgs a;
    a.coords = m->d_coords; a.conn = b.d_conn; a.emat = b.d_emat; a.owned = b.d_owned;
    a.mat_kind = m->d_mat_kind; a.mat_par = m->d_mat_par; a.dNdR = b.d_dNdR; a.w = b.d_w;
    a.state = m->d_state; a.nip_total = m->nip_total; a.ip_off = b.ip_off; a.th = m->th;
    a.dU = dU; a.f = f; a.status = m->d_status; a.mode = mode;
    a.axi = m->stressmodel == AMARU_STRESS_AXISYMMETRIC; a.Nf = b.d_N;
    constexpr int EPB = NT / NIP;
    for (size_t c = 0; c + 1 < b.color_off.size(); c++) {
        a.e_begin = b.color_off[c];
        a.e_end = b.color_off[c + 1];
        const int64_t n = a.e_end - a.e_begin;
        if (n <= 0) continue;
        k_update<NN, ND, NIP, NT><<<(unsigned)((n + EPB - 1) / EPB), NT, 0, m->stream>>>(a);
        m->launches++;
    }
    CUDA_CHECK(cudaGetLastError());
}

// ABI order <-> device order of the IP state.  io: [nip*ncomp] (ABI, IP-major), planes: [ncomp][nip_total]
__global__ void k_state_permute(int nip, int64_t nelem, int64_t elem_off, int64_t ip_off, const int64_t *perm,
                                double *io, double *state, int64_t nip_total, int plane0, int ncomp, int to_device) {
    const int64_t total = nelem * nip;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t es = i / nip;
        const int q = (int)(i - es * nip);
        const int64_t src_ip = ip_off + perm[es] * nip + q;       // ABI position: batch after batch (batches may differ in nip)
        const int64_t dev_ip = ip_off + i;
        for (int c = 0; c < ncomp; c++) {
            if (to_device)
                state[(int64_t)(plane0 + c) * nip_total + dev_ip] = io[src_ip * ncomp + c];
            else
                io[src_ip * ncomp + c] = state[(int64_t)(plane0 + c) * nip_total + dev_ip];
        }
    }
}

}  // namespace

void amaru_launch_update(amaru_model *m, const double *d_dU_nodes, double *d_f_nodes, int mode) {
    CUDA_CHECK(cudaMemsetAsync(d_f_nodes, 0, (size_t)m->nnodes * m->nd * sizeof(double), m->stream));
    for (Batch &b : m->batches) {
        switch (b.shape) {
        case AMARU_SHAPE_QUAD4: launch_upd<4, 2, 4, 128>(m, b, d_dU_nodes, d_f_nodes, mode); break;
        case AMARU_SHAPE_QUAD8: launch_upd<8, 2, 4, 128>(m, b, d_dU_nodes, d_f_nodes, mode); break;
        case AMARU_SHAPE_HEX8: launch_upd<8, 3, 8, 128>(m, b, d_dU_nodes, d_f_nodes, mode); break;
        case AMARU_SHAPE_HEX20: launch_upd<20, 3, 8, 128>(m, b, d_dU_nodes, d_f_nodes, mode); break;
        case AMARU_SHAPE_TET10: launch_upd<10, 3, 4, 128>(m, b, d_dU_nodes, d_f_nodes, mode); break;
        default: throw AmaruError{AMARU_ERR_UNSUPPORTED, "update: unsupported shape"};
        }
    }
    // plane stress (LinearElastic, linear-elastic.jl:99-108: the zz row of De is zero): the kernels run on the equivalent
    // plane-strain constants (abi.cu), whose σzz = c*ν*(εxx+εyy) is the one entry that differs — cleared here
    if (mode == 0 && m->stressmodel == AMARU_STRESS_PLANESTRESS && m->nip_total > 0)
        CUDA_CHECK(cudaMemsetAsync(m->d_state + 2 * m->nip_total, 0, (size_t)m->nip_total * sizeof(double), m->stream));
}

void amaru_state_permute(amaru_model *m, double *d_io, int plane0, int ncomp, bool to_device) {
    for (Batch &b : m->batches) {
        const int64_t total = b.nelem * b.nip;
        if (total == 0) continue;
        const int64_t blocks = std::min<int64_t>((total + 255) / 256, (int64_t)m->nsm * 32);
        k_state_permute<<<(unsigned)blocks, 256, 0, m->stream>>>(b.nip, b.nelem, b.elem_off, b.ip_off, b.d_perm, d_io,
                                                                 m->d_state, m->nip_total, plane0, ncomp, to_device ? 1 : 0);
        m->launches++;
    }
    CUDA_CHECK(cudaGetLastError());
}
