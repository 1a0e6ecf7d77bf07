// K1: tangent stiffness assembly  K = Σ_e Σ_ip (detJ·w·th)·Bᵀ·D(state)·B  straight into the block-CSR values.
//
// Replaces elem_stiffness (reference src/mech/elem/mech-solid.jl:124-166, setB :82-121) + mount_K
// (src/mech/mech-solver.jl:78-110).  Where the reference builds a dense Ke per element, pushes ne² COO triplets and
// sorts/merges them in sparse(), this kernel
//   * recomputes J, J⁻¹, detJ and ∇N at each IP from the gathered nodal coordinates (cheaper than reading a cache),
//   * evaluates the constitutive tangent D of each IP from the device-resident IP state (materials.cuh),
//   * stages T_b = D·B_b (6 x nd per node) and the coef-scaled ∇N_a in shared memory, and
//   * lets one thread own one node pair (a,b): it accumulates the nd x nd block Σ_ip coef·B_aᵀ·T_b in registers
//     (B_a has 3 non-zeros per column, so 27 FMAs per IP instead of a dense 6x60 contraction) and adds it to the
//     destination block found through the precomputed scatter map.
// Determinism without atomics: elements are processed colour by colour (elements of one colour share no node, hence
// no destination block), one launch per colour, so every block receives its contributions in a fixed order.
// K3 (consistent mass, mech-solid.jl:169-205 + dyn-solver.jl:72-103) uses the same map with N_a·N_b·I blocks.
#include <atomic>

#include "materials.cuh"

namespace {

template <int NN, int ND, int NIP, int EPB>
struct AsmSmem {
    static constexpr int NV = (ND == 3) ? 6 : 3;          // strain rows that B touches
    static constexpr int TROW = NV * ND;                  // entries of T_b
    static constexpr size_t doubles = (size_t)NIP * NN * ND   // dNdR table
                                      + NIP                   // weights
                                      + (size_t)EPB * NN * ND // X
                                      + (size_t)EPB * NIP * ND * ND  // J⁻¹
                                      + (size_t)EPB * NIP            // coef
                                      + (size_t)EPB * NIP * 36       // D
                                      + (size_t)EPB * NIP * NN * 2 * ND  // coef·g, coef·g/√2 of node a
                                      + (size_t)EPB * NIP * TROW * NN;   // T_b, component-major
};

struct AsmArgs {
    const double *coords;
    const int32_t *conn;
    const int32_t *emat;
    const int32_t *map;
    const int32_t *mat_kind;
    const double *mat_par;
    const double *dNdR;
    const double *w;
    const double *state;   // planes
    int64_t nip_total;
    int64_t ip_off;        // batch offset into the planes
    double th;
    double *K;
    int *status;
    int64_t e_begin, e_end;  // colour-sorted element range of this launch
};

template <int NN, int ND, int NIP, int EPB, int NT>
__global__ void __launch_bounds__(NT) k_assemble_K(AsmArgs p) {
    using L = AsmSmem<NN, ND, NIP, EPB>;
    constexpr int NV = L::NV, TROW = L::TROW, BS2 = ND * ND;
    extern __shared__ double smem[];
    double *sdN = smem;
    double *sw = sdN + NIP * NN * ND;
    double *sX = sw + NIP;
    double *sJi = sX + EPB * NN * ND;
    double *scoef = sJi + EPB * NIP * ND * ND;
    double *sD = scoef + EPB * NIP;
    double *sGa = sD + EPB * NIP * 36;
    double *sT = sGa + EPB * NIP * NN * 2 * ND;

    const int tid = threadIdx.x;
    const int64_t e0 = p.e_begin + (int64_t)blockIdx.x * EPB;
    const int ne = (int)min((int64_t)EPB, p.e_end - e0);

    for (int i = tid; i < NIP * NN * ND; i += NT) sdN[i] = p.dNdR[i];
    if (tid < NIP) sw[tid] = p.w[tid];
    for (int i = tid; i < ne * NN; i += NT) {
        const int64_t node = p.conn[e0 * NN + i];
#pragma unroll
        for (int d = 0; d < ND; d++) sX[i * ND + d] = p.coords[node * 3 + d];
    }
    __syncthreads();

    // one thread per (element, ip): Jacobian, coef, tangent D
    for (int i = tid; i < ne * NIP; i += NT) {
        const int e = i / NIP, q = i - e * NIP;
        double Ji[ND * ND];
        const double det = am_jacobian<NN, ND>(sX + e * NN * ND, sdN + q * NN * ND, Ji);
        if (!(det > 0.0)) atomicMax(p.status, AMARU_FAIL_NEG_JACOBIAN);   // mech-solid.jl:150
#pragma unroll
        for (int k = 0; k < ND * ND; k++) sJi[i * ND * ND + k] = Ji[k];
        scoef[i] = det * sw[q] * p.th;
        const int64_t ip = p.ip_off + (e0 + e) * NIP + q;
        double sig[6];
#pragma unroll
        for (int c = 0; c < 6; c++) sig[c] = p.state[(int64_t)c * p.nip_total + ip];
        const double dlam = p.state[(int64_t)13 * p.nip_total + ip];
        const MatPar mp = load_mat(p.mat_kind, p.mat_par, p.emat[e0 + e]);
        double D[36];
        const int st = am_calcD(mp, sig, dlam, D);
        if (st) atomicMax(p.status, st);
#pragma unroll
        for (int k = 0; k < 36; k++) sD[i * 36 + k] = D[k];
    }
    __syncthreads();

    // one thread per (element, ip, node): ∇N = dNdR·J⁻¹, T = D·B_node
    for (int i = tid; i < ne * NIP * NN; i += NT) {
        const int eq = i / NN, b = i - eq * NN;   // eq = e*NIP + q
        const int q = eq % NIP;
        const double *dn = sdN + (q * NN + b) * ND;
        const double *Ji = sJi + eq * ND * ND;
        double g[ND], gs[ND];
#pragma unroll
        for (int j = 0; j < ND; j++) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < ND; k++) v += dn[k] * Ji[k * ND + j];
            g[j] = v;
            gs[j] = v / AM_SR2;               // mech-solid.jl:115-117 divides by SR2
        }
        const double cf = scoef[eq];
        double *ga = sGa + (size_t)i * 2 * ND;
#pragma unroll
        for (int j = 0; j < ND; j++) {
            ga[j] = cf * g[j];
            ga[ND + j] = cf * gs[j];
        }
        const double *D = sD + eq * 36;
        double *T = sT + (size_t)eq * TROW * NN + b;
        if constexpr (ND == 3) {
#pragma unroll
            for (int k = 0; k < 6; k++) {
                T[(k * 3 + 0) * NN] = D[6 * k + 0] * g[0] + D[6 * k + 4] * gs[2] + D[6 * k + 5] * gs[1];
                T[(k * 3 + 1) * NN] = D[6 * k + 1] * g[1] + D[6 * k + 3] * gs[2] + D[6 * k + 5] * gs[0];
                T[(k * 3 + 2) * NN] = D[6 * k + 2] * g[2] + D[6 * k + 3] * gs[1] + D[6 * k + 4] * gs[0];
            }
        } else {
            const int rows[3] = {0, 1, 5};
#pragma unroll
            for (int kk = 0; kk < 3; kk++) {
                const int k = rows[kk];
                T[(kk * 2 + 0) * NN] = D[6 * k + 0] * g[0] + D[6 * k + 5] * gs[1];
                T[(kk * 2 + 1) * NN] = D[6 * k + 1] * g[1] + D[6 * k + 5] * gs[0];
            }
        }
    }
    __syncthreads();

    // one thread per (element, a <= b): nd x nd block K_ab; the tangent of all three materials is symmetric (associated flow),
    // so K_ba = K_abᵀ is written from the same registers — half the contractions of the round-1 kernel
    constexpr int NPAIR = NN * (NN + 1) / 2;
    for (int i = tid; i < ne * NPAIR; i += NT) {
        const int e = i / NPAIR, t = i - e * NPAIR;
        int a = 0, rem = t;                              // t -> (a, b), rows of the upper triangle
        while (rem >= NN - a) {
            rem -= NN - a;
            a++;
        }
        const int b = a + rem, ab = a * NN + b;
        double acc[BS2];
#pragma unroll
        for (int k = 0; k < BS2; k++) acc[k] = 0.0;
#pragma unroll 2
        for (int q = 0; q < NIP; q++) {
            const int eq = e * NIP + q;
            const double *ga = sGa + ((size_t)eq * NN + a) * 2 * ND;
            const double *T = sT + (size_t)eq * TROW * NN + b;
            if constexpr (ND == 3) {
                const double gx = ga[0], gy = ga[1], gz = ga[2], hx = ga[3], hy = ga[4], hz = ga[5];
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const double t0 = T[(0 * 3 + j) * NN], t1 = T[(1 * 3 + j) * NN], t2 = T[(2 * 3 + j) * NN];
                    const double t3 = T[(3 * 3 + j) * NN], t4 = T[(4 * 3 + j) * NN], t5 = T[(5 * 3 + j) * NN];
                    acc[0 * 3 + j] += gx * t0 + hz * t4 + hy * t5;
                    acc[1 * 3 + j] += gy * t1 + hz * t3 + hx * t5;
                    acc[2 * 3 + j] += gz * t2 + hy * t3 + hx * t4;
                }
            } else {
                const double gx = ga[0], gy = ga[1], hx = ga[2], hy = ga[3];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const double t0 = T[(0 * 2 + j) * NN], t1 = T[(1 * 2 + j) * NN], t5 = T[(2 * 2 + j) * NN];
                    acc[0 * 2 + j] += gx * t0 + hy * t5;
                    acc[1 * 2 + j] += gy * t1 + hx * t5;
                }
            }
        }
        const int32_t dst = p.map[(e0 + e) * (int64_t)(NN * NN) + ab];
        if (dst >= 0) {
            double *Kb = p.K + (int64_t)dst * BS2;
#pragma unroll
            for (int k = 0; k < BS2; k++) Kb[k] += acc[k];
        }
        if (a != b) {
            const int32_t dst2 = p.map[(e0 + e) * (int64_t)(NN * NN) + b * NN + a];
            if (dst2 >= 0) {
                double *Kb = p.K + (int64_t)dst2 * BS2;
#pragma unroll
                for (int r = 0; r < ND; r++)
#pragma unroll
                    for (int c = 0; c < ND; c++) Kb[r * ND + c] += acc[c * ND + r];
            }
        }
    }
}


// Axisymmetric tangent (stressmodel = :axisymmetric, 2D cells): B gets the hoop row ε_θθ = N_a/r (setB, mech-solid.jl:94-108)
// and th = 2π·r of the integration point (:143).  One CTA per element, one thread per node pair, the four strain rows
// (rr, zz, θθ, rz) written out; such models are small, so this kernel is plain rather than tuned.  Same colour-ordered
// read-modify-write as k_assemble_K (no atomics, fixed order).
template <int NN, int NIP>
__global__ void __launch_bounds__(64) k_assemble_K_axi(AsmArgs p, const double *__restrict__ Nf) {
    __shared__ double sX[NN * 2], sG[NIP][NN][2], sNr[NIP][NN], sD[NIP][16], sc[NIP];
    const int tid = threadIdx.x;
    const int64_t e = p.e_begin + blockIdx.x;
    for (int i = tid; i < NN; i += 64) {
        const int64_t node = p.conn[e * NN + i];
        sX[i * 2] = p.coords[node * 3];
        sX[i * 2 + 1] = p.coords[node * 3 + 1];
    }
    __syncthreads();
    if (tid < NIP) {
        const int q = tid;
        double Ji[4];
        const double det = am_jacobian<NN, 2>(sX, p.dNdR + q * NN * 2, Ji);
        if (!(det > 0.0)) atomicMax(p.status, AMARU_FAIL_NEG_JACOBIAN);   // mech-solid.jl:150
        double r = 0.0;
        for (int a = 0; a < NN; a++) r += Nf[q * NN + a] * sX[a * 2];      // ip.coord.x
        sc[q] = det * p.w[q] * 2.0 * 3.14159265358979323846 * r;
        for (int a = 0; a < NN; a++) {
            const double *dn = p.dNdR + (q * NN + a) * 2;
            sG[q][a][0] = dn[0] * Ji[0] + dn[1] * Ji[2];
            sG[q][a][1] = dn[0] * Ji[1] + dn[1] * Ji[3];
            sNr[q][a] = Nf[q * NN + a] / r;
        }
        const int64_t ip = p.ip_off + e * NIP + q;
        double sig[6];
#pragma unroll
        for (int c = 0; c < 6; c++) sig[c] = p.state[(int64_t)c * p.nip_total + ip];
        const double dlam = p.state[(int64_t)13 * p.nip_total + ip];
        const MatPar mp = load_mat(p.mat_kind, p.mat_par, p.emat[e]);
        double D[36];
        const int st = am_calcD(mp, sig, dlam, D);
        if (st) atomicMax(p.status, st);
        const int rows[4] = {0, 1, 2, 5};
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) sD[q][i * 4 + j] = D[6 * rows[i] + rows[j]];
    }
    __syncthreads();
    for (int ab = tid; ab < NN * NN; ab += 64) {
        const int a = ab / NN, b = ab - a * NN;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int q = 0; q < NIP; q++) {
            // B_n (rows rr, zz, θθ, rz; columns u_r, u_z) = [g0 0; 0 g1; N/r 0; g1/√2 g0/√2]
            const double Ba[4][2] = {{sG[q][a][0], 0.0}, {0.0, sG[q][a][1]}, {sNr[q][a], 0.0},
                                     {sG[q][a][1] / AM_SR2, sG[q][a][0] / AM_SR2}};
            const double Bb[4][2] = {{sG[q][b][0], 0.0}, {0.0, sG[q][b][1]}, {sNr[q][b], 0.0},
                                     {sG[q][b][1] / AM_SR2, sG[q][b][0] / AM_SR2}};
            double DB[4][2];
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 2; j++) {
                    double v = 0.0;
                    for (int k = 0; k < 4; k++) v += sD[q][i * 4 + k] * Bb[k][j];
                    DB[i][j] = v;
                }
            for (int i = 0; i < 2; i++)
                for (int j = 0; j < 2; j++) {
                    double v = 0.0;
                    for (int k = 0; k < 4; k++) v += Ba[k][i] * DB[k][j];
                    acc[i * 2 + j] += sc[q] * v;
                }
        }
        const int32_t dst = p.map[e * (int64_t)(NN * NN) + ab];
        if (dst >= 0) {
            double *Kb = p.K + (int64_t)dst * 4;
#pragma unroll
            for (int k = 0; k < 4; k++) Kb[k] += acc[k];
        }
    }
}

// consistent mass: M_ab = Σ_ip ρ·detJ·w·th·N_a·N_b · I   (mech-solid.jl:169-205; th = 2π·r when axisymmetric, :180)
struct MassArgs {
    const double *coords;
    const int32_t *conn;
    const int32_t *map;
    const double *rho;
    const double *dNdR;
    const double *N;
    const double *w;
    double th;
    int axi;
    double *M;
    int *status;
    int64_t e_begin, e_end;
};

template <int NN, int ND, int NIP, int NT>
__global__ void __launch_bounds__(NT) k_assemble_M(MassArgs p) {
    // one element per CTA; small kernel, run once per analysis (M is constant)
    __shared__ double sX[NN * ND], sc[NIP];
    const int tid = threadIdx.x;
    const int64_t e = p.e_begin + blockIdx.x;
    for (int i = tid; i < NN; i += NT) {
        const int64_t node = p.conn[e * NN + i];
#pragma unroll
        for (int d = 0; d < ND; d++) sX[i * ND + d] = p.coords[node * 3 + d];
    }
    __syncthreads();
    if (tid < NIP) {
        double Ji[ND * ND];
        const double det = am_jacobian<NN, ND>(sX, p.dNdR + tid * NN * ND, Ji);
        if (!(det > 0.0)) atomicMax(p.status, AMARU_FAIL_NEG_JACOBIAN);
        double th = p.th;
        if (p.axi) {
            double r = 0.0;
            for (int a = 0; a < NN; a++) r += p.N[tid * NN + a] * sX[a * ND];
            th = 2.0 * 3.14159265358979323846 * r;
        }
        sc[tid] = p.rho[e] * det * p.w[tid] * th;
    }
    __syncthreads();
    for (int ab = tid; ab < NN * NN; ab += NT) {
        const int a = ab / NN, b = ab - a * NN;
        double v = 0.0;
        for (int q = 0; q < NIP; q++) v += sc[q] * p.N[q * NN + a] * p.N[q * NN + b];
        const int32_t dst = p.map[e * (int64_t)(NN * NN) + ab];
        if (dst >= 0) {
            double *Mb = p.M + (int64_t)dst * ND * ND;
#pragma unroll
            for (int d = 0; d < ND; d++) Mb[d * ND + d] += v;
        }
    }
}

// scatter map: destination block of every (element, a, b) by binary search in the sorted block row of node a
__global__ void k_build_map(int nn, int64_t nelem, int64_t nrows, const int32_t *conn, const int32_t *rowptr,
                            const int32_t *col, int32_t *map) {
    const int64_t total = nelem * nn * nn;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i / (nn * nn);
        const int ab = (int)(i - e * nn * nn);
        const int a = ab / nn, b = ab - a * nn;
        const int32_t row = conn[e * nn + a], c = conn[e * nn + b];
        int32_t res = -1;
        if (row < nrows) {
            int32_t lo = rowptr[row], hi = rowptr[row + 1] - 1;
            while (lo <= hi) {
                const int32_t mid = (lo + hi) >> 1;
                const int32_t v = col[mid];
                if (v == c) {
                    res = mid;
                    break;
                }
                if (v < c)
                    lo = mid + 1;
                else
                    hi = mid - 1;
            }
        }
        map[i] = res;
    }
}

template <int NN, int ND, int NIP, int EPB, int NT>
void launch_K(amaru_model *m, Batch &b) {
    using L = AsmSmem<NN, ND, NIP, EPB>;
    const size_t smem = L::doubles * sizeof(double);
    if (m->stressmodel == AMARU_STRESS_AXISYMMETRIC) {
        if constexpr (ND == 2) {
            AsmArgs a;
            a.coords = m->d_coords; a.conn = b.d_conn; a.emat = b.d_emat; a.map = b.d_map;
            a.mat_kind = m->d_mat_kind; a.mat_par = m->d_mat_par; a.dNdR = b.d_dNdR; a.w = b.d_w;
            a.state = m->d_state; a.nip_total = m->nip_total; a.ip_off = b.ip_off; a.th = m->th;
            a.K = m->d_K; a.status = m->d_status;
            for (size_t c = 0; c + 1 < b.color_off.size(); c++) {
                a.e_begin = b.color_off[c];
                a.e_end = b.color_off[c + 1];
                const int64_t n = a.e_end - a.e_begin;
                if (n <= 0) continue;
                k_assemble_K_axi<NN, NIP><<<(unsigned)n, 64, 0, m->stream>>>(a, b.d_N);
                m->launches++;
            }
            CUDA_CHECK(cudaGetLastError());
        }
        return;
    }
    // function attributes are per device (multi-GPU handles drive several devices from one process): one bit per ordinal
    static std::atomic<uint64_t> attr_set{0};
    const uint64_t bit = 1ull << (m->device & 63);
    if (!(attr_set.load(std::memory_order_acquire) & bit)) {
        CUDA_CHECK(cudaFuncSetAttribute(k_assemble_K<NN, ND, NIP, EPB, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
        attr_set.fetch_or(bit, std::memory_order_release);
    }
    AsmArgs a;
    a.coords = m->d_coords; a.conn = b.d_conn; a.emat = b.d_emat; a.map = b.d_map;
    a.mat_kind = m->d_mat_kind; a.mat_par = m->d_mat_par; a.dNdR = b.d_dNdR; a.w = b.d_w;
    a.state = m->d_state; a.nip_total = m->nip_total; a.ip_off = b.ip_off; a.th = m->th;
    a.K = m->d_K; a.status = m->d_status;
    for (size_t c = 0; c + 1 < b.color_off.size(); c++) {
        a.e_begin = b.color_off[c];
        a.e_end = b.color_off[c + 1];
        const int64_t n = a.e_end - a.e_begin;
        if (n <= 0) continue;
        const unsigned grid = (unsigned)((n + EPB - 1) / EPB);
        k_assemble_K<NN, ND, NIP, EPB, NT><<<grid, NT, smem, m->stream>>>(a);
        m->launches++;
    }
    CUDA_CHECK(cudaGetLastError());
}

template <int NN, int ND, int NIP>
void launch_M(amaru_model *m, Batch &b) {
    MassArgs a;
    a.coords = m->d_coords; a.conn = b.d_conn; a.map = b.d_map; a.rho = b.d_rho; a.dNdR = b.d_dNdR; a.N = b.d_N;
    a.w = b.d_w; a.th = m->th; a.axi = m->stressmodel == AMARU_STRESS_AXISYMMETRIC; a.M = m->d_M; a.status = m->d_status;
    for (size_t c = 0; c + 1 < b.color_off.size(); c++) {
        a.e_begin = b.color_off[c];
        a.e_end = b.color_off[c + 1];
        const int64_t n = a.e_end - a.e_begin;
        if (n <= 0) continue;
        k_assemble_M<NN, ND, NIP, 128><<<(unsigned)n, 128, 0, m->stream>>>(a);
        m->launches++;
    }
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace

void amaru_build_map(amaru_model *m, Batch &b) {
    const int64_t total = b.nelem * b.nn * b.nn;
    if (total == 0) return;
    const int64_t blocks = std::min<int64_t>((total + 255) / 256, (int64_t)m->nsm * 32);
    k_build_map<<<(unsigned)blocks, 256, 0, m->stream>>>(b.nn, b.nelem, m->nowned, b.d_conn, m->d_rowptr, m->d_col,
                                                         b.d_map);
    m->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void amaru_launch_assemble(amaru_model *m, int what) {
    const size_t bytes = (size_t)m->nblk * m->nd * m->nd * sizeof(double);
    CUDA_CHECK(cudaMemsetAsync(what == 0 ? m->d_K : m->d_M, 0, bytes, m->stream));
    for (Batch &b : m->batches) {
        if (what == 0) {
            switch (b.shape) {
            case AMARU_SHAPE_QUAD4: launch_K<4, 2, 4, 8, 128>(m, b); break;
            case AMARU_SHAPE_QUAD8: launch_K<8, 2, 4, 4, 256>(m, b); break;
            case AMARU_SHAPE_HEX8: launch_K<8, 3, 8, 4, 256>(m, b); break;
            case AMARU_SHAPE_HEX20: launch_K<20, 3, 8, 1, 128>(m, b); break;
            case AMARU_SHAPE_TET10: launch_K<10, 3, 4, 4, 128>(m, b); break;
            default: throw AmaruError{AMARU_ERR_UNSUPPORTED, "assemble: unsupported shape"};
            }
        } else {
            switch (b.shape) {
            case AMARU_SHAPE_QUAD4: launch_M<4, 2, 4>(m, b); break;
            case AMARU_SHAPE_QUAD8: launch_M<8, 2, 4>(m, b); break;
            case AMARU_SHAPE_HEX8: launch_M<8, 3, 8>(m, b); break;
            case AMARU_SHAPE_HEX20: launch_M<20, 3, 8>(m, b); break;
            case AMARU_SHAPE_TET10: launch_M<10, 3, 4>(m, b); break;
            default: throw AmaruError{AMARU_ERR_UNSUPPORTED, "assemble: unsupported shape"};
            }
        }
    }
}
