// K5-K8: the linear-system step.  Replaces solve_system! (reference src/solver.jl:5-79): the sparse direct
// lu(K11)\(F1 - K12*U2) (solver.jl:38-43) becomes a Jacobi / block-Jacobi preconditioned CG on the device, and the
// K12/K21/K22 slice products (solver.jl:32,38,57) become full-matrix products with the prescribed dofs masked.
//
// Storage: node-blocked CSR (nd x nd blocks) in node-major dof order; the reference's unknown-first eq_id numbering
// exists only at the ABI (k_eq_to_nodes / k_nodes_to_eq).  Prescribed dofs stay in the matrix; CG runs on the full
// vector space with the rows of prescribed dofs masked to zero, which is algebraically the K11 system.
//
// Kernels here (all FP64, hand-written; the SpMV lives in spmv.cu):
//   k_cg_init      r = b - A[0;U2] on free dofs, z = M⁻¹r, p = z, partial r·z and b·b
//   k_cg_update    x += αp, r -= αq, z = M⁻¹r, partial r·z and r·r                                               (K5-K7)
//   k_cg_pupdate   p = z + βp                                                                                    (K6)
//   k_block_inverse  Jacobi / block-Jacobi setup from the diagonal blocks, prescribed rows/cols -> identity      (K7)
// Scalars (α, β, convergence flag, iteration count) live on the device; reductions are two-stage with a fixed grid
// and a fixed summation order (last-block pattern), so results are bitwise reproducible run to run.  The host only
// polls the convergence flag every CG_BATCH iterations; kernels of iterations launched past convergence exit at once.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "p2p.cuh"
#include "reduce.cuh"

namespace {

constexpr int CG_BATCH = 20;
constexpr int ROW_THREADS = 256;

// z = M⁻¹ r for one node.  Block-Jacobi: the inverse of the (symmetric) diagonal block is stored as its upper triangle,
// BS(BS+1)/2 doubles per node in row order (00 01 02 11 12 22): 48 instead of 72 bytes per node in every CG iteration.
template <int BS, bool BLOCKJ>
__device__ __forceinline__ void apply_minv(const double *__restrict__ Minv, int64_t node, const double *r, double *z) {
    if (BLOCKJ) {
        constexpr int NS = BS * (BS + 1) / 2;
        const double *M = Minv + node * NS;
        double m[NS];
#pragma unroll
        for (int k = 0; k < NS; k++) m[k] = M[k];
        if constexpr (BS == 3) {
            z[0] = m[0] * r[0] + m[1] * r[1] + m[2] * r[2];
            z[1] = m[1] * r[0] + m[3] * r[1] + m[4] * r[2];
            z[2] = m[2] * r[0] + m[4] * r[1] + m[5] * r[2];
        } else {
            z[0] = m[0] * r[0] + m[1] * r[1];
            z[1] = m[1] * r[0] + m[2] * r[1];
        }
    } else {
#pragma unroll
        for (int i = 0; i < BS; i++) z[i] = Minv[node * BS + i] * r[i];
    }
}

// r = fixed ? 0 : b - t ; x_free = 0 ; z = M⁻¹r ; p = z ; rz, bb
template <int BS, bool BLOCKJ>
__global__ void __launch_bounds__(ROW_THREADS)
k_cg_init(int64_t nnodes, const double *__restrict__ b, const double *__restrict__ t, const uint8_t *__restrict__ fixed,
          const double *__restrict__ Minv, double *x, double *r, double *z, double *p, double *partial,
          CgScalars *scal, double tol2, int maxit, int finalize) {
    double s[2] = {0.0, 0.0};
    for (int64_t n = blockIdx.x * (int64_t)ROW_THREADS + threadIdx.x; n < nnodes; n += (int64_t)gridDim.x * ROW_THREADS) {
        double rr[BS], zz[BS];
#pragma unroll
        for (int i = 0; i < BS; i++) {
            const int64_t k = n * BS + i;
            const bool fx = fixed[k] != 0;
            rr[i] = fx ? 0.0 : b[k] - t[k];
            if (!fx) x[k] = 0.0;
            r[k] = rr[i];
        }
        apply_minv<BS, BLOCKJ>(Minv, n, rr, zz);
#pragma unroll
        for (int i = 0; i < BS; i++) {
            z[n * BS + i] = zz[i];
            p[n * BS + i] = zz[i];
            s[0] += rr[i] * zz[i];
            s[1] += rr[i] * rr[i];
        }
    }
    block_sum<2, ROW_THREADS>(s);
    if (publish_partials<2>(s, partial, &scal->counter[1])) {
        sum_partials<2, ROW_THREADS>(s, partial);
        if (threadIdx.x == 0) {
            scal->rz_new = s[0];
            scal->rr = s[1];
            scal->acc[0] = s[0];   // multi-GPU: all-reduced in place, then k_finalize_scalars
            scal->acc[1] = s[1];
            if (finalize) {
                scal->rz_old = s[0];
                scal->bb = s[1];
                scal->tol2 = tol2;
                scal->maxit = maxit;
                scal->iters = 0;
                scal->done = (s[1] == 0.0) ? 1 : 0;
            }
        }
    }
}

template <int BS, bool BLOCKJ>
__global__ void __launch_bounds__(ROW_THREADS)
k_cg_update(int64_t nnodes, const double *__restrict__ p, const double *__restrict__ q,
            const double *__restrict__ Minv, double *x, double *r, double *z, double *partial, CgScalars *scal,
            int finalize) {
    if (scal->done) return;
    // single GPU: alpha was finalised by the SpMV's last block; multi-GPU: acc[0] holds the all-reduced p.Ap
    const double pq_all = scal->acc[0];
    const double alpha = finalize ? scal->alpha : scal->rz_old / pq_all;
    double s[2] = {0.0, 0.0};
    for (int64_t n = blockIdx.x * (int64_t)ROW_THREADS + threadIdx.x; n < nnodes; n += (int64_t)gridDim.x * ROW_THREADS) {
        double rr[BS], zz[BS];
#pragma unroll
        for (int i = 0; i < BS; i++) {
            const int64_t k = n * BS + i;
            x[k] += alpha * p[k];
            rr[i] = r[k] - alpha * q[k];
            r[k] = rr[i];
        }
        apply_minv<BS, BLOCKJ>(Minv, n, rr, zz);
#pragma unroll
        for (int i = 0; i < BS; i++) {
            z[n * BS + i] = zz[i];
            s[0] += rr[i] * zz[i];
            s[1] += rr[i] * rr[i];
        }
    }
    block_sum<2, ROW_THREADS>(s);
    if (publish_partials<2>(s, partial, &scal->counter[1])) {
        sum_partials<2, ROW_THREADS>(s, partial);
        if (threadIdx.x == 0) {
            scal->rz_new = s[0];
            scal->rr = s[1];
            if (!finalize) scal->pq = pq_all;   // kept for the breakdown test of k_finalize_scalars (acc is reused below)
            scal->acc[0] = s[0];
            scal->acc[1] = s[1];
            if (finalize) {
                scal->beta = s[0] / scal->rz_old;
                scal->rz_old = s[0];
                scal->iters += 1;
                if (s[1] <= scal->tol2 * scal->bb) scal->done = 1;
                else if (scal->iters >= scal->maxit) scal->done = 2;
                else if (!(s[1] == s[1])) scal->done = 3;
            }
        }
    }
}


// ---- fused multi-GPU loop (p2p.cuh): the scalar all-reduces and the halo push live inside the vector kernels ---------------
// k_cg_update_f: every CTA collects the all-reduced p.Ap (pushed by the operator kernel's last CTA) -> alpha; the last CTA
// pushes this rank's {r.z, r.r} to every rank and advances scal_epoch.
template <int BS, bool BLOCKJ>
__global__ void __launch_bounds__(ROW_THREADS)
k_cg_update_f(int64_t nnodes, const double *__restrict__ p, const double *__restrict__ q, const double *__restrict__ Minv,
              double *x, double *r, double *z, double *partial, CgScalars *scal, P2PFused fz) {
    if (scal->done) return;
    __shared__ double s_pq;
    __shared__ int s_ok;
    P2PWin *me = fz.pd.win[fz.pd.rank];
    const unsigned long long base = *reinterpret_cast<volatile unsigned long long *>(&me->scal_epoch);
    if (threadIdx.x < 32) {
        double v[1] = {0.0};
        const bool ok = p2p_collect(fz.pd, base + 1, v, 1);
        if (threadIdx.x == 0) {
            s_ok = ok ? 1 : 0;
            s_pq = v[0];
        }
    }
    __syncthreads();
    if (!s_ok) return;                                   // a peer left the sequence: the host reports AMARU_ERR_COMM
    const double pq_all = s_pq;
    const double alpha = scal->rz_old / pq_all;
    double s[2] = {0.0, 0.0};
    for (int64_t n = blockIdx.x * (int64_t)ROW_THREADS + threadIdx.x; n < nnodes; n += (int64_t)gridDim.x * ROW_THREADS) {
        double rr[BS], zz[BS];
#pragma unroll
        for (int i = 0; i < BS; i++) {
            const int64_t k = n * BS + i;
            x[k] += alpha * p[k];
            rr[i] = r[k] - alpha * q[k];
            r[k] = rr[i];
        }
        apply_minv<BS, BLOCKJ>(Minv, n, rr, zz);
#pragma unroll
        for (int i = 0; i < BS; i++) {
            z[n * BS + i] = zz[i];
            s[0] += rr[i] * zz[i];
            s[1] += rr[i] * rr[i];
        }
    }
    block_sum<2, ROW_THREADS>(s);
    if (publish_partials<2>(s, partial, &scal->counter[1])) {
        sum_partials<2, ROW_THREADS>(s, partial);
        __shared__ double s_out[2];
        if (threadIdx.x == 0) {
            s_out[0] = s[0];
            s_out[1] = s[1];
            scal->pq = pq_all;                           // for the breakdown test of the p-update
        }
        __syncthreads();
        if (threadIdx.x < 32) p2p_push_scalars(fz.pd, base + 2, s_out, 2, threadIdx.x);
        if (threadIdx.x == 0) me->scal_epoch = base + 1;
    }
}

// k_cg_pupdate_f: every CTA collects the all-reduced {r.z, r.r} -> beta and the convergence decision (identical everywhere);
// p = z + beta p with the entries of boundary nodes stored straight into the neighbours' ghost slots; the last CTA raises
// the halo flags (the next operator kernel waits for them), does the scalar bookkeeping and advances scal_epoch.
template <int BS>
__global__ void __launch_bounds__(ROW_THREADS)
k_cg_pupdate_f(int64_t nnodes, const double *__restrict__ z, double *p, CgScalars *scal, P2PFused fz) {
    if (scal->done) return;
    __shared__ double s_v[2];
    __shared__ int s_ok, s_last;
    P2PWin *me = fz.pd.win[fz.pd.rank];
    const unsigned long long base = *reinterpret_cast<volatile unsigned long long *>(&me->scal_epoch);
    const unsigned long long he = *reinterpret_cast<volatile unsigned long long *>(&me->halo_epoch) + 1ull;
    if (threadIdx.x < 32) {
        double v[2] = {0.0, 0.0};
        const bool ok = p2p_collect(fz.pd, base + 1, v, 2);
        if (threadIdx.x == 0) {
            s_ok = ok ? 1 : 0;
            s_v[0] = v[0];
            s_v[1] = v[1];
        }
    }
    __syncthreads();
    if (!s_ok) return;
    const double rz_new = s_v[0], rr = s_v[1];
    const double beta = rz_new / scal->rz_old;
    int done = 0;
    if (!(scal->pq > 0.0)) done = 3;                     // not SPD / breakdown
    else if (rr <= scal->tol2 * scal->bb) done = 1;
    else if (scal->iters + 1 >= scal->maxit) done = 2;
    else if (!(rr == rr)) done = 3;
    if (!done) {
        for (int64_t n = blockIdx.x * (int64_t)ROW_THREADS + threadIdx.x; n < nnodes; n += (int64_t)gridDim.x * ROW_THREADS) {
            double pn[BS];
#pragma unroll
            for (int i = 0; i < BS; i++) {
                pn[i] = z[n * BS + i] + beta * p[n * BS + i];
                p[n * BS + i] = pn[i];
            }
            const int32_t b = fz.bidx[n];
            if (b >= 0)
                for (int32_t e = fz.bent_ptr[b]; e < fz.bent_ptr[b + 1]; e++) {
                    double *dst = fz.peer_p[fz.bent_q[e]] + fz.bent_remote[e] * BS;
#pragma unroll
                    for (int i = 0; i < BS; i++) dst[i] = pn[i];
                }
        }
        __threadfence_system();
    }
    __syncthreads();                                     // every thread read the scalars and fenced its remote stores
    if (threadIdx.x == 0) s_last = atomicInc(fz.counter, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    if (!done && threadIdx.x < fz.nneigh) {
        __threadfence_system();
        st_release_sys(&fz.pd.win[fz.neigh[threadIdx.x]]->hflag[fz.pd.rank], he);
    }
    if (threadIdx.x == 0) {
        scal->beta = beta;
        scal->rz_old = rz_new;
        scal->rr = rr;
        scal->iters += 1;
        scal->done = done;
        me->scal_epoch = base + 1;
    }
}

// multi-GPU: scalar bookkeeping after the all-reduce of the partial dots (1 thread)
__global__ void k_finalize_scalars(CgScalars *scal, int stage, double tol2, int maxit) {
    if (stage == 0) {  // after init: acc = {rz, bb}
        scal->rz_old = scal->acc[0];
        scal->bb = scal->acc[1];
        scal->tol2 = tol2;
        scal->maxit = maxit;
        scal->iters = 0;
        scal->done = (scal->acc[1] == 0.0) ? 1 : 0;
    } else if (stage == 1) {  // after spmv: acc = {pq}
        if (scal->done) return;
        if (!(scal->acc[0] > 0.0)) scal->done = 3;
        scal->pq = scal->acc[0];
        scal->alpha = scal->rz_old / scal->acc[0];
    } else {  // after update: acc = {rz_new, rr}; a breakdown (p.Ap <= 0 or NaN) shows up as a non-finite r.r here
        if (scal->done) return;
        scal->beta = scal->acc[0] / scal->rz_old;
        scal->rz_old = scal->acc[0];
        scal->rr = scal->acc[1];
        scal->iters += 1;
        if (!(scal->pq > 0.0)) scal->done = 3;   // not SPD / breakdown (spmv_dot_epilogue's test on one GPU)
        else if (scal->acc[1] <= scal->tol2 * scal->bb) scal->done = 1;
        else if (scal->iters >= scal->maxit) scal->done = 2;
        else if (!(scal->acc[1] == scal->acc[1])) scal->done = 3;
    }
}

__global__ void __launch_bounds__(ROW_THREADS)
k_cg_pupdate(int64_t n, const double *__restrict__ z, double *p, const CgScalars *scal) {
    if (scal->done) return;
    const double beta = scal->beta;
    for (int64_t i = blockIdx.x * (int64_t)ROW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * ROW_THREADS)
        p[i] = z[i] + beta * p[i];
}

// Jacobi / block-Jacobi setup: invert the diagonal block of each owned node with prescribed rows/cols -> identity
template <int BS, bool BLOCKJ>
__global__ void k_block_inverse(int64_t nnodes, const int32_t *__restrict__ diag, const double *__restrict__ A,
                                const uint8_t *__restrict__ fixed, double *Minv) {
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < nnodes; n += (int64_t)gridDim.x * blockDim.x) {
        double D[BS * BS];
        const int32_t d = diag[n];
#pragma unroll
        for (int i = 0; i < BS; i++)
#pragma unroll
            for (int j = 0; j < BS; j++) {
                const bool fx = fixed[n * BS + i] || fixed[n * BS + j];
                D[i * BS + j] = (d < 0 || fx) ? (i == j ? 1.0 : 0.0) : A[(int64_t)d * BS * BS + i * BS + j];
            }
        if (!BLOCKJ) {
#pragma unroll
            for (int i = 0; i < BS; i++) Minv[n * BS + i] = 1.0 / D[i * BS + i];
        } else if (BS == 2) {
            // the diagonal block of K is symmetric up to rounding (it is summed from symmetric element blocks in a fixed order);
            // the stored inverse is that of its symmetric part, so that M⁻¹ is exactly symmetric (PCG needs an SPD preconditioner)
            const double o = 0.5 * (D[1] + D[2]);
            const double det = D[0] * D[3] - o * o;
            Minv[n * 3 + 0] = D[3] / det; Minv[n * 3 + 1] = -o / det; Minv[n * 3 + 2] = D[0] / det;
        } else {
            const double a01 = 0.5 * (D[1] + D[3]), a02 = 0.5 * (D[2] + D[6]), a12 = 0.5 * (D[5] + D[7]);
            const double a00 = D[0], a11 = D[4], a22 = D[8];
            const double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
            const double det = a00 * c00 + a01 * c01 + a02 * c02;
            double *M = Minv + n * 6;
            M[0] = c00 / det; M[1] = c01 / det; M[2] = c02 / det;
            M[3] = (a00 * a22 - a02 * a02) / det; M[4] = (a01 * a02 - a00 * a12) / det;
            M[5] = (a00 * a11 - a01 * a01) / det;
        }
    }
}

// ------------------------------------------------------------------------------------------------ ABI order <-> node order
__global__ void k_eq_to_nodes(int64_t n, const int32_t *__restrict__ eqid, const double *__restrict__ src, double *dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[eqid[i]];
}
__global__ void k_nodes_to_eq(int64_t n, const int32_t *__restrict__ eqid, const uint8_t *__restrict__ fixed,
                              const double *__restrict__ src, double *dst, int which) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool fx = fixed[i] != 0;
        if (which == 0 || (which == 1 && !fx) || (which == 2 && fx)) dst[eqid[i]] = src[i];
    }
}
__global__ void k_axpby_matrix(int64_t n, double a, const double *K, double b, const double *M, double *A) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        A[i] = a * K[i] + b * M[i];
}
__global__ void k_zero_free(int64_t n, const uint8_t *__restrict__ fixed, double *x) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (!fixed[i]) x[i] = 0.0;
}
__global__ void k_maxabs_nan(int64_t n, const double *__restrict__ v, const uint8_t *__restrict__ fixed, int only_free,
                             CgScalars *scal) {
    double mx = 0.0;
    int nan = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (only_free && fixed[i]) continue;
        const double a = fabs(v[i]);
        if (a != a) nan = 1;
        else if (a > mx) mx = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        nan |= __shfl_xor_sync(0xffffffffu, nan, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&scal->maxabs_bits, (unsigned long long)__double_as_longlong(mx));  // order-independent, exact
        if (nan) atomicOr(&scal->nanflag, 1);
    }
}
__global__ void k_stage_maxabs(CgScalars *scal) {
    scal->acc[0] = scal->nanflag ? __longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double((long long)scal->maxabs_bits);
}
__global__ void k_reset_flags(CgScalars *scal) {
    scal->maxabs_bits = 0ull;
    scal->nanflag = 0;
}

int blocks_for(amaru_model *m, int64_t n, int threads) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, (int64_t)m->nsm * 8));
}
int node_grid(amaru_model *m, int64_t n) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + ROW_THREADS - 1) / ROW_THREADS, (int64_t)m->nsm * 8));
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host side
void amaru_pcg_setup(amaru_model *m) {
    const size_t nv = (size_t)m->nnodes * m->nd * sizeof(double);
    CUDA_CHECK(cudaMalloc(&m->d_x, nv));
    CUDA_CHECK(cudaMalloc(&m->d_r, nv));
    CUDA_CHECK(cudaMalloc(&m->d_z, nv));
    CUDA_CHECK(cudaMalloc(&m->d_p, nv));
    CUDA_CHECK(cudaMalloc(&m->d_q, nv));
    CUDA_CHECK(cudaMalloc(&m->d_b, nv));
    CUDA_CHECK(cudaMalloc(&m->d_f, nv));
    for (double *v : {m->d_x, m->d_r, m->d_z, m->d_p, m->d_q, m->d_b, m->d_f}) CUDA_CHECK(cudaMemset(v, 0, nv));
    CUDA_CHECK(cudaMalloc(&m->d_Minv, (size_t)std::max<int64_t>(m->nowned, 1) * m->nd * m->nd * sizeof(double)));
    amaru_spmv_setup(m);
    const int maxgrid = std::max(std::max(m->grid_rows, m->grid_tma), m->nsm * 8);
    CUDA_CHECK(cudaMalloc(&m->d_partial, (size_t)4 * maxgrid * sizeof(double)));
    CUDA_CHECK(cudaMalloc(&m->d_scal, sizeof(CgScalars)));
    CUDA_CHECK(cudaMemset(m->d_scal, 0, sizeof(CgScalars)));
    CUDA_CHECK(cudaMallocHost(&m->h_pinned, sizeof(CgScalars) + 64));
    const char *eg = getenv("AMARU_CG_GRAPH");
    m->cg_graph = !(eg && atoi(eg) == 0);
}

// y = A x on the owned rows; mask_mode 1 zeroes the rows of prescribed dofs.  x must hold valid ghost entries.
void amaru_spmv(amaru_model *m, const double *A, const double *x, double *y, int mask_mode) {
    amaru_spmv_launch(m, A, x, y, mask_mode, 0, 0, 0);
}

// y = A x with the system matrix of the solve (unmasked, no dot): matrix-free or from the assembled block-CSR values
static void amaru_system_product(amaru_model *m, const double *x, double *y) {
    if (m->op_ebe && !m->blended) amaru_ebe_apply(m, x, y, 0, 0, 0, 0);
    else amaru_spmv(m, m->d_A, x, y, 0);
}

static void spmv_dot(amaru_model *m, const double *A, const double *x, double *y, int finalize, int fused = 0) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (m->profiling) {   // the pool owns the events from the moment they exist (released with the profile / the handle)
        CUDA_CHECK(cudaEventCreate(&e0));
        m->ev_pool.push_back(e0);
        CUDA_CHECK(cudaEventCreate(&e1));
        m->ev_pool.push_back(e1);
        CUDA_CHECK(cudaEventRecord(e0, m->stream));
    }
    if (m->op_ebe && !m->blended && A == m->d_A) amaru_ebe_apply(m, x, y, 1, 1, 1, finalize, fused);        // matrix-free (ebe.cu)
    else if (m->use_sym && A == m->d_A) amaru_spmv_sym_launch(m, x, y, 1, 1, 1, finalize);   // half the bytes (spmv.cu)
    else amaru_spmv_launch(m, A, x, y, 1, 1, 1, finalize);
    if (m->profiling) CUDA_CHECK(cudaEventRecord(e1, m->stream));
}

// test / measurement hook: q = A p with the handle's CG operator; masked != 0 also runs the fused p.Ap (returned)
double amaru_operator_product(amaru_model *m, int masked) {
    amaru_spmv_sym_refresh(m);
    CUDA_CHECK(cudaMemsetAsync(m->d_scal, 0, sizeof(CgScalars), m->stream));
    amaru_ebe_begin(m);
    if (m->nranks > 1) amaru_halo_exchange(m, m->d_p);
    if (masked) {
        const bool prof = m->profiling;
        m->profiling = false;
        spmv_dot(m, m->d_A, m->d_p, m->d_q, 0);
        m->profiling = prof;
        if (m->nranks > 1) amaru_allreduce_sum(m, m->d_scal->acc, 1);
    } else {
        amaru_system_product(m, m->d_p, m->d_q);
    }
    CgScalars *h = reinterpret_cast<CgScalars *>(m->h_pinned);
    CUDA_CHECK(cudaMemcpyAsync(h, m->d_scal, sizeof(CgScalars), cudaMemcpyDeviceToHost, m->stream));
    CUDA_CHECK(cudaStreamSynchronize(m->stream));
    return m->nranks > 1 ? h->acc[0] : h->pq;
}

static void build_preconditioner(amaru_model *m, int precond) {
    if (m->minv_kind == precond) return;
    const int g = blocks_for(m, m->nowned, 256);
    const bool bj = precond == AMARU_PRECOND_BLOCK_JACOBI;
    if (m->nd == 3) {
        if (bj) k_block_inverse<3, true><<<g, 256, 0, m->stream>>>(m->nowned, m->d_diag, m->d_A, m->d_fixed, m->d_Minv);
        else k_block_inverse<3, false><<<g, 256, 0, m->stream>>>(m->nowned, m->d_diag, m->d_A, m->d_fixed, m->d_Minv);
    } else {
        if (bj) k_block_inverse<2, true><<<g, 256, 0, m->stream>>>(m->nowned, m->d_diag, m->d_A, m->d_fixed, m->d_Minv);
        else k_block_inverse<2, false><<<g, 256, 0, m->stream>>>(m->nowned, m->d_diag, m->d_A, m->d_fixed, m->d_Minv);
    }
    m->launches++;
    CUDA_CHECK(cudaGetLastError());
    m->minv_kind = precond;
}

template <int BS, bool BJ>
static void cg_loop(amaru_model *m, double rtol, int maxit, SolveInfo &info) {
    const int gn = node_grid(m, m->nowned);
    const bool multi = m->nranks > 1;
    const int fin = multi ? 0 : 1;
    CgScalars *h = reinterpret_cast<CgScalars *>(m->h_pinned);
    // t = A*[0;U2] (the caller zeroed the free entries of d_x), r = b - t on the free dofs
    if (multi) amaru_halo_exchange(m, m->d_x);
    amaru_system_product(m, m->d_x, m->d_q);
    k_cg_init<BS, BJ><<<gn, ROW_THREADS, 0, m->stream>>>(m->nowned, m->d_b, m->d_q, m->d_fixed, m->d_Minv, m->d_x, m->d_r,
                                                          m->d_z, m->d_p, m->d_partial, m->d_scal, rtol * rtol, maxit, fin);
    m->launches++;
    if (multi) {   // the last block of k_cg_init left the local {r.z, b.b} in acc
        amaru_allreduce_sum(m, m->d_scal->acc, 2);
        k_finalize_scalars<<<1, 1, 0, m->stream>>>(m->d_scal, 0, rtol * rtol, maxit);
        m->launches++;
    }
    const int64_t nloc = m->nowned * BS;
    // One batch of CG_BATCH iterations is a fixed sequence of launches with fixed arguments (all scalars live on the device),
    // so on one GPU it is captured once per solve into a CUDA graph and replayed: small systems (config 1: 1 322 dofs,
    // 3 launches of a few microseconds per iteration) are launch-bound otherwise.  Not used while the SpMV is being timed
    // with events, nor on partitioned handles (NCCL calls sit between the kernels).
    cudaGraphExec_t gexec = nullptr;
    // peer-memory exchanges are plain kernels with device-resident epochs, so partitioned handles replay graphs too; with
    // NCCL calls between the kernels (the default between processes until amaru_p2p_enable) the batch is launched directly
    const bool p2p = multi && amaru_comm_is_p2p(m);
    // fused loop: operator + 2 vector kernels per iteration, no stand-alone exchange kernels (DMMA forms of the matrix-free operator)
    const bool fused = p2p && amaru_comm_fused(m) && m->op_ebe && !m->blended && amaru_ebe_fusable(m);
    if (fused) amaru_halo_push(m, m->d_p);               // p of the first iteration; consumed by the first operator kernel
    const bool use_graph = !m->profiling && m->cg_graph && (!multi || p2p);
    bool finished = false;
    int64_t batch_launches = 0;
    auto one_batch = [&]() {
        for (int it = 0; it < CG_BATCH && fused; it++) {
            spmv_dot(m, m->d_A, m->d_p, m->d_q, 0, 1);
            const P2PFused fz = amaru_comm_fused_args(m);
            k_cg_update_f<BS, BJ><<<gn, ROW_THREADS, 0, m->stream>>>(m->nowned, m->d_p, m->d_q, m->d_Minv, m->d_x, m->d_r, m->d_z,
                                                                      m->d_partial, m->d_scal, fz);
            k_cg_pupdate_f<BS><<<gn, ROW_THREADS, 0, m->stream>>>(m->nowned, m->d_z, m->d_p, m->d_scal, fz);
            m->launches += 2;
        }
        for (int it = 0; it < CG_BATCH && !fused; it++) {
            if (multi) amaru_halo_exchange(m, m->d_p);
            spmv_dot(m, m->d_A, m->d_p, m->d_q, fin);
            // multi-GPU: per iteration = halo exchange, operator, all-reduce(p.Ap), update, all-reduce(r.z, r.r), one scalar
            // kernel, p-update; the partial dots are staged by the producing kernels' last blocks and alpha is formed by
            // the consumer (k_cg_update) from the all-reduced value
            if (multi) amaru_allreduce_sum(m, m->d_scal->acc, 1);
            k_cg_update<BS, BJ><<<gn, ROW_THREADS, 0, m->stream>>>(m->nowned, m->d_p, m->d_q, m->d_Minv, m->d_x, m->d_r,
                                                                    m->d_z, m->d_partial, m->d_scal, fin);
            if (multi) {
                amaru_allreduce_sum(m, m->d_scal->acc, 2);
                k_finalize_scalars<<<1, 1, 0, m->stream>>>(m->d_scal, 2, 0.0, 0);
                m->launches++;
            }
            k_cg_pupdate<<<node_grid(m, nloc), ROW_THREADS, 0, m->stream>>>(nloc, m->d_z, m->d_p, m->d_scal);
            m->launches += 2;
        }
    };
    try {
        while (!finished) {
            if (use_graph && gexec) {
                CUDA_CHECK(cudaGraphLaunch(gexec, m->stream));
                m->launches += batch_launches;
            } else if (use_graph) {
                const int64_t l0 = m->launches;
                CUDA_CHECK(cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal));
                cudaGraph_t graph = nullptr;
                try {
                    one_batch();
                } catch (...) {   // leave capture mode before unwinding, else the stream stays unusable
                    cudaStreamEndCapture(m->stream, &graph);
                    if (graph) cudaGraphDestroy(graph);
                    throw;
                }
                CUDA_CHECK(cudaStreamEndCapture(m->stream, &graph));
                const cudaError_t ie = cudaGraphInstantiate(&gexec, graph, 0);
                cudaGraphDestroy(graph);
                CUDA_CHECK(ie);
                CUDA_CHECK(cudaGraphLaunch(gexec, m->stream));
                batch_launches = m->launches - l0;
            } else {
                one_batch();
            }
            CUDA_CHECK(cudaGetLastError());
            CUDA_CHECK(cudaMemcpyAsync(h, m->d_scal, sizeof(CgScalars), cudaMemcpyDeviceToHost, m->stream));
            CUDA_CHECK(cudaStreamSynchronize(m->stream));
            if (p2p) amaru_comm_check(m);
            finished = h->done != 0;
        }
    } catch (...) {
        if (gexec) cudaGraphExecDestroy(gexec);
        throw;
    }
    if (gexec) cudaGraphExecDestroy(gexec);
    info.iters = h->iters;
    info.relres = h->bb > 0.0 ? std::sqrt(h->rr / h->bb) : 0.0;
    info.converged = (h->done == 1);
    if (m->profiling) {
        // launches past convergence return at once: only the first `iters` pairs are real SpMVs
        for (size_t i = 0; i + 1 < m->ev_pool.size(); i += 2) {
            if ((int)(i / 2) < info.iters) {
                float ms = 0.f;
                CUDA_CHECK(cudaEventElapsedTime(&ms, m->ev_pool[i], m->ev_pool[i + 1]));
                m->prof_spmv_ms += ms;
                m->prof_spmv_n++;
            }
            cudaEventDestroy(m->ev_pool[i]);
            cudaEventDestroy(m->ev_pool[i + 1]);
        }
        m->ev_pool.clear();
    }
}

void amaru_zero_free(amaru_model *m, double *x) {
    const int64_t n = m->nnodes * m->nd;
    k_zero_free<<<blocks_for(m, n, 256), 256, 0, m->stream>>>(n, m->d_fixed, x);
    m->launches++;
}

// Solve on node-major vectors: in  d_x = prescribed values at prescribed dofs (free entries ignored),
//                                  d_b = known forces at free dofs;
//                              out d_x = full displacement vector, d_q = A*x (reactions at prescribed dofs).
void amaru_pcg_solve(amaru_model *m, double rtol, int maxit, int precond, SolveInfo &info) {
    build_preconditioner(m, precond);
    amaru_spmv_sym_refresh(m);   // upper blocks of the current system matrix for the CG products
    amaru_ebe_begin(m);
    const int64_t nloc = m->nowned * m->nd;
    amaru_zero_free(m, m->d_x);   // the first product must be K*[0;U2]
    const bool bj = precond == AMARU_PRECOND_BLOCK_JACOBI;
    if (m->nd == 3) {
        if (bj) cg_loop<3, true>(m, rtol, maxit, info);
        else cg_loop<3, false>(m, rtol, maxit, info);
    } else {
        if (bj) cg_loop<2, true>(m, rtol, maxit, info);
        else cg_loop<2, false>(m, rtol, maxit, info);
    }
    // reactions: q = A*x over all rows (solver.jl:32,57: F2 = K22*U2 + K21*U1)
    if (m->nranks > 1) amaru_halo_exchange(m, m->d_x);
    amaru_system_product(m, m->d_x, m->d_q);
    // max |U1| (solver.jl:68-71)
    k_reset_flags<<<1, 1, 0, m->stream>>>(m->d_scal);
    k_maxabs_nan<<<blocks_for(m, nloc, 256), 256, 0, m->stream>>>(nloc, m->d_x, m->d_fixed, 1, m->d_scal);
    m->launches += 2;
    if (m->nranks > 1) {   // max over ranks (a NaN anywhere becomes +inf everywhere)
        k_stage_maxabs<<<1, 1, 0, m->stream>>>(m->d_scal);
        amaru_allreduce_max(m, m->d_scal->acc, 1);
        m->launches++;
    }
    CgScalars *h = reinterpret_cast<CgScalars *>(m->h_pinned);
    CUDA_CHECK(cudaMemcpyAsync(h, m->d_scal, sizeof(CgScalars), cudaMemcpyDeviceToHost, m->stream));
    CUDA_CHECK(cudaStreamSynchronize(m->stream));
    double mx;
    std::memcpy(&mx, &h->maxabs_bits, sizeof(double));
    if (m->nranks > 1) mx = h->acc[0];
    info.maxabs = h->nanflag ? NAN : mx;
}

void amaru_eq_to_nodes(amaru_model *m, const double *d_eq, double *d_nodes) {
    const int64_t n = m->nnodes * m->nd;
    k_eq_to_nodes<<<blocks_for(m, n, 256), 256, 0, m->stream>>>(n, m->d_eqid, d_eq, d_nodes);
    m->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void amaru_nodes_to_eq(amaru_model *m, const double *d_nodes, double *d_eq, int which) {
    const int64_t n = m->nowned * m->nd;   // only rows this rank owns
    k_nodes_to_eq<<<blocks_for(m, n, 256), 256, 0, m->stream>>>(n, m->d_eqid, m->d_fixed, d_nodes, d_eq, which);
    m->launches++;
    CUDA_CHECK(cudaGetLastError());
}

// out = a*x + b*y (elementwise; out may alias x or y)
void amaru_axpby(amaru_model *m, int64_t n, double a, const double *x, double b, const double *y, double *out) {
    if (n <= 0) return;
    k_axpby_matrix<<<blocks_for(m, n, 256), 256, 0, m->stream>>>(n, a, x, b, y, out);
    m->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void amaru_combine_matrix(amaru_model *m) {
    const int64_t n = m->nblk * m->nd * m->nd;
    if (m->sysB == 0.0 && m->sysA == 1.0) {
        m->d_A = m->d_K;
    } else {
        AMARU_REQUIRE(m->d_M != nullptr, AMARU_ERR_ARG, "set_system_matrix: mass matrix not assembled");
        if (!m->d_Abuf) CUDA_CHECK(cudaMalloc(&m->d_Abuf, (size_t)n * sizeof(double) + 256));
        m->d_A = m->d_Abuf;
        k_axpby_matrix<<<blocks_for(m, n, 256), 256, 0, m->stream>>>(n, m->sysA, m->d_K, m->sysB, m->d_M, m->d_A);
        m->launches++;
        CUDA_CHECK(cudaGetLastError());
    }
    m->minv_kind = -1;
    m->sym_fresh = false;
}

// returns 1 if v[0..n) holds a NaN
int amaru_check_nan(amaru_model *m, const double *d_v, int64_t n) {
    k_reset_flags<<<1, 1, 0, m->stream>>>(m->d_scal);
    k_maxabs_nan<<<blocks_for(m, n, 256), 256, 0, m->stream>>>(n, d_v, m->d_fixed, 0, m->d_scal);
    m->launches += 2;
    CgScalars *h = reinterpret_cast<CgScalars *>(m->h_pinned);
    CUDA_CHECK(cudaMemcpyAsync(h, m->d_scal, sizeof(CgScalars), cudaMemcpyDeviceToHost, m->stream));
    CUDA_CHECK(cudaStreamSynchronize(m->stream));
    return h->nanflag;
}

// measurement hook: `reps` back-to-back launches of one CG kernel class (0 SpMV+dot, 3 fused update, 4 p-update)
void amaru_time_cg_kernel(amaru_model *m, int kind, int precond, int reps) {
    const bool bj = precond == AMARU_PRECOND_BLOCK_JACOBI;
    build_preconditioner(m, precond);
    amaru_spmv_sym_refresh(m);
    CUDA_CHECK(cudaMemsetAsync(m->d_scal, 0, sizeof(CgScalars), m->stream));   // done = 0, alpha = beta = 0
    amaru_ebe_begin(m);
    const int gn = node_grid(m, m->nowned);
    const int64_t nloc = m->nowned * m->nd;
    const bool prof = m->profiling;
    m->profiling = false;
    for (int i = 0; i < reps; i++) {
        if (kind == 0) {
            spmv_dot(m, m->d_A, m->d_p, m->d_q, 0);
        } else if (kind == 3) {
            if (m->nd == 3) {
                if (bj) k_cg_update<3, true><<<gn, ROW_THREADS, 0, m->stream>>>(m->nowned, m->d_p, m->d_q, m->d_Minv, m->d_x, m->d_r, m->d_z, m->d_partial, m->d_scal, 0);
                else k_cg_update<3, false><<<gn, ROW_THREADS, 0, m->stream>>>(m->nowned, m->d_p, m->d_q, m->d_Minv, m->d_x, m->d_r, m->d_z, m->d_partial, m->d_scal, 0);
            } else {
                if (bj) k_cg_update<2, true><<<gn, ROW_THREADS, 0, m->stream>>>(m->nowned, m->d_p, m->d_q, m->d_Minv, m->d_x, m->d_r, m->d_z, m->d_partial, m->d_scal, 0);
                else k_cg_update<2, false><<<gn, ROW_THREADS, 0, m->stream>>>(m->nowned, m->d_p, m->d_q, m->d_Minv, m->d_x, m->d_r, m->d_z, m->d_partial, m->d_scal, 0);
            }
            m->launches++;
        } else {
            k_cg_pupdate<<<node_grid(m, nloc), ROW_THREADS, 0, m->stream>>>(nloc, m->d_z, m->d_p, m->d_scal);
            m->launches++;
        }
    }
    m->profiling = prof;
    CUDA_CHECK(cudaGetLastError());
}
