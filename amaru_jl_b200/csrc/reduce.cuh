// Device-side scalar block of the PCG and the deterministic two-stage reduction helpers shared by pcg.cu / spmv.cu.
#pragma once
#include "amaru_internal.h"

struct CgScalars {
    double rz_old, pq, rz_new, rr, bb, alpha, beta, tol2;
    double acc[4];               // scratch for all-reduce (multi-GPU)
    unsigned long long maxabs_bits;
    int done;                    // 0 running, 1 converged, 2 maxit, 3 breakdown
    int iters, maxit;
    unsigned int counter[4];
    int nanflag;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of NV values per thread; result valid in thread 0
template <int NV, int NT>
__device__ __forceinline__ void block_sum(double (&v)[NV]) {
    __shared__ double sh[NV][NT / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        const double s = warp_sum(v[k]);
        if (lane == 0) sh[k][w] = s;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double s = lane < NT / 32 ? sh[k][lane] : 0.0;
            s = warp_sum(s);
            v[k] = s;
        }
    }
    __syncthreads();
}

// publishes this block's partial sums and returns true (for all threads) in the last block to arrive
template <int NV>
__device__ __forceinline__ bool publish_partials(const double (&v)[NV], double *partial, unsigned int *counter) {
    __shared__ bool last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) partial[(size_t)k * gridDim.x + blockIdx.x] = v[k];
        __threadfence();
        const unsigned int t = atomicInc(counter, gridDim.x - 1);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    return last;
}

// fixed-order sum of the per-block partials (called by every thread of the last block); result valid in thread 0
template <int NV, int NT>
__device__ __forceinline__ void sum_partials(double (&v)[NV], const double *partial) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += NT) s += __ldcg(&partial[(size_t)k * gridDim.x + i]);
        v[k] = s;
    }
    block_sum<NV, NT>(v);
}

// p·(Ap) epilogue shared by the SpMV kernels
template <int NT>
__device__ __forceinline__ void spmv_dot_epilogue(double (&dsum)[1], double *partial, CgScalars *scal, int finalize) {
    block_sum<1, NT>(dsum);
    if (publish_partials<1>(dsum, partial, &scal->counter[0])) {
        sum_partials<1, NT>(dsum, partial);
        if (threadIdx.x == 0) {
            scal->pq = dsum[0];
            scal->acc[0] = dsum[0];   // multi-GPU: the local p.Ap is all-reduced in place (no staging kernel)
            if (finalize) {
                if (!(dsum[0] > 0.0)) scal->done = 3;   // not SPD / breakdown
                scal->alpha = scal->rz_old / dsum[0];
            }
        }
    }
}
