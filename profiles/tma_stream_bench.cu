// Microbenchmark: how fast can one B200 stream a large array HBM -> shared memory with 1-D bulk async copies
// (cp.async.bulk / UBLKCP) as a function of copy size, stages in flight and CTAs per SM?  No compute.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_stream_bench tma_stream_bench.cu && ./tma_stream_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_stream(const unsigned char *src, size_t total, int chunk, int S, int ncopies, int hint, unsigned long long *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar[16];
    const int tid = threadIdx.x;
    uint64_t policy = 0;
    if (tid == 0) {
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        for (int s = 0; s < S; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nchunks = total / chunk;
    const size_t first = blockIdx.x, stride = gridDim.x;
    const size_t nloc = first < nchunks ? (nchunks - first + stride - 1) / stride : 0;
    auto issue = [&](size_t i) {
        const int s = (int)(i % S);
        const unsigned char *g = src + (first + i * stride) * (size_t)chunk;
        unsigned char *d = smem + (size_t)s * chunk;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(chunk) : "memory");
        const int part = chunk / ncopies;
        for (int c = 0; c < ncopies; c++) {
            if (hint)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                             ::"r"(smem_u32(d + c * part)), "l"(g + c * part), "r"(part), "r"(smem_u32(&bar[s])), "l"(policy) : "memory");
            else
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(d + c * part)), "l"(g + c * part), "r"(part), "r"(smem_u32(&bar[s])) : "memory");
        }
    };
    if (tid == 0)
        for (size_t i = 0; i < (size_t)S && i < nloc; i++) issue(i);
    unsigned long long acc = 0;
    for (size_t i = 0; i < nloc; i++) {
        const int s = (int)(i % S);
        const uint32_t parity = (uint32_t)((i / S) & 1);
        asm volatile(
            "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
            ::"r"(smem_u32(&bar[s])), "r"(parity) : "memory");
        acc += reinterpret_cast<const unsigned long long *>(smem + (size_t)s * chunk)[tid];   // touch the data
        __syncthreads();
        if (tid == 0 && i + S < nloc) issue(i + S);
    }
    if (acc == 0x1234567ull) *sink = acc;
}

__global__ void k_stream_timed(const unsigned char *src, size_t total, int chunk, int S, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar[16];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < S; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nchunks = total / chunk;
    const size_t first = blockIdx.x, stride = gridDim.x;
    const size_t nloc = first < nchunks ? (nchunks - first + stride - 1) / stride : 0;
    long long t_wait = 0, t_sync = 0, t_exp = 0, t_cp = 0, t_lat = 0;
    long long issued_at[16];
    auto issue = [&](size_t i) {
        const int s = (int)(i % S);
        const unsigned char *g = src + (first + i * stride) * (size_t)chunk;
        unsigned char *d = smem + (size_t)s * chunk;
        long long a = clock64();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(chunk) : "memory");
        long long b = clock64();
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(d)), "l"(g), "r"(chunk), "r"(smem_u32(&bar[s])) : "memory");
        long long c = clock64();
        t_exp += b - a; t_cp += c - b; issued_at[s] = c;
    };
    if (tid == 0)
        for (size_t i = 0; i < (size_t)S && i < nloc; i++) issue(i);
    unsigned long long acc = 0;
    for (size_t i = 0; i < nloc; i++) {
        const int s = (int)(i % S);
        const uint32_t parity = (uint32_t)((i / S) & 1);
        long long a = clock64();
        asm volatile(
            "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
            ::"r"(smem_u32(&bar[s])), "r"(parity) : "memory");
        long long b = clock64();
        acc += reinterpret_cast<const unsigned long long *>(smem + (size_t)s * chunk)[tid];
        __syncthreads();
        long long c = clock64();
        if (tid == 0) { t_wait += b - a; t_sync += c - b; t_lat += b - issued_at[s]; }
        if (tid == 0 && i + S < nloc) issue(i + S);
    }
    if (tid == 0 && blockIdx.x == 0) { out[0] = t_wait; out[1] = t_sync; out[2] = t_exp; out[3] = t_cp; out[4] = t_lat; out[5] = (long long)nloc; out[6] = (long long)acc; }
}

int main() {
    const size_t total = (size_t)8 << 30;
    unsigned char *d;
    unsigned long long *sink;
    cudaMalloc(&d, total);
    cudaMalloc(&sink, 8);
    cudaMemset(d, 1, total);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int cfgs[][5] = {  // chunk bytes, stages, ncopies, threads, hint
        {16384, 2, 1, 128, 1}, {16384, 4, 1, 128, 1}, {16384, 8, 1, 128, 1}, {16384, 12, 1, 128, 1},
        {32768, 2, 1, 128, 1}, {32768, 3, 1, 128, 1}, {32768, 6, 1, 128, 1}, {65536, 3, 1, 128, 1},
        {8192, 8, 1, 128, 1},  {8192, 16, 1, 128, 1}, {16384, 8, 4, 128, 1}, {16384, 8, 1, 128, 0},
        {16384, 4, 1, 256, 1}, {4096, 16, 1, 128, 1}, {16384, 3, 1, 128, 1}, {12288, 5, 3, 256, 1}};
    for (auto &c : cfgs) {
        const int chunk = c[0], S = c[1], nc = c[2], nt = c[3], hint = c[4];
        const size_t smem = (size_t)chunk * S;
        cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stream, nt, smem);
        if (occ < 1) { printf("chunk %d S %d: does not fit\n", chunk, S); continue; }
        for (int ctas = 1; ctas <= occ && ctas <= 4; ctas++) {
            const int grid = 148 * ctas;
            k_stream<<<grid, nt, smem>>>(d, total, chunk, S, nc, hint, sink);
            cudaEventRecord(e0);
            k_stream<<<grid, nt, smem>>>(d, total, chunk, S, nc, hint, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("chunk %6d  stages %2d  copies/chunk %d  threads %3d  hint %d  CTAs/SM %d  in-flight/SM %4zu KB : %7.1f GB/s  (%s)\n",
                   chunk, S, nc, nt, hint, ctas, smem * ctas / 1024, total / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
        }
    }
    {
        long long *dout, hout[8];
        cudaMalloc(&dout, 64);
        for (int S : {2, 4, 8}) {
            const int chunk = 16384;
            const size_t smem = (size_t)chunk * S;
            cudaFuncSetAttribute(k_stream_timed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_stream_timed<<<148, 128, smem>>>(d, total, chunk, S, dout);
            cudaMemcpy(hout, dout, 56, cudaMemcpyDeviceToHost);
            double n = (double)hout[5];
            printf("timed chunk %d S %d (1 CTA/SM): per iteration cycles: wait %.0f  touch+sync %.0f  expect_tx %.0f  ublkcp %.0f  issue->landed-seen %.0f  (%s)\n",
                   chunk, S, hout[0] / n, hout[1] / n, hout[2] / n, hout[3] / n, hout[4] / n, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
