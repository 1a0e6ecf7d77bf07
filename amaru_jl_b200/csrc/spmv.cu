// K4/K5/K8: block-CSR SpMV  y = A x  (+ fused p·Ap partial dot), the dominant kernel of the PCG that replaces
// lu(K11) (reference src/solver.jl:38-43) and of the K12/K21/K22 products (solver.jl:32,57).
//
// k_spmv_stream (default) — TMA-streamed, tile-staged, warp-specialised:
//   The matrix is ~97 % of the bytes of a CG iteration and is read exactly once, so it is streamed with bulk async
//   copies (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier) instead of per-lane loads.  On the host the block
//   rows are grouped into tiles of whole consecutive rows (<= tile_blks blocks, <= tile_rows rows); block rows are
//   contiguous in block-CSR, so a tile is ONE contiguous byte range of the value array.  Per tile the host also
//   builds one packed record: one int per row (local block offset | prescribed-dof mask << 16 | local index of the
//   diagonal << 19), the sorted list of unique column nodes `ucol`, and a 16-bit tile-local column index per block.
//   Each CTA = 1 producer warp + NCW consumer warps.  The producer streams the values of `nstages` tiles and the
//   records of `nstages + xd` tiles through two shared-memory rings (two bulk copies per tile with an L2 evict_first
//   policy so that x stays L2-resident; tile headers are prefetched 32 tiles ahead), paced by empty/full mbarriers.
//   The consumers start the gather of the x entries of tile i+xd with cp.async (8-byte copies, nothing held in
//   registers) and contract tile i entirely out of shared memory: lane = (block-in-step, row), so one warp step
//   covers 10 (3x3) / 16 (2x2) blocks.  No global load sits on the critical path of the inner loop; bytes in flight
//   are decoupled from registers and occupancy.
//   (profiles/tma_stream_bench.cu: pure bulk-copy streaming reads 7.3 TB/s on this B200; profiles/README.md: the
//   kernel's DRAM traffic equals its algorithmic bytes, what is left is consumer instruction issue.)
// k_spmv (fallback when a single row exceeds a tile, or AMARU_SPMV_SIMPLE=1) — one warp per block row, per-lane loads.
//
// Algorithmic bytes per launch: nblk*(8*bs^2 + 2) + 4*(rows + 1 + unique columns) per tile + 32 per tile
//   + 8*n (x once, L2-resident) + 8*n (y).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "reduce.cuh"

namespace {

constexpr int ROW_THREADS = 256;
constexpr int MAX_STAGES = 8;
constexpr int NCW = 4;                      // consumer warps per CTA (8 with 2x larger tiles measured the same: 3.11 ms)
constexpr int XD_MAX = 4;                   // x gathers are started xd (<= XD_MAX) tiles ahead of the contraction

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// producer-side wait: sleeps between probes so that the spinning warp does not steal issue slots from the consumer
// warp that shares its scheduler (ncu: the bare try_wait loop executed 4x more instructions than the contraction)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity, unsigned sleep_ns) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(sleep_ns);
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
    switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    }
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory"); }

struct SpmvTile {   // rows [r0, r0+nrows), blocks [b0, b0+nb), record at trec[moff], nu unique columns
    int32_t r0, nrows, b0, nb, moff, nu, recints, pad;
};

__host__ __device__ inline size_t r128(size_t b) { return (b + 127) / 128 * 128; }
struct StageLayout {
    size_t vbytes, rbytes, xbytes, sbytes;
};
__host__ __device__ inline StageLayout stage_layout(int bs, int tile_blks, int tile_rows, int xcap) {
    StageLayout L;
    L.vbytes = r128((size_t)tile_blks * bs * bs * 8 + 16);
    L.rbytes = r128((size_t)(tile_rows + 1 + xcap + (tile_blks + 1) / 2) * 4 + 16);
    L.xbytes = r128((size_t)xcap * bs * 8);
    L.sbytes = L.vbytes + L.rbytes + L.xbytes;   // one value stage + one record stage
    return L;
}

// reduce the per-lane slot accumulators of one block row; result valid in the writer lanes
template <int BS>
__device__ __forceinline__ double row_reduce(double acc) {
    if constexpr (BS == 3) {
        const double t1 = __shfl_down_sync(0xffffffffu, acc, 9), t2 = __shfl_down_sync(0xffffffffu, acc, 18);
        const double sm = acc + t1 + t2;                       // lanes 0..8: slot rc summed over the 3 block groups
        const double u1 = __shfl_down_sync(0xffffffffu, sm, 1), u2 = __shfl_down_sync(0xffffffffu, sm, 2);
        return sm + u1 + u2;                                   // lanes 0,3,6: row r = lane/3
    } else {
        double sm = acc;
        sm += __shfl_xor_sync(0xffffffffu, sm, 4);
        sm += __shfl_xor_sync(0xffffffffu, sm, 8);
        sm += __shfl_xor_sync(0xffffffffu, sm, 16);
        return sm + __shfl_xor_sync(0xffffffffu, sm, 1);       // lanes 0 and 2
    }
}

// ------------------------------------------------------------------------------------------------ streamed kernel
// Shared memory: a ring of SV value stages (the big ones) and a deeper ring of SR = SV + XD record stages
// (record + staged x, small), so that the x gathers can run XD tiles ahead without holding value buffers hostage:
// up to SV-1 value tiles are in TMA flight per CTA.
// (the first version of this kernel, whose consumer warps gathered x themselves, is in the history: 3.12 ms against 2.90 ms)

// The x gather has its own warp and nothing synchronises the consumer warps with each other:
//   warp NCW   : TMA producer (values ring SV, record ring SR): per tile two cp.async.bulk copies completing on `full` mbarriers;
//   warp NCW+1 : gather warp — waits for a record, issues the cp.async copies of the tile's unique x entries and lets
//                the copies themselves arrive on the stage's `xfull` mbarrier (cp.async.mbarrier.arrive.noinc, one
//                arrival per lane), then moves on: it runs ahead as far as records have landed;
//   warps 0..NCW-1 : consumers — wait xfull + full_v of their tile, contract their rows, release the two stages.
// No bar.sync, no cp.async.wait_group: a consumer is never held back by a slower sibling, and the gather bookkeeping
// is executed once per tile instead of once per consumer warp.
constexpr int STREAM2_THREADS = (NCW + 2) * 32;

template <int BS, bool DOT>
__global__ void __launch_bounds__(STREAM2_THREADS)
k_spmv_stream2(int ntiles, const SpmvTile *__restrict__ tiles, const int32_t *__restrict__ trec,
               const double *__restrict__ A, const double *__restrict__ x, double *__restrict__ y, int mask_rows,
               int tile_blks, int tile_rows, int xcap, int nstages, int rextra, int sleep_ns, double *partial,
               CgScalars *scal, int check_done, int finalize) {
    if (check_done && scal->done) return;
    constexpr int B2 = BS * BS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_v[MAX_STAGES], empty_v[MAX_STAGES], full_r[MAX_STAGES + XD_MAX], empty_r[MAX_STAGES + XD_MAX],
        xfull[MAX_STAGES + XD_MAX];
    __shared__ SpmvTile shdr[MAX_STAGES + XD_MAX];
    const StageLayout L = stage_layout(BS, tile_blks, tile_rows, xcap);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int SV = nstages, SR = nstages + rextra;
    unsigned char *vring = smem_raw;
    unsigned char *rring = smem_raw + (size_t)SV * L.vbytes;
    const size_t rstride = L.rbytes + L.xbytes;
    if (tid == 0) {
        for (int s = 0; s < SV; s++) {
            mbar_init(&full_v[s], 1);
            mbar_init(&empty_v[s], NCW);
        }
        for (int s = 0; s < SR; s++) {
            mbar_init(&full_r[s], 1);
            mbar_init(&empty_r[s], NCW);
            mbar_init(&xfull[s], 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int first = blockIdx.x, stride = gridDim.x;
    const int nloc = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
    double dsum[1] = {0.0};

    if (warp == NCW) {
        // ============================ TMA producer: record of tile j as soon as its record stage is free, values of tile
        // j - rextra as soon as its value stage is free (records run ahead of values by `rextra` tiles)
        uint64_t policy = 0;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        SpmvTile hcur{}, hnext{};
        if (lane < nloc) hcur = tiles[first + lane * stride];
        if (32 + lane < nloc) hnext = tiles[first + (32 + lane) * stride];
        int vb0 = 0, vnb = 0;
        int psr = 0, psv = 0;
        uint32_t prph = 1u, pvph = 1u;
        for (int j = 0; j < nloc + rextra; j++) {
            if (j < nloc) {
                if ((j & 31) == 0 && j > 0) {
                    hcur = hnext;
                    if (j + 32 + lane < nloc) hnext = tiles[first + (j + 32 + lane) * stride];
                }
                SpmvTile ti;
                ti.r0 = __shfl_sync(0xffffffffu, hcur.r0, j & 31);
                ti.nrows = __shfl_sync(0xffffffffu, hcur.nrows, j & 31);
                ti.b0 = __shfl_sync(0xffffffffu, hcur.b0, j & 31);
                ti.nb = __shfl_sync(0xffffffffu, hcur.nb, j & 31);
                ti.moff = __shfl_sync(0xffffffffu, hcur.moff, j & 31);
                ti.nu = __shfl_sync(0xffffffffu, hcur.nu, j & 31);
                ti.recints = __shfl_sync(0xffffffffu, hcur.recints, j & 31);
                ti.pad = 0;
                if (lane == (j & 31)) {
                    vb0 = ti.b0;
                    vnb = ti.nb;
                }
                if (lane == 0) {
                    const int sr = psr;
                    if (j >= SR) mbar_wait_backoff(&empty_r[sr], prph, (unsigned)sleep_ns);
                    shdr[sr] = ti;
                    const uint64_t rs = ((uint64_t)ti.recints * 4 + 15) & ~15ull;
                    mbar_arrive_expect_tx(&full_r[sr], (uint32_t)rs);
                    tma_bulk_g2s(rring + (size_t)sr * rstride, reinterpret_cast<const unsigned char *>(trec + ti.moff),
                                 (uint32_t)rs, &full_r[sr], policy);
                }
                if (++psr == SR) {
                    psr = 0;
                    prph ^= 1u;
                }
            }
            const int jv = j - rextra;
            const int b0 = __shfl_sync(0xffffffffu, vb0, jv & 31), nb = __shfl_sync(0xffffffffu, vnb, jv & 31);
            if (jv >= 0 && lane == 0) {
                const int sv = psv;
                if (jv >= SV) mbar_wait_backoff(&empty_v[sv], pvph, (unsigned)sleep_ns);
                const uint64_t v0 = (uint64_t)b0 * (B2 * 8), va = v0 & ~15ull;
                const uint64_t vs = ((v0 + (uint64_t)nb * (B2 * 8) - va) + 15) & ~15ull;
                mbar_arrive_expect_tx(&full_v[sv], (uint32_t)vs);
                tma_bulk_g2s(vring + (size_t)sv * L.vbytes, reinterpret_cast<const unsigned char *>(A) + va, (uint32_t)vs,
                             &full_v[sv], policy);
            }
            if (jv >= 0 && ++psv == SV) {
                psv = 0;
                pvph ^= 1u;
            }
            __syncwarp();
        }
    } else if (warp == NCW + 1) {
        // ============================ gather warp: x entries of every tile -> its record stage (cp.async, 8 B per copy)
        int gs = 0;
        uint32_t gph = 0;
        for (int j = 0; j < nloc; j++) {
            mbar_wait(&full_r[gs], gph);   // record landed (and, transitively, the stage was released by the consumers)
            const int nrows_j = shdr[gs].nrows, nu_j = shdr[gs].nu;
            unsigned char *bj = rring + (size_t)gs * rstride;
            const int32_t *ucol = reinterpret_cast<const int32_t *>(bj) + nrows_j + 1;
            const uint32_t sxa = smem_u32(bj + L.rbytes);
            for (int k = lane; k < nu_j; k += 32) {
                const double *src = x + (int64_t)ucol[k] * BS;
                const uint32_t dst = sxa + (uint32_t)k * (BS * 8u);
#pragma unroll
                for (int d = 0; d < BS; d++)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + d * 8u), "l"(src + d) : "memory");
            }
            // the lane's copies arrive on xfull[gs] when they have landed (32 arrivals complete the phase)
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&xfull[gs])) : "memory");
            if (++gs == SR) {
                gs = 0;
                gph ^= 1u;
            }
        }
    } else {
        // ============================ consumer warps (independent of each other)
        constexpr int BPS = 32 / BS;
        const int b = lane / BS, r = lane - b * BS;
        const bool act = lane < BPS * BS;
        const bool writer = lane < BS;
        int sr = 0, sv = 0;
        uint32_t rph = 0, vph = 0;
        for (int i = 0; i < nloc; i++) {
            mbar_wait(&full_r[sr], rph);   // record + header (already complete by the time x is staged; cheap acquire)
            mbar_wait(&xfull[sr], rph);    // x staged
            const int h_r0 = shdr[sr].r0, h_nrows = shdr[sr].nrows, h_b0 = shdr[sr].b0, h_nu = shdr[sr].nu;
            const unsigned char *rb = rring + (size_t)sr * rstride;
            const int32_t *rec = reinterpret_cast<const int32_t *>(rb);
            const uint16_t *sl = reinterpret_cast<const uint16_t *>(rec + h_nrows + 1 + h_nu);
            const double *sx = reinterpret_cast<const double *>(rb + L.rbytes);
            mbar_wait(&full_v[sv], vph);
            const double *sval = reinterpret_cast<const double *>(vring + (size_t)sv * L.vbytes + (((uint64_t)h_b0 * (B2 * 8)) & 15ull));
            for (int lr = warp; lr < h_nrows; lr += NCW) {
                const int32_t e0 = rec[lr], e1 = rec[lr + 1];
                const int k0 = e0 & 0xffff, nbr = (e1 & 0xffff) - k0;
                const double *pv = sval + (k0 + b) * B2 + r * BS;
                const uint16_t *pl = sl + k0 + b;
                const int nfull = nbr / BPS, rem = nbr - nfull * BPS;
                double acc0 = 0.0, acc1 = 0.0;
                if (act) {
                    int s = 0;
                    for (; s + 2 <= nfull; s += 2) {
                        const double *x0 = sx + pl[0], *x1 = sx + pl[BPS];
#pragma unroll
                        for (int j = 0; j < BS; j++) {
                            acc0 += pv[j] * x0[j];
                            acc1 += pv[BPS * B2 + j] * x1[j];
                        }
                        pl += 2 * BPS;
                        pv += 2 * BPS * B2;
                    }
                    if (s < nfull) {
                        const double *x0 = sx + pl[0];
#pragma unroll
                        for (int j = 0; j < BS; j++) acc0 += pv[j] * x0[j];
                        pl += BPS;
                        pv += BPS * B2;
                    }
                    if (b < rem) {
                        const double *x0 = sx + pl[0];
#pragma unroll
                        for (int j = 0; j < BS; j++) acc1 += pv[j] * x0[j];
                    }
                }
                double tot = acc0 + acc1;
                if constexpr (BS == 3) {
                    const double s1 = tot + __shfl_down_sync(0xffffffffu, tot, 15);
                    const double u = s1 + __shfl_down_sync(0xffffffffu, s1, 3);
                    const double v = u + __shfl_down_sync(0xffffffffu, u, 6);
                    tot = v + __shfl_down_sync(0xffffffffu, s1, 12);
                } else {
                    tot += __shfl_xor_sync(0xffffffffu, tot, 2);
                    tot += __shfl_xor_sync(0xffffffffu, tot, 4);
                    tot += __shfl_xor_sync(0xffffffffu, tot, 8);
                    tot += __shfl_xor_sync(0xffffffffu, tot, 16);
                }
                if (writer) {
                    const int64_t idx = (int64_t)(h_r0 + lr) * BS + lane;
                    if (mask_rows && ((e0 >> (16 + lane)) & 1)) tot = 0.0;
                    y[idx] = tot;
                    if (DOT && nbr > 0) dsum[0] += tot * sx[((e0 >> 19) & 0x1fff) * BS + lane];
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&empty_v[sv]);
                mbar_arrive(&empty_r[sr]);
            }
            if (++sr == SR) {
                sr = 0;
                rph ^= 1u;
            }
            if (++sv == SV) {
                sv = 0;
                vph ^= 1u;
            }
        }
    }
    if (DOT) spmv_dot_epilogue<STREAM2_THREADS>(dsum, partial, scal, finalize);
}


// ------------------------------------------------------------------------------------------------ symmetric storage
// K (and a*K + b*M) is symmetric for the three materials of the path (associated flow), so the CG product needs only the
// blocks with column >= row: HALF the bytes of the dominant kernel.  Stored: for every owned row the blocks (i, j >= i) in
// local numbering — ghosts are numbered after the owned nodes, so every owned x ghost block is stored too, and the
// transposed contribution of those (and of the diagonal blocks) is skipped: the neighbour rank owns that row and stores the
// mirrored block itself.  y_i += A_ij x_j is contracted as in k_spmv_stream2; the transposed part y_j += A_ij' x_i is
// accumulated per consumer warp in a PRIVATE shared-memory image of the tile's unique columns (no shared atomics: inside a
// row the block columns are distinct, rows of a warp are separated by __syncwarp), and a seventh warp — the flush warp —
// sums the four images once all consumers are done with the tile and sends ONE red.global.add.f64 per unique column
// entry (0.14 reductions per stored value at HEX20) to y, which the host zeroed before the launch.  Prescribed dofs are
// masked at the flush (3 mask bits ride in the unique-column list).
//   warps 0..NCW-1 consumers | NCW TMA producer | NCW+1 x-gather | NCW+2 flush
// The p.Ap dot is accumulated per lane with weight 2 on the blocks whose mirror image is implied and 1 on the others
// (diagonal, owned x ghost), so it is still one fused, deterministic reduction.  The y accumulation order through the L2
// reductions varies run to run (last-bit differences in y).  OPT-IN (AMARU_SPMV_SYM=1): on B200 it does not beat the
// full-storage kernel, see setup_sym; a variant that forms the transposed products with shuffles instead of a second
// pass over the values in shared memory measured slower still (3.65 ms).
constexpr int SYM_THREADS = (NCW + 3) * 32;
constexpr int MAX_YST = 4;
constexpr uint32_t UCOL_NODE = 0x0fffffffu;   // ucol entry: node | fixed-dof mask << 28 | ghost << 31
constexpr int LCOL_SKIP = 0x8000;             // lcol entry: (local column)*bs | skip-transposed flag

template <int BS, bool DOT>
__global__ void __launch_bounds__(SYM_THREADS)
k_spmv_sym(int ntiles, const SpmvTile *__restrict__ tiles, const int32_t *__restrict__ trec, const double *__restrict__ A,
           const double *__restrict__ x, double *__restrict__ y, int mask_rows, int tile_blks, int tile_rows, int xcap,
           int nstages, int rextra, int ystages, int sleep_ns, double *partial, CgScalars *scal, int check_done, int finalize) {
    if (check_done && scal->done) return;
    constexpr int B2 = BS * BS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_v[MAX_STAGES], empty_v[MAX_STAGES], full_r[MAX_STAGES + XD_MAX], empty_r[MAX_STAGES + XD_MAX],
        xfull[MAX_STAGES + XD_MAX], ydone[MAX_YST], empty_y[MAX_YST];
    __shared__ SpmvTile shdr[MAX_STAGES + XD_MAX];
    const StageLayout L = stage_layout(BS, tile_blks, tile_rows, xcap);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int SV = nstages, SR = nstages + rextra, SY = ystages;
    unsigned char *vring = smem_raw;
    unsigned char *rring = smem_raw + (size_t)SV * L.vbytes;
    const size_t rstride = L.rbytes + L.xbytes;
    double *yring = reinterpret_cast<double *>(rring + (size_t)SR * rstride);   // [SY][NCW][xcap*BS]
    const int ylen = (int)(L.xbytes / 8);                                       // doubles per image (>= xcap*BS)
    if (tid == 0) {
        for (int s = 0; s < SV; s++) {
            mbar_init(&full_v[s], 1);
            mbar_init(&empty_v[s], NCW);
        }
        for (int s = 0; s < SR; s++) {
            mbar_init(&full_r[s], 1);
            mbar_init(&empty_r[s], NCW + 1);   // consumers + flush warp
            mbar_init(&xfull[s], 32);
        }
        for (int s = 0; s < SY; s++) {
            mbar_init(&ydone[s], NCW);
            mbar_init(&empty_y[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < SY * NCW * ylen; i += SYM_THREADS) yring[i] = 0.0;
    __syncthreads();
    const int first = blockIdx.x, stride = gridDim.x;
    const int nloc = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
    double dsum[1] = {0.0};

    if (warp == NCW) {
        // ============================ TMA producer (as in k_spmv_stream2)
        uint64_t policy = 0;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        SpmvTile hcur{}, hnext{};
        if (lane < nloc) hcur = tiles[first + lane * stride];
        if (32 + lane < nloc) hnext = tiles[first + (32 + lane) * stride];
        int vb0 = 0, vnb = 0;
        int psr = 0, psv = 0;
        uint32_t prph = 1u, pvph = 1u;
        for (int j = 0; j < nloc + rextra; j++) {
            if (j < nloc) {
                if ((j & 31) == 0 && j > 0) {
                    hcur = hnext;
                    if (j + 32 + lane < nloc) hnext = tiles[first + (j + 32 + lane) * stride];
                }
                SpmvTile ti;
                ti.r0 = __shfl_sync(0xffffffffu, hcur.r0, j & 31);
                ti.nrows = __shfl_sync(0xffffffffu, hcur.nrows, j & 31);
                ti.b0 = __shfl_sync(0xffffffffu, hcur.b0, j & 31);
                ti.nb = __shfl_sync(0xffffffffu, hcur.nb, j & 31);
                ti.moff = __shfl_sync(0xffffffffu, hcur.moff, j & 31);
                ti.nu = __shfl_sync(0xffffffffu, hcur.nu, j & 31);
                ti.recints = __shfl_sync(0xffffffffu, hcur.recints, j & 31);
                ti.pad = 0;
                if (lane == (j & 31)) {
                    vb0 = ti.b0;
                    vnb = ti.nb;
                }
                if (lane == 0) {
                    const int sr = psr;
                    if (j >= SR) mbar_wait_backoff(&empty_r[sr], prph, (unsigned)sleep_ns);
                    shdr[sr] = ti;
                    const uint64_t rs = ((uint64_t)ti.recints * 4 + 15) & ~15ull;
                    mbar_arrive_expect_tx(&full_r[sr], (uint32_t)rs);
                    tma_bulk_g2s(rring + (size_t)sr * rstride, reinterpret_cast<const unsigned char *>(trec + ti.moff),
                                 (uint32_t)rs, &full_r[sr], policy);
                }
                if (++psr == SR) {
                    psr = 0;
                    prph ^= 1u;
                }
            }
            const int jv = j - rextra;
            const int b0 = __shfl_sync(0xffffffffu, vb0, jv & 31), nb = __shfl_sync(0xffffffffu, vnb, jv & 31);
            if (jv >= 0 && lane == 0) {
                const int sv = psv;
                if (jv >= SV) mbar_wait_backoff(&empty_v[sv], pvph, (unsigned)sleep_ns);
                const uint64_t v0 = (uint64_t)b0 * (B2 * 8), va = v0 & ~15ull;
                const uint64_t vs = ((v0 + (uint64_t)nb * (B2 * 8) - va) + 15) & ~15ull;
                mbar_arrive_expect_tx(&full_v[sv], (uint32_t)vs);
                tma_bulk_g2s(vring + (size_t)sv * L.vbytes, reinterpret_cast<const unsigned char *>(A) + va, (uint32_t)vs,
                             &full_v[sv], policy);
            }
            if (jv >= 0 && ++psv == SV) {
                psv = 0;
                pvph ^= 1u;
            }
            __syncwarp();
        }
    } else if (warp == NCW + 1) {
        // ============================ gather warp: x entries of the tile's unique columns -> record stage
        int gs = 0;
        uint32_t gph = 0;
        for (int j = 0; j < nloc; j++) {
            mbar_wait(&full_r[gs], gph);
            const int nrows_j = shdr[gs].nrows, nu_j = shdr[gs].nu;
            unsigned char *bj = rring + (size_t)gs * rstride;
            const uint32_t *ucol = reinterpret_cast<const uint32_t *>(bj) + nrows_j + 1;
            const uint32_t sxa = smem_u32(bj + L.rbytes);
            for (int k = lane; k < nu_j; k += 32) {
                const double *src = x + (int64_t)(ucol[k] & UCOL_NODE) * BS;
                const uint32_t dst = sxa + (uint32_t)k * (BS * 8u);
#pragma unroll
                for (int d = 0; d < BS; d++)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + d * 8u), "l"(src + d) : "memory");
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&xfull[gs])) : "memory");
            if (++gs == SR) {
                gs = 0;
                gph ^= 1u;
            }
        }
    } else if (warp == NCW + 2) {
        // ============================ flush warp: sum the consumers' column images of a finished tile, reduce into y
        int sr = 0, sy = 0;
        uint32_t rph = 0, dph = 0;
        for (int i = 0; i < nloc; i++) {
            mbar_wait(&full_r[sr], rph);
            mbar_wait_backoff(&ydone[sy], dph, 40u);
            const int h_nrows = shdr[sr].nrows, h_nu = shdr[sr].nu;
            const uint32_t *ucol = reinterpret_cast<const uint32_t *>(rring + (size_t)sr * rstride) + h_nrows + 1;
            double *img = yring + (size_t)sy * NCW * ylen;
            for (int u = lane; u < h_nu; u += 32) {
                const uint32_t uc = ucol[u];
                double s[BS];
#pragma unroll
                for (int c = 0; c < BS; c++) {
                    double t = 0.0;
#pragma unroll
                    for (int w = 0; w < NCW; w++) {
                        t += img[w * ylen + u * BS + c];
                        img[w * ylen + u * BS + c] = 0.0;
                    }
                    s[c] = t;
                }
                if (!(uc >> 31)) {
                    double *dst = y + (int64_t)(uc & UCOL_NODE) * BS;
#pragma unroll
                    for (int c = 0; c < BS; c++)
                        if (s[c] != 0.0 && !(mask_rows && ((uc >> (28 + c)) & 1u))) atomicAdd(dst + c, s[c]);
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&empty_y[sy]);
                mbar_arrive(&empty_r[sr]);
            }
            if (++sr == SR) {
                sr = 0;
                rph ^= 1u;
            }
            if (++sy == SY) {
                sy = 0;
                dph ^= 1u;
            }
        }
    } else {
        // ============================ consumer warps
        constexpr int BPS = 32 / BS;
        const int b = lane / BS, r = lane - b * BS;
        const bool act = lane < BPS * BS;
        const bool writer = lane < BS;
        int sr = 0, sv = 0, sy = 0;
        uint32_t rph = 0, vph = 0, yph = 1u;
        for (int i = 0; i < nloc; i++) {
            mbar_wait(&full_r[sr], rph);
            mbar_wait(&xfull[sr], rph);
            const int h_nrows = shdr[sr].nrows, h_b0 = shdr[sr].b0, h_nu = shdr[sr].nu;
            const unsigned char *rb = rring + (size_t)sr * rstride;
            const int32_t *rec = reinterpret_cast<const int32_t *>(rb);
            const uint16_t *sl = reinterpret_cast<const uint16_t *>(rec + h_nrows + 1 + h_nu);
            const double *sx = reinterpret_cast<const double *>(rb + L.rbytes);
            mbar_wait(&full_v[sv], vph);
            if (i >= SY) mbar_wait(&empty_y[sy], yph);   // the flush warp has emptied (and zeroed) this image
            double *img = yring + ((size_t)sy * NCW + warp) * ylen;
            const double *sval = reinterpret_cast<const double *>(vring + (size_t)sv * L.vbytes + (((uint64_t)h_b0 * (B2 * 8)) & 15ull));
            for (int lr = warp; lr < h_nrows; lr += NCW) {
                const int32_t e0 = rec[lr], e1 = rec[lr + 1];
                const int k0 = e0 & 0xffff, nbr = (e1 & 0xffff) - k0;
                const int dl = ((e0 >> 19) & 0x1fff) * BS;
                double xi[BS];   // x of the row node: the transposed products need all of it, the dot its component r
#pragma unroll
                for (int j = 0; j < BS; j++) xi[j] = sx[dl + j];
                const double xir = act ? sx[dl + r] : 0.0;
                double acc = 0.0;
                if (act) {
                    for (int k = b; k < nbr; k += BPS) {
                        const int lc = sl[k0 + k];
                        const int off = lc & (LCOL_SKIP - 1);
                        const double *pv = sval + (k0 + k) * B2;
                        const double *xj = sx + off;
                        double part = 0.0;
#pragma unroll
                        for (int j = 0; j < BS; j++) part += pv[r * BS + j] * xj[j];
                        acc += part;
                        if (lc & LCOL_SKIP) {
                            if (DOT) dsum[0] += xir * part;
                        } else {
                            if (DOT) dsum[0] += 2.0 * (xir * part);
                            double t = 0.0;   // lane (b, c = r): column c of the block times x_i
#pragma unroll
                            for (int j = 0; j < BS; j++) t += pv[j * BS + r] * xi[j];
                            img[off + r] += t;
                        }
                    }
                }
                double tot = acc;
                if constexpr (BS == 3) {
                    const double s1 = tot + __shfl_down_sync(0xffffffffu, tot, 15);
                    const double u = s1 + __shfl_down_sync(0xffffffffu, s1, 3);
                    const double v = u + __shfl_down_sync(0xffffffffu, u, 6);
                    tot = v + __shfl_down_sync(0xffffffffu, s1, 12);
                } else {
                    tot += __shfl_xor_sync(0xffffffffu, tot, 2);
                    tot += __shfl_xor_sync(0xffffffffu, tot, 4);
                    tot += __shfl_xor_sync(0xffffffffu, tot, 8);
                    tot += __shfl_xor_sync(0xffffffffu, tot, 16);
                }
                if (writer) img[dl + lane] += tot;   // the row node is one of the tile's unique columns (diagonal block)
                __syncwarp();                        // the next row of this warp may hit the same columns
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&empty_v[sv]);
                mbar_arrive(&empty_r[sr]);
                mbar_arrive(&ydone[sy]);
            }
            if (++sr == SR) {
                sr = 0;
                rph ^= 1u;
            }
            if (++sv == SV) {
                sv = 0;
                vph ^= 1u;
            }
            if (++sy == SY) {
                sy = 0;
                yph ^= 1u;
            }
        }
    }
    if (DOT) spmv_dot_epilogue<SYM_THREADS>(dsum, partial, scal, finalize);
}

__global__ void k_extract_upper(int64_t nublk, int b2, const int32_t *__restrict__ usrc, const double *__restrict__ A,
                                double *__restrict__ Asym) {
    const int64_t n = nublk * b2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i / b2;
        Asym[i] = A[(int64_t)usrc[k] * b2 + (i - k * b2)];
    }
}

// ------------------------------------------------------------------------------------------------ per-lane fallback
template <int BS, bool DOT>
__global__ void __launch_bounds__(ROW_THREADS)
k_spmv(int64_t nrows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
       const double *__restrict__ A, const double *__restrict__ x, double *__restrict__ y,
       const uint8_t *__restrict__ fixed, int mask_rows, double *partial, CgScalars *scal, int check_done,
       int finalize) {
    if (check_done && scal->done) return;
    constexpr int B2 = BS * BS, BPI = 32 / B2, ACT = BPI * B2;
    const int lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (ROW_THREADS / 32);
    const int bl = lane / B2, rc = lane - bl * B2;
    const int c = rc % BS;
    const bool act = lane < ACT;
    const bool writer = (BS == 3) ? (lane < 9 && lane % 3 == 0) : (lane == 0 || lane == 2);
    const int wr = (BS == 3) ? lane / 3 : lane / 2;
    double dsum[1] = {0.0};
    for (int64_t row = gw; row < nrows; row += nw) {
        const int32_t s = rowptr[row], e = rowptr[row + 1];
        double acc = 0.0;
        if (act) {
            int32_t k = s + bl;
            for (; k + 3 * BPI < e; k += 4 * BPI) {
                const int32_t c0 = __ldg(col + k), c1 = __ldg(col + k + BPI), c2 = __ldg(col + k + 2 * BPI),
                              c3 = __ldg(col + k + 3 * BPI);
                const double v0 = __ldg(A + (int64_t)k * B2 + rc), v1 = __ldg(A + (int64_t)(k + BPI) * B2 + rc),
                             v2 = __ldg(A + (int64_t)(k + 2 * BPI) * B2 + rc),
                             v3 = __ldg(A + (int64_t)(k + 3 * BPI) * B2 + rc);
                acc += v0 * x[(int64_t)c0 * BS + c];
                acc += v1 * x[(int64_t)c1 * BS + c];
                acc += v2 * x[(int64_t)c2 * BS + c];
                acc += v3 * x[(int64_t)c3 * BS + c];
            }
            for (; k < e; k += BPI) acc += __ldg(A + (int64_t)k * B2 + rc) * x[(int64_t)__ldg(col + k) * BS + c];
        }
        double tot = row_reduce<BS>(acc);
        if (writer) {
            const int64_t i = row * BS + wr;
            if (mask_rows && fixed[i]) tot = 0.0;
            y[i] = tot;
            if (DOT) dsum[0] += tot * x[i];
        }
    }
    if (DOT) spmv_dot_epilogue<ROW_THREADS>(dsum, partial, scal, finalize);
}

template <class F>
void parallel_chunks(int64_t n, F f) {
    int nt = std::min(amaru_host_threads(), 32);
    if (n < 64) nt = 1;
    if (nt == 1) {
        f(0, n);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([=] { f(n * t / nt, n * (t + 1) / nt); });
    for (auto &x : th) x.join();
}

template <int BS, bool DOT>
void launch_stream(amaru_model *m, const double *A, const double *x, double *y, int mask, int check_done, int finalize) {
    const StageLayout L = stage_layout(BS, m->tile_blks, m->tile_rows, m->tile_xcap);
    const size_t smem = m->spmv_stages * L.vbytes + (size_t)(m->spmv_stages + m->spmv_xd) * (L.rbytes + L.xbytes);
    k_spmv_stream2<BS, DOT><<<m->grid_tma, STREAM2_THREADS, smem, m->stream>>>(
        m->ntiles, reinterpret_cast<const SpmvTile *>(m->d_tiles), m->d_tmeta, A, x, y, mask, m->tile_blks, m->tile_rows,
        m->tile_xcap, m->spmv_stages, m->spmv_xd, m->spmv_sleep, m->d_partial, m->d_scal, check_done, finalize);
}

// The attribute is per function and device, and the dynamic size depends on the model's tile geometry: handles of different
// sizes live on one device (a group part next to a single-GPU handle), so the cap is the same constant for all of them.
constexpr int SPMV_SMEM_CAP = 220 * 1024;

template <int BS>
bool configure_stream(amaru_model *m) {
    const StageLayout L = stage_layout(BS, m->tile_blks, m->tile_rows, m->tile_xcap);
    const size_t smem = m->spmv_stages * L.vbytes + (size_t)(m->spmv_stages + m->spmv_xd) * (L.rbytes + L.xbytes);
    if (smem > (size_t)SPMV_SMEM_CAP) return false;
    int occ = 0;
    CUDA_CHECK(cudaFuncSetAttribute(k_spmv_stream2<BS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SPMV_SMEM_CAP));
    CUDA_CHECK(cudaFuncSetAttribute(k_spmv_stream2<BS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SPMV_SMEM_CAP));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv_stream2<BS, true>, STREAM2_THREADS, smem));
    if (occ < 1) return false;
    m->grid_tma = std::min(m->nsm * occ, m->ntiles);
    return true;
}

int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}


// ---- symmetric-storage tiles: upper pattern (col >= row in local numbering), same record format as the full tiles with
// ucol = node | fixed mask << 28 | ghost << 31 and lcol = (local column)*bs | skip-transposed << 15
template <int BS>
bool configure_sym(amaru_model *m) {
    const StageLayout L = stage_layout(BS, m->tile_blks, m->tile_rows, m->stile_xcap);
    const size_t smem = m->spmv_stages * L.vbytes + (size_t)(m->spmv_stages + m->spmv_xd) * (L.rbytes + L.xbytes) +
                        (size_t)m->sym_ystages * NCW * L.xbytes;
    if (smem > (size_t)SPMV_SMEM_CAP) return false;
    int occ = 0;
    CUDA_CHECK(cudaFuncSetAttribute(k_spmv_sym<BS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SPMV_SMEM_CAP));
    CUDA_CHECK(cudaFuncSetAttribute(k_spmv_sym<BS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SPMV_SMEM_CAP));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv_sym<BS, true>, SYM_THREADS, smem));
    if (occ < 1) return false;
    m->sgrid = std::min(m->nsm * occ, m->nstiles);
    return true;
}

void setup_sym(amaru_model *m) {
    const int bs = m->nd;
    m->use_sym = false;
    // Off by default: measured on B200 at 1 M HEX20 elements (profiles/README.md, ncu_spmv_sym_r1.txt) the kernel reads half
    // the DRAM bytes (9.24 GB vs 18.19 GB) but is bound by the shared-memory / LSU pipe (72 % of peak: the transposed products
    // re-read the values and read-modify-write the column images) and takes 3.23 ms vs 2.90 ms for the full-storage kernel.
    if (!m->use_tma || !env_int("AMARU_SPMV_SYM", 0) || m->nnodes >= (1 << 28)) return;
    m->sym_ystages = std::min(std::max(env_int("AMARU_SPMV_YSTAGES", 2), 1), MAX_YST);
    const std::vector<int32_t> &rp = m->h_rowptr, &cl = m->h_col;
    // upper pattern
    std::vector<int32_t> urp((size_t)m->nowned + 1, 0);
    for (int64_t r = 0; r < m->nowned; r++) {
        const int32_t *b = cl.data() + rp[r], *e = cl.data() + rp[r + 1];
        const int32_t *d = std::lower_bound(b, e, (int32_t)r);
        if (d == e || *d != (int32_t)r) return;   // a row without diagonal block: keep the full kernel
        urp[(size_t)r + 1] = urp[(size_t)r] + (int32_t)(e - d);
    }
    const int64_t nublk = urp[(size_t)m->nowned];
    std::vector<int32_t> ucl((size_t)nublk), usrc((size_t)nublk);
    parallel_chunks(m->nowned, [&](int64_t lo, int64_t hi) {
        for (int64_t r = lo; r < hi; r++) {
            const int32_t n = urp[(size_t)r + 1] - urp[(size_t)r], s0 = rp[r + 1] - n;
            for (int32_t k = 0; k < n; k++) {
                ucl[(size_t)urp[(size_t)r] + k] = cl[(size_t)s0 + k];
                usrc[(size_t)urp[(size_t)r] + k] = s0 + k;
            }
        }
    });
    std::vector<SpmvTile> tiles;
    for (int64_t r = 0; r < m->nowned;) {
        const int32_t b0 = urp[r];
        int64_t e = r;
        while (e < m->nowned && urp[e + 1] - b0 <= m->tile_blks && e - r < m->tile_rows) e++;
        if (e == r) return;
        SpmvTile t{};
        t.r0 = (int32_t)r; t.nrows = (int32_t)(e - r); t.b0 = b0; t.nb = urp[e] - b0;
        tiles.push_back(t);
        r = e;
    }
    const int64_t nt = (int64_t)tiles.size();
    std::vector<int32_t> nu((size_t)nt);
    parallel_chunks(nt, [&](int64_t lo, int64_t hi) {
        std::vector<int32_t> tmp;
        for (int64_t t = lo; t < hi; t++) {
            tmp.assign(ucl.begin() + tiles[t].b0, ucl.begin() + tiles[t].b0 + tiles[t].nb);
            std::sort(tmp.begin(), tmp.end());
            nu[(size_t)t] = (int32_t)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
        }
    });
    int64_t moff = 0;
    m->stile_xcap = 1;
    for (int64_t t = 0; t < nt; t++) {
        m->stile_xcap = std::max(m->stile_xcap, nu[(size_t)t]);
        tiles[t].nu = nu[(size_t)t];
        tiles[t].moff = (int32_t)moff;
        tiles[t].recints = tiles[t].nrows + 1 + tiles[t].nu + (tiles[t].nb + 1) / 2;
        moff += (tiles[t].recints + 3) / 4 * 4;
        if (moff > 2000000000LL) return;
    }
    if ((int64_t)m->stile_xcap * bs >= LCOL_SKIP) return;
    std::vector<int32_t> trec((size_t)moff + 16, 0);
    const std::vector<uint8_t> &fx = m->h_fixed;
    parallel_chunks(nt, [&](int64_t lo, int64_t hi) {
        std::vector<int32_t> tmp;
        for (int64_t t = lo; t < hi; t++) {
            const SpmvTile &T = tiles[t];
            tmp.assign(ucl.begin() + T.b0, ucl.begin() + T.b0 + T.nb);
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            int32_t *rec = trec.data() + T.moff;
            for (int32_t lr = 0; lr < T.nrows; lr++) {
                const int64_t row = T.r0 + lr;
                const int32_t dl = (int32_t)(std::lower_bound(tmp.begin(), tmp.end(), (int32_t)row) - tmp.begin());
                rec[lr] = (urp[row] - T.b0) | (dl << 19);   // masks live in the column list here
            }
            rec[T.nrows] = T.nb;
            uint32_t *uc = reinterpret_cast<uint32_t *>(rec + T.nrows + 1);
            for (size_t k = 0; k < tmp.size(); k++) {
                const int64_t node = tmp[k];
                uint32_t v = (uint32_t)node;
                for (int d = 0; d < bs; d++) v |= (uint32_t)(fx[(size_t)node * bs + d] ? 1u : 0u) << (28 + d);
                if (node >= m->nowned) v |= 1u << 31;   // ghost: its row belongs to the neighbour rank
                uc[k] = v;
            }
            uint16_t *lc = reinterpret_cast<uint16_t *>(rec + T.nrows + 1 + T.nu);
            for (int32_t lr = 0; lr < T.nrows; lr++) {
                const int64_t row = T.r0 + lr;
                for (int32_t k = urp[row] - T.b0; k < urp[row + 1] - T.b0; k++) {
                    const int32_t c = ucl[(size_t)T.b0 + k];
                    const int32_t l = (int32_t)(std::lower_bound(tmp.begin(), tmp.end(), c) - tmp.begin()) * bs;
                    const bool skip = c == (int32_t)row || c >= m->nowned;   // diagonal block / ghost column
                    lc[k] = (uint16_t)(l | (skip ? LCOL_SKIP : 0));
                }
            }
        }
    });
    m->nstiles = (int)nt;
    m->nublk = nublk;
    CUDA_CHECK(cudaMalloc(&m->d_stiles, tiles.size() * sizeof(SpmvTile)));
    CUDA_CHECK(cudaMemcpy(m->d_stiles, tiles.data(), tiles.size() * sizeof(SpmvTile), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&m->d_stmeta, trec.size() * sizeof(int32_t)));
    CUDA_CHECK(cudaMemcpy(m->d_stmeta, trec.data(), trec.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&m->d_usrc, std::max<size_t>(usrc.size(), 1) * sizeof(int32_t)));
    CUDA_CHECK(cudaMemcpy(m->d_usrc, usrc.data(), usrc.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&m->d_Asym, (size_t)nublk * bs * bs * sizeof(double) + 256));
    CUDA_CHECK(cudaMemset(m->d_Asym, 0, (size_t)nublk * bs * bs * sizeof(double) + 256));
    m->sym_meta_bytes = (int64_t)moff * 4 + (int64_t)nt * sizeof(SpmvTile);
    m->use_sym = (bs == 3) ? configure_sym<3>(m) : configure_sym<2>(m);
    m->sym_fresh = false;
}

}  // namespace

// Row tiles + tile-local column compression of the streamed SpMV (host, threaded), uploaded once per pattern.
void amaru_spmv_setup(amaru_model *m) {
    int occ = 0;
    if (m->nd == 3)
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv<3, true>, ROW_THREADS, 0));
    else
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv<2, true>, ROW_THREADS, 0));
    m->grid_rows = m->nsm * std::max(occ, 1);
    m->use_tma = false;
    m->grid_tma = 0;
    if (env_int("AMARU_SPMV_SIMPLE", 0) || m->nowned == 0) return;
    const int bs = m->nd;
    m->spmv_warps = NCW;
    // defaults from the sweep on B200 at 1 M HEX20 elements (profiles/spmv_sweep_r1.txt): the consumers are issue-bound,
    // so small stages (-> 3 CTAs / 15 warps per SM) beat deeper pipelines
    m->spmv_ver = 2;
    m->spmv_xd = std::min(std::max(env_int("AMARU_SPMV_XD", m->spmv_ver == 2 ? 2 : 1), 1), XD_MAX);
    m->spmv_stages = std::min(std::max(env_int("AMARU_SPMV_STAGES", 2), 2), MAX_STAGES);
    m->tile_blks = std::min(env_int("AMARU_SPMV_TILE", bs == 3 ? 232 : 522), 8191);   // 13-bit local diagonal index
    if ((int64_t)m->tile_blks * bs > 65535) m->tile_blks = 65535 / bs;   // lcol holds (local column)*bs in 16 bits
    m->tile_rows = env_int("AMARU_SPMV_TILE_ROWS", 32);
    m->spmv_sleep = env_int("AMARU_SPMV_SLEEP", 100);   // ns between probes of the producer's empty-barrier wait

    const std::vector<int32_t> &rp = m->h_rowptr, &cl = m->h_col;
    std::vector<SpmvTile> tiles;
    for (int64_t r = 0; r < m->nowned;) {
        const int32_t b0 = rp[r];
        int64_t e = r;
        while (e < m->nowned && rp[e + 1] - b0 <= m->tile_blks && e - r < m->tile_rows) e++;
        if (e == r) return;   // one row is longer than a tile: keep the per-lane kernel
        SpmvTile t{};
        t.r0 = (int32_t)r; t.nrows = (int32_t)(e - r); t.b0 = b0; t.nb = rp[e] - b0;
        tiles.push_back(t);
        r = e;
    }
    const int64_t nt = (int64_t)tiles.size();
    std::vector<int32_t> nu((size_t)nt);
    parallel_chunks(nt, [&](int64_t lo, int64_t hi) {
        std::vector<int32_t> tmp;
        for (int64_t t = lo; t < hi; t++) {
            tmp.assign(cl.begin() + tiles[t].b0, cl.begin() + tiles[t].b0 + tiles[t].nb);
            std::sort(tmp.begin(), tmp.end());
            nu[(size_t)t] = (int32_t)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
        }
    });
    int64_t moff = 0;
    m->tile_xcap = 1;
    for (int64_t t = 0; t < nt; t++) {
        m->tile_xcap = std::max(m->tile_xcap, nu[(size_t)t]);
        tiles[t].nu = nu[(size_t)t];
        tiles[t].moff = (int32_t)moff;
        tiles[t].recints = tiles[t].nrows + 1 + tiles[t].nu + (tiles[t].nb + 1) / 2;
        moff += (tiles[t].recints + 3) / 4 * 4;   // records start 16-byte aligned
        if (moff > 2000000000LL) return;
    }
    std::vector<int32_t> trec((size_t)moff + 16, 0);
    const std::vector<uint8_t> &fx = m->h_fixed;
    parallel_chunks(nt, [&](int64_t lo, int64_t hi) {
        std::vector<int32_t> tmp;
        for (int64_t t = lo; t < hi; t++) {
            const SpmvTile &T = tiles[t];
            tmp.assign(cl.begin() + T.b0, cl.begin() + T.b0 + T.nb);
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            int32_t *rec = trec.data() + T.moff;
            for (int32_t lr = 0; lr < T.nrows; lr++) {
                const int64_t row = T.r0 + lr;
                int32_t mask = 0;
                for (int d = 0; d < bs; d++) mask |= (fx[(size_t)row * bs + d] ? 1 : 0) << d;
                auto it = std::lower_bound(tmp.begin(), tmp.end(), (int32_t)row);
                const int32_t dl = (it != tmp.end() && *it == (int32_t)row) ? (int32_t)(it - tmp.begin()) : 0;
                rec[lr] = (rp[row] - T.b0) | (mask << 16) | (dl << 19);
            }
            rec[T.nrows] = T.nb;
            std::memcpy(rec + T.nrows + 1, tmp.data(), tmp.size() * sizeof(int32_t));
            uint16_t *lc = reinterpret_cast<uint16_t *>(rec + T.nrows + 1 + T.nu);
            for (int32_t k = 0; k < T.nb; k++)
                lc[k] = (uint16_t)((std::lower_bound(tmp.begin(), tmp.end(), cl[(size_t)T.b0 + k]) - tmp.begin()) * bs);   // pre-scaled by bs
        }
    });
    m->ntiles = (int)nt;
    CUDA_CHECK(cudaMalloc(&m->d_tiles, tiles.size() * sizeof(SpmvTile)));
    CUDA_CHECK(cudaMemcpy(m->d_tiles, tiles.data(), tiles.size() * sizeof(SpmvTile), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&m->d_tmeta, trec.size() * sizeof(int32_t)));
    CUDA_CHECK(cudaMemcpy(m->d_tmeta, trec.data(), trec.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    m->spmv_meta_bytes = (int64_t)moff * 4 + (int64_t)nt * sizeof(SpmvTile);
    m->use_tma = (bs == 3) ? configure_stream<3>(m) : configure_stream<2>(m);
    setup_sym(m);
}

// y = A x on the owned rows (+ p·Ap partial dot and CG scalar finalisation when dot != 0)
void amaru_spmv_launch(amaru_model *m, const double *A, const double *x, double *y, int mask, int dot, int check_done,
                       int finalize) {
    if (m->use_tma) {
        if (m->nd == 3) {
            if (dot) launch_stream<3, true>(m, A, x, y, mask, check_done, finalize);
            else launch_stream<3, false>(m, A, x, y, mask, check_done, finalize);
        } else {
            if (dot) launch_stream<2, true>(m, A, x, y, mask, check_done, finalize);
            else launch_stream<2, false>(m, A, x, y, mask, check_done, finalize);
        }
    } else {
        const int64_t need = (m->nowned + (ROW_THREADS / 32) - 1) / (ROW_THREADS / 32);
        const int g = (int)std::max<int64_t>(1, std::min<int64_t>(need, m->grid_rows));
        if (m->nd == 3) {
            if (dot) k_spmv<3, true><<<g, ROW_THREADS, 0, m->stream>>>(m->nowned, m->d_rowptr, m->d_col, A, x, y, m->d_fixed, mask, m->d_partial, m->d_scal, check_done, finalize);
            else k_spmv<3, false><<<g, ROW_THREADS, 0, m->stream>>>(m->nowned, m->d_rowptr, m->d_col, A, x, y, m->d_fixed, mask, m->d_partial, m->d_scal, check_done, finalize);
        } else {
            if (dot) k_spmv<2, true><<<g, ROW_THREADS, 0, m->stream>>>(m->nowned, m->d_rowptr, m->d_col, A, x, y, m->d_fixed, mask, m->d_partial, m->d_scal, check_done, finalize);
            else k_spmv<2, false><<<g, ROW_THREADS, 0, m->stream>>>(m->nowned, m->d_rowptr, m->d_col, A, x, y, m->d_fixed, mask, m->d_partial, m->d_scal, check_done, finalize);
        }
    }
    m->launches++;
    CUDA_CHECK(cudaGetLastError());
}

// d_Asym <- the upper blocks of the current system matrix (called whenever d_A changed, before the next CG loop)
void amaru_spmv_sym_refresh(amaru_model *m) {
    if (!m->use_sym || m->sym_fresh) return;
    const int b2 = m->nd * m->nd;
    const int64_t n = m->nublk * b2;
    if (n > 0) {
        const int g = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)m->nsm * 16));
        k_extract_upper<<<g, 256, 0, m->stream>>>(m->nublk, b2, m->d_usrc, m->d_A, m->d_Asym);
        m->launches++;
        CUDA_CHECK(cudaGetLastError());
    }
    m->sym_fresh = true;
}

// y = A x on the owned rows from the symmetric storage (+ p.Ap); y is zeroed here, the kernel accumulates into it
void amaru_spmv_sym_launch(amaru_model *m, const double *x, double *y, int mask, int dot, int check_done, int finalize) {
    CUDA_CHECK(cudaMemsetAsync(y, 0, (size_t)m->nowned * m->nd * sizeof(double), m->stream));
    const int bs = m->nd;
    const StageLayout L = stage_layout(bs, m->tile_blks, m->tile_rows, m->stile_xcap);
    const size_t smem = m->spmv_stages * L.vbytes + (size_t)(m->spmv_stages + m->spmv_xd) * (L.rbytes + L.xbytes) +
                        (size_t)m->sym_ystages * NCW * L.xbytes;
#define SYMLAUNCH(BS, DOT)                                                                                                   \
    k_spmv_sym<BS, DOT><<<m->sgrid, SYM_THREADS, smem, m->stream>>>(                                                           \
        m->nstiles, reinterpret_cast<const SpmvTile *>(m->d_stiles), m->d_stmeta, m->d_Asym, x, y, mask, m->tile_blks, m->tile_rows, \
        m->stile_xcap, m->spmv_stages, m->spmv_xd, m->sym_ystages, m->spmv_sleep, m->d_partial, m->d_scal, check_done, finalize)
    if (bs == 3) {
        if (dot) SYMLAUNCH(3, true);
        else SYMLAUNCH(3, false);
    } else {
        if (dot) SYMLAUNCH(2, true);
        else SYMLAUNCH(2, false);
    }
#undef SYMLAUNCH
    m->launches++;
    CUDA_CHECK(cudaGetLastError());
}
