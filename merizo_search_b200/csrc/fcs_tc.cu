// fcs_tc.cu -- K3/K4: the large-batch path.  bf16 tcgen05 GEMM with a fused threshold filter, then
// an exact fp32 rescore of the surviving candidates.
//
// Replaces knn_exact_faiss (reference dbsearch.py:213-248: IndexFlat inner product + ResultHeap) for
// large query batches.  A genuine dense contraction S = Q[nq,128] x DB[N,128]^T, so it runs on the
// 5th-generation tensor cores:
//
//   K3  tc_gemm_filter_kernel -- persistent, warp-specialised, one CTA per SM (21 warps):
//        * 512 queries (4 tiles of M=128) stay resident in shared memory as bf16 UMMA operand images;
//          DB rows stream through a 3-stage ring of 128-row bf16 operand images (32 KB each).  Both the
//          query images and the DB images are stored in HBM already in the canonical K-major
//          no-swizzle core-matrix layout, so every operand load is ONE contiguous 1-D bulk copy
//          (cp.async.bulk / UBLKCP, the TMA engine) -- no tensor maps.
//        * warp 0: TMA producer.  warps 1-4: one thread each issues tcgen05.mma.kind::f16, M128 x N128 x
//          K16, 8 per (query tile, DB tile), one query tile per issuer (a single thread can only issue one
//          MMA per ~75 cycles; two or more issuers reach the 64-cycle math rate, scripts/micro/
//          umma_two_issuers.cu).  Accumulators: fp32 in TMEM, one 128-column buffer per query tile = all
//          512 columns; the epilogue of tile t drains its buffer while the MMAs of the other three run.
//        * warps 5..20: epilogue, one THREAD PER QUERY (tcgen05.ld 32x32b: lane = accumulator row).
//          Each thread reduces its scores 32 at a time with 3-input max, compares the group maximum
//          with the query's threshold and only on a hit re-reads the 8-column sub-group from TMEM and appends
//          (score,row) keys to the query's candidate buffer in HBM (slots reserved with one atomic per 16).
//
//   The threshold is a per-query constant during a launch.  It comes from a SAMPLE of the database, not from a
//   running top-k': the rows above the m-th best score of a sample of S rows number about m*N/S in the whole
//   shard, so a small m and a large N/S give a good threshold after very little work.  A search is therefore
//     round 0   8-32 tiles spread evenly over the shard (1024-4096 rows), every score recorded (slot = sample row);
//     round i   a nested strided sample 8x the size of what has been seen, threshold = the rank-m_i score of that
//               (m_i ~ 18: about 128 rows of the round qualify -- hits are what the epilogue pays for);
//     sweep     every tile that was not sampled (~94 % of the rows), threshold = rank-24 score of the last sample:
//               about max(2.4 k', 384) rows qualify, a hit density the epilogue absorbs;
//   each followed by tc_select_kernel (one block per query, radix select on the order-preserving score word), and each
//   GEMM round after the first launched programmatically behind its selection kernel (setup and operand prefetch overlap
//   the selection's tail).  Every tile is multiplied exactly once.  cfg3 (10 M rows x 4096 queries) runs 5 GEMM launches:
//   16 / 92 / 617 / 4158 tiles of samples (0.8 ms, 6 % of the rows) and a 73 k-tile sweep (6.8 ms) -- where round 1's
//   geometric rounds (x3 rows per round at 2k' hits each) needed 9 launches and 2.8 ms for the first 2.2 M rows.
//
//   K4  tc_rescore_kernel -- one block (4 warps) per query.  Phase A: the k' best approximate candidates are gathered
//        (fp32 rows), their inner products recomputed exactly, the k best kept; certificate
//        s_k(exact) > t' + eps with t' = the k'-th approximate score (every non-candidate has an approximate
//        score <= t').  Phase B, only when A fails: ALL rows above the sweep threshold (they are still in the
//        buffer, behind the k' best) are rescored; certificate with t' = the sweep threshold, which lies far
//        below the k-th score.  eps = |q|*max|r - bf16(r)| + |q - bf16(q)|*max|bf16(r)| + fp32 accumulation
//        slack: a rigorous Cauchy-Schwarz bound from the measured rounding errors of both operands.
//        Queries that fail both (or lost candidates to an overflowing buffer) are queued ON THE DEVICE for
//        the exact fp32 scan (K2, launched with a device-side count): no host round trip anywhere in a search.
#include <cuda_bf16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "fcs_common.cuh"
#include "fcs_internal.h"
#include "fcs_tc.h"


#ifndef FCS_TC_BPOLICY
#define FCS_TC_BPOLICY 1  // L2 policy of the DB tile copies: 1 evict_normal (default: the CTAs that sweep the same tiles for other
                         // query groups find them in L2; cfg4b 8.97 vs 9.22 ms), 0 evict_first
#endif

namespace fcs {

namespace {

constexpr int TC_M = 128;                        // queries per UMMA tile
constexpr int TC_QT = 4;                         // query tiles resident per CTA
constexpr int TC_QGROUP = TC_M * TC_QT;          // 512
constexpr int TC_N = 128;                        // DB rows per B tile = N of one MMA
constexpr int A_TILE_BYTES = TC_M * DIM * 2;     // 32 KB
constexpr int B_TILE_BYTES = TC_N * DIM * 2;     // 32 KB
constexpr int TC_STAGES = 3;
constexpr int TC_PREFETCH = 6;                     // DB tiles prefetched into L2 ahead of the ring
constexpr int TC_RES_MAX = 16;                    // candidate slots reserved per atomic (upper bound, see res_block)
constexpr int TC_MMA_WARPS = 4;                   // MMA-issuing warps (one thread each), one query tile apiece
constexpr int TC_THREADS = (1 + TC_MMA_WARPS + 16) * 32;  // producer + MMA issuers + 16 epilogue warps
constexpr int TC_CAP = 4096;                     // candidate slots per query (k' up to ~850)
constexpr int TC_CAP_BIG = 16384;                // ... for larger k (up to FCS_MAX_K = 2048)
constexpr int TC_R0_TILES = TC_CAP / TC_N;       // round 0 records every score of 32 tiles
constexpr int TC_SMEM = TC_QT * A_TILE_BYTES + TC_STAGES * B_TILE_BYTES + 32 * 8 + 16;
constexpr int TC_MAX_KPRIME = 3328;               // k' for k = FCS_MAX_K
constexpr int TC_WARP_K = 128;                    // k up to which the rescore keeps its top-k list in registers
constexpr int TC_MAX_RANK = 1024;                // largest selection rank used for a threshold

// Operand image layout (K-major, SWIZZLE_NONE "interleave" canonical layout, cute mma_sm100_desc):
// 8x8 core matrices of 128 contiguous bytes (8 rows x 16 B); the 16 core matrices along K of one
// 8-row group are contiguous (LBO = 128 B), 8-row groups are SBO = 2048 B apart.
__host__ __device__ __forceinline__ uint32_t image_offset(int row_in_tile, int k8) {
    return uint32_t((row_in_tile >> 3) * 2048 + k8 * 128 + (row_in_tile & 7) * 16);
}
constexpr uint64_t DESC_BASE = (uint64_t(128 >> 4) << 16) | (uint64_t(2048 >> 4) << 32) | (uint64_t(1) << 46);
// instruction descriptor, kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major both, N>>3 at 17, M>>4 at 24
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(TC_N >> 3) << 17) | (uint32_t(TC_M >> 4) << 24);

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) { return DESC_BASE | uint64_t((smem_addr >> 4) & 0x3FFF); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ------------------------------------------------------------------------------------------------
// one-time: fp32 (row-swizzled) rows -> bf16 operand images.  Also, for the exactness certificate, the largest
// norm of a ROUNDED row and the largest norm of a row's rounding error over the shard (stats[0], stats[1]:
// squared, as float bits -- non-negative floats order like unsigned integers).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tc_build_bimg_kernel(const float* __restrict__ rows, int64_t n_rows, int64_t n_tiles,
                                                            uint8_t* __restrict__ b_img, unsigned* __restrict__ stats) {
    const int64_t total = n_tiles * TC_N * 16;  // one thread per (row, 8-element k-chunk)
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int k8 = int(t & 15);
        const int64_t row = t >> 4;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (row < n_rows) {
            const float4* src = reinterpret_cast<const float4*>(rows + row * DIM);
            const float4 a = src[swz_chunk(2 * k8, row)];
            const float4 b = src[swz_chunk(2 * k8 + 1, row)];
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
        float ss = 0.f, dd = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float r = bf16_round(v[i]);
            ss = fmaf(r, r, ss);
            dd = fmaf(v[i] - r, v[i] - r, dd);
        }
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {  // 16 threads = one row
            ss += __shfl_xor_sync(FULL, ss, o);
            dd += __shfl_xor_sync(FULL, dd, o);
        }
        if (k8 == 0) {
            atomicMax(stats + 0, __float_as_uint(ss));
            atomicMax(stats + 1, __float_as_uint(dd));
        }
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
        o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
        const int64_t tile = row / TC_N;
        const int rit = int(row - tile * TC_N);
        *reinterpret_cast<uint4*>(b_img + tile * B_TILE_BYTES + image_offset(rit, k8)) = o;
    }
}

// ------------------------------------------------------------------------------------------------
// per search: normalise queries, write fp32 copy + bf16 operand images, per-query certificate slack,
// reset per-query state
// ------------------------------------------------------------------------------------------------
struct TcPrepParams {
    const float* q_raw;
    int nq, nq_pad, qnorm;
    float* qn;
    uint8_t* a_img;
    float* thr;
    float* thr_sel;
    unsigned* cnt;
    unsigned* sel_cnt;
    unsigned* flags;
    float* eps;
    unsigned* n_flagged;
    unsigned first_round_slots;
    float max_rhat;   // max |bf16(row)|
    float max_dr;     // max |row - bf16(row)|
};

__global__ void __launch_bounds__(128) tc_prep_kernel(const TcPrepParams p) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (q == 0 && lane == 0) *p.n_flagged = 0u;
    if (q >= p.nq_pad) return;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < p.nq) {
        v = reinterpret_cast<const float4*>(p.q_raw + size_t(q) * DIM)[lane];
        float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(FULL, ss, o);
        if (p.qnorm != FCS_QNORM_NONE) {
            const float d = fmaxf(sqrtf(ss), (p.qnorm == FCS_QNORM_COSINE) ? 1e-8f : 1e-12f);
            v.x = v.x / d; v.y = v.y / d; v.z = v.z / d; v.w = v.w / d;
        }
        reinterpret_cast<float4*>(p.qn + size_t(q) * DIM)[lane] = v;
        // |q| and |q - bf16(q)| of the query as it is multiplied
        const float ex = v.x - bf16_round(v.x), ey = v.y - bf16_round(v.y), ez = v.z - bf16_round(v.z), ew = v.w - bf16_round(v.w);
        float qq = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        float dq = ex * ex + ey * ey + ez * ez + ew * ew;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            qq += __shfl_xor_sync(FULL, qq, o);
            dq += __shfl_xor_sync(FULL, dq, o);
        }
        if (lane == 0) {
            const float nq_ = sqrtf(qq), ndq = sqrtf(dq);
            // q.r - q^.r^ = q.(r - r^) + (q - q^).r^  =>  |error| <= |q| max|r - r^| + |q - q^| max|r^|;
            // + fp32 accumulation of 128 exact bf16 products in the tensor core (3e-5 |q||r| is generous)
            // + the rounding of the fp32 rescore itself
            p.eps[q] = 1.002f * (nq_ * p.max_dr + ndq * p.max_rhat) + 3.2e-5f * nq_ * p.max_rhat;
            p.thr[q] = -INFINITY;
            p.thr_sel[q] = -INFINITY;
            p.cnt[q] = p.first_round_slots;  // round 0 indexes its candidates by sample row
            p.sel_cnt[q] = 0u;
            p.flags[q] = 0u;
        }
    }
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&p0);
    o.y = *reinterpret_cast<uint32_t*>(&p1);
    const int tile = q / TC_M, m = q % TC_M;
    *reinterpret_cast<uint2*>(p.a_img + size_t(tile) * A_TILE_BYTES + image_offset(m, lane >> 1) + (lane & 1) * 8) = o;
}

// ------------------------------------------------------------------------------------------------
// K3
// ------------------------------------------------------------------------------------------------
struct TcGemmParams {
    const uint8_t* a_img;  // [n_qgroups*4][32 KB]
    const uint8_t* b_img;  // [n_tiles][32 KB]
    int64_t n_rows;
    int nq;
    int n_qgroups;
    // the DB tiles of this round, enumerated by idx in [0, n_idx):
    //   sample round      (comp_T == 0): tile = (j0 + idx) * stride
    //   complement sweep  (comp_T  > 0): every tile except the sampled ones {j * stride : j < comp_T}
    int64_t n_idx, j0, stride, comp_T;
    const float* thr;      // [nq] approximate-score threshold (strict >)
    unsigned* cnt;         // [nq] append counters
    uint64_t* cand;        // [nq][cap] approximate keys (score, LOCAL row); 0 = unused slot
    int cap;               // candidate slots per query (TC_CAP or TC_CAP_BIG)
    int first_round;       // round 0: thresholds are -inf, slot = idx * 128 + column, no atomics
    int trace_on;          // debug builds (FCS_TC_TRACE): record cycle stamps in this launch
    int res_block;         // slots reserved per atomic: small when many CTAs share a query group (unused slots are waste)
    int pdl;               // launched with programmatic stream serialization behind a selection kernel (rounds >= 1): only
                           // the epilogue depends on it (thresholds, candidate buffers); operands are immutable by then
};

__device__ __forceinline__ int64_t tile_of(const TcGemmParams& p, int64_t idx) {
    if (p.comp_T == 0) return (p.j0 + idx) * p.stride;
    const int64_t gap = p.stride - 1;            // unsampled tiles per stride block
    const int64_t body = p.comp_T * gap;
    if (idx < body) {
        const int64_t b = idx / gap;
        return b * p.stride + 1 + (idx - b * gap);
    }
    return idx + p.comp_T;
}

#ifdef FCS_TC_TRACE
// debug build only: cycle stamps of CTA 0 (MMA thread: row 0/1; epilogue warp of tile 0, quadrant 2: rows 2..4)
__device__ long long g_trace[8][256];
#define TRACE(row, idx, cond) do { if (p.trace_on && (cond) && (idx) < 256) g_trace[row][idx] = clock64(); } while (0)
#else
#define TRACE(row, idx, cond) do { } while (0)
#endif

// Per-thread slot reservation: one atomic buys res_block slots of the query's buffer (x = first slot, y = slots used;
// y == ~0u: nothing reserved yet).  Deliberately NOT inlined: there are 32 call sites per 32-column part, and the
// epilogue loop has to stay small enough for the instruction cache (an inlined append per column cost 3-5 k cycles
// per hit in instruction-cache misses).
__device__ __noinline__ uint2 tc_append(unsigned* cnt_q, uint64_t* cand_q, uint2 res, int res_block, uint32_t score_bits, uint32_t row, int cap) {
    if (res.y >= unsigned(res_block)) {
        res.x = atomicAdd(cnt_q, unsigned(res_block));
        res.y = 0;
    }
    const unsigned slot = res.x + res.y++;
    if (slot < unsigned(cap)) cand_q[slot] = make_key(__uint_as_float(score_bits), row);
    return res;
}
// unused slots of the last reservation must read as empty
__device__ __forceinline__ void tc_close_reservation(uint64_t* cand_q, uint2 res, int res_block, int cap) {
    if (res.y == ~0u) return;  // this thread never appended
    for (; res.y < unsigned(res_block); ++res.y) {
        const unsigned slot = res.x + res.y;
        if (slot < unsigned(cap)) cand_q[slot] = 0ull;
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_filter_kernel(const TcGemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + TC_QT * A_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + TC_STAGES * B_TILE_BYTES);
    uint64_t* full_b = bars;             // [TC_STAGES]
    uint64_t* empty_b = bars + 4;        // [TC_STAGES]
    uint64_t* tmem_full = bars + 8;      // [tile]
    uint64_t* tmem_empty = bars + 12;    // [tile]
    uint64_t* a_full = bars + 16;
    uint64_t* a_empty = bars + 17;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < TC_STAGES; ++i) {
            mbar_init(&full_b[i], 1);
            mbar_init(&empty_b[i], TC_MMA_WARPS);
        }
        for (int i = 0; i < TC_QT; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);  // the 4 epilogue warps of a query tile
        }
        mbar_init(a_full, 1);
        mbar_init(a_empty, TC_MMA_WARPS);
        mbar_fence_init();
    }
    if (warp == 1) {  // whole warp: TMEM allocation (all 512 columns; one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // contiguous share of the linearised (query group, tile index) steps of this round
    const int64_t n_tiles = p.n_idx;
    const int64_t steps = int64_t(p.n_qgroups) * n_tiles;
    const int64_t s_begin = steps * blockIdx.x / gridDim.x;
    const int64_t s_end = steps * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (lane == 0) {
#if FCS_TC_BPOLICY == 0
            const uint64_t pol_stream = policy_evict_first();
#else
            const uint64_t pol_stream = policy_evict_normal();
#endif
            const uint64_t pol_keep = policy_evict_normal();
            uint32_t it = 0, seg = 0;
            for (int64_t s = s_begin; s < s_end; ++seg) {
                const int64_t qg = s / n_tiles, ti = s - qg * n_tiles;
                const int64_t seg_len = (s_end - s < n_tiles - ti) ? (s_end - s) : (n_tiles - ti);
                mbar_wait(a_empty, (seg & 1u) ^ 1u);  // previous segment's MMAs are done with the query images
                mbar_arrive_expect_tx(a_full, TC_QT * A_TILE_BYTES);
                for (int t = 0; t < TC_QT; ++t)
                    bulk_g2s(sA + t * A_TILE_BYTES, p.a_img + (size_t(qg) * TC_QT + t) * A_TILE_BYTES, A_TILE_BYTES, a_full, pol_keep);
                for (int64_t j = 0; j < seg_len; ++j, ++it) {
                    const uint32_t st = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
                    // only 2 stages (64 KB) can be in flight behind the one being consumed: pull tiles further
                    // ahead into L2 so the bulk copies below see L2 latency, not DRAM latency
                    if (j + TC_PREFETCH < seg_len)
                        bulk_prefetch_l2(p.b_img + size_t(tile_of(p, ti + j + TC_PREFETCH)) * B_TILE_BYTES, B_TILE_BYTES);
                    mbar_wait(&empty_b[st], ph ^ 1u);
                    TRACE(5, it, blockIdx.x == 0);
                    mbar_arrive_expect_tx(&full_b[st], B_TILE_BYTES);
                    bulk_g2s(sB + st * B_TILE_BYTES, p.b_img + size_t(tile_of(p, ti + j)) * B_TILE_BYTES, B_TILE_BYTES, &full_b[st], pol_stream);
                }
                s += seg_len;
            }
        }
        __syncwarp();
    } else if (warp <= TC_MMA_WARPS) {
        // ---------------------------------------------------------------- MMA issuers (one thread per warp).
        // One thread can issue a tcgen05.mma only every ~75 cycles (scripts/micro/umma_two_issuers.cu); an
        // M128 x N128 x K16 MMA is 64 cycles of tensor-pipe work, so two issuers keep the pipe full.
        const int t_first = (warp - 1) * (TC_QT / TC_MMA_WARPS);
        if (lane == 0) {
            const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
            uint32_t it = 0, seg = 0;
            for (int64_t s = s_begin; s < s_end; ++seg) {
                const int64_t qg = s / n_tiles, ti = s - qg * n_tiles;
                const int64_t seg_len = (s_end - s < n_tiles - ti) ? (s_end - s) : (n_tiles - ti);
                mbar_wait(a_full, seg & 1u);
                tc_fence_after();
                for (int64_t j = 0; j < seg_len; ++j, ++it) {
                    const uint32_t st = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
                    const uint32_t tph = it & 1u;
                    mbar_wait(&full_b[st], ph);
                    tc_fence_after();
                    TRACE(6, it, blockIdx.x == 0 && t_first == 0);
#pragma unroll
                    for (int tt = 0; tt < TC_QT / TC_MMA_WARPS; ++tt) {
                        const int t = t_first + tt;
                        mbar_wait(&tmem_empty[t], tph ^ 1u);  // epilogue has drained this tile's accumulator
                        tc_fence_after();
                        TRACE(0, it, blockIdx.x == 0 && t == 0);
                        const uint32_t d_tmem = tmem_base + uint32_t(t * TC_N);
#pragma unroll
                        for (int k = 0; k < DIM / 16; ++k) {
                            const uint64_t ad = make_desc(a_addr + t * A_TILE_BYTES + k * 256);
                            const uint64_t bd = make_desc(b_addr + st * B_TILE_BYTES + k * 256);
                            tc_mma(d_tmem, ad, bd, k > 0 ? 1u : 0u);
                        }
                        tc_commit(&tmem_full[t]);
                        TRACE(1, it, blockIdx.x == 0 && t == 0);
                    }
                    tc_commit(&empty_b[st]);  // B stage free once every issuer's MMAs have read it
                }
                tc_commit(a_empty);
                s += seg_len;
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- epilogue: thread = query
        const int t = (warp - 1 - TC_MMA_WARPS) >> 2;  // query tile (4 consecutive warps cover the 4 lane quadrants)
        const int quad = warp & 3;      // TMEM lane quadrant this warp may access
        const uint32_t lane_base = uint32_t(quad * 32) << 16;
        uint32_t it = 0;
        // Everything the preceding selection kernel writes (thresholds, counters, compacted candidates) is touched by the
        // epilogue only: with a programmatic launch the producer and the MMA issuers have been filling the pipeline while
        // that kernel was finishing.
        if (p.pdl) pdl_wait();
        for (int64_t s = s_begin; s < s_end;) {
            const int64_t qg = s / n_tiles, ti = s - qg * n_tiles;
            const int64_t seg_len = (s_end - s < n_tiles - ti) ? (s_end - s) : (n_tiles - ti);
            const int q = int(qg) * TC_QGROUP + t * TC_M + quad * 32 + lane;
            const float thr = (q < p.nq) ? p.thr[q] : INFINITY;
            const int qc = q < p.nq ? q : 0;
            unsigned* cnt_q = p.cnt + qc;
            uint64_t* cand_q = p.cand + size_t(qc) * p.cap;
            uint2 res = make_uint2(0u, ~0u);
            for (int64_t j = 0; j < seg_len; ++j, ++it) {
                const uint32_t tph = it & 1u;
                // the tile's first row does not depend on the accumulator: computed BEFORE the wait (the complement
                // mapping of the sweep has a 64-bit division; after the wait it sat on the tile's critical chain)
                int64_t row_base = tile_of(p, ti + j) * TC_N;
                asm volatile("" : "+l"(row_base));
                const int64_t left = p.n_rows - row_base;
                int rows_here = left < TC_N ? int(left) : TC_N;  // the last tile of the shard is zero-padded
                asm volatile("" : "+r"(rows_here));  // materialise it here: volatile statements keep their order
                mbar_wait(&tmem_full[t], tph);
                tc_fence_after();
                TRACE(2, it, blockIdx.x == 0 && warp == 5 && lane == 0);
                if (p.first_round) {
                    // round 0: every score is recorded, slot = row of the sample (no threshold, no atomics)
                    uint64_t* dst = cand_q + (ti + j) * TC_N;
#pragma unroll 1
                    for (int part = 0; part < TC_N / 32; ++part) {
                        uint32_t r[32];
                        tc_ld32(tmem_base + lane_base + uint32_t(t * TC_N + part * 32), r);
                        tc_wait_ld();
                        if (q < p.nq) {
#pragma unroll
                            for (int c = 0; c < 32; c += 2) {
                                const int64_t row = row_base + part * 32 + c;
                                ulonglong2 kk;
                                kk.x = row < p.n_rows ? make_key(__uint_as_float(r[c]), uint32_t(row)) : 0ull;
                                kk.y = row + 1 < p.n_rows ? make_key(__uint_as_float(r[c + 1]), uint32_t(row + 1)) : 0ull;
                                *reinterpret_cast<ulonglong2*>(dst + part * 32 + c) = kk;
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[t]);
                    continue;
                }
                // One accumulator buffer per tile: the tile's chain MMA -> epilogue -> next MMA sets the step period (pipeline
                // trace in profiles/r02_experiments.md), so everything between "accumulator full" and "accumulator handed
                // back" is on the critical path of the whole CTA.
                //  * Fast pass: four 32-column tcgen05.ld, each reduced with 3-input max to one maximum per 8 columns; the
                //    comparison results go into a 16-bit mask (bit 4*part + group) -- no warp vote inside the loop.  ONE warp
                //    OR-reduction per tile then says whether any lane has a hit anywhere (rare once the threshold is warm).
                //  * Slow path: the hit 8-column groups are re-read from TMEM (taking the values from the registers of the
                //    32-column load instead was 7-10 % slower on every workload: the live registers delay the next load).
                //  * Up to two hits per tile are parked in registers and appended AFTER the accumulator has been handed back:
                //    the append's call, store and (every 16th time) L2 atomic are off the chain.  A third hit in one tile is
                //    appended at once.  The append is not inlined: the loop has to stay instruction-cache resident.
                int pend_n = 0;
                uint32_t pend_v0 = 0, pend_v1 = 0;
                int pend_c0 = 0, pend_c1 = 0;
                unsigned gm = 0;
#pragma unroll
                for (int part = 0; part < TC_N / 32; ++part) {
                    uint32_t r[32];
                    tc_ld32(tmem_base + lane_base + uint32_t(t * TC_N + part * 32), r);
                    tc_wait_ld();
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float* f = reinterpret_cast<const float*>(&r[g * 8]);
                        const float mg = fmaxf(max3(f[0], f[1], f[2]), max3(max3(f[3], f[4], f[5]), f[6], f[7]));
                        gm |= (mg > thr) ? (1u << (part * 4 + g)) : 0u;
                    }
                }
                unsigned wgm = __reduce_or_sync(FULL, gm);
                while (wgm) {  // slow path: re-read the hit 8-column groups from TMEM (warp-uniform loop)
                    const int pg = __ffs(wgm) - 1;
                    wgm &= wgm - 1;
                    uint32_t v8[8];
                    __syncwarp();  // lanes left the divergent loop below at different times
                    tc_ld8(tmem_base + lane_base + uint32_t(t * TC_N + pg * 8), v8);
                    tc_wait_ld();
                    unsigned hits = 0;
#pragma unroll
                    for (int c = 0; c < 8; ++c) hits |= (__uint_as_float(v8[c]) > thr) ? (1u << c) : 0u;
                    while (hits) {
                        const int c = __ffs(hits) - 1;
                        hits &= hits - 1;
                        uint32_t bits = v8[0];
#pragma unroll
                        for (int cc = 1; cc < 8; ++cc) bits = (c == cc) ? v8[cc] : bits;
                        const int col = pg * 8 + c;
                        if (col >= rows_here) continue;  // zero padding rows of the shard's last tile
                        if (pend_n == 0) {
                            pend_v0 = bits;
                            pend_c0 = col;
                            pend_n = 1;
                        } else if (pend_n == 1) {
                            pend_v1 = bits;
                            pend_c1 = col;
                            pend_n = 2;
                        } else {
                            res = tc_append(cnt_q, cand_q, res, p.res_block, bits, uint32_t(row_base) + uint32_t(col), p.cap);
                        }
                    }
                }
                TRACE(3, it, blockIdx.x == 0 && warp == 5 && lane == 0);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[t]);
                if (pend_n > 0) res = tc_append(cnt_q, cand_q, res, p.res_block, pend_v0, uint32_t(row_base) + uint32_t(pend_c0), p.cap);
                if (pend_n > 1) res = tc_append(cnt_q, cand_q, res, p.res_block, pend_v1, uint32_t(row_base) + uint32_t(pend_c1), p.cap);
                TRACE(4, it, blockIdx.x == 0 && warp == 5 && lane == 0);
            }
            if (!p.first_round) tc_close_reservation(cand_q, res, p.res_block, p.cap);
            s += seg_len;
        }
    }


    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// after every round: one block per query finds the rank-th largest approximate score of the query's candidates
// (radix select on the 32-bit order-preserving score word), publishes it as the threshold and moves the candidates
// at or above it to the front of the buffer.
//   compact   (sample rounds): only the survivors are kept (cnt = survivors); thr[q] = the rank-th score.
//   partition (last round):    survivors in front, every other real candidate behind them (cnt = all real
//                              candidates, sel_cnt = survivors); thr_sel[q] = the rank-th score, thr[q] untouched.
// Fewer than `rank` candidates: everything is kept and the threshold stays what it was.
// ------------------------------------------------------------------------------------------------
constexpr int SEL_NT = 128;
struct TcSelectParams {
    uint64_t* cand;
    unsigned* cnt;
    float* thr;
    float* thr_sel;
    unsigned* sel_cnt;
    unsigned* flags;
    int rank;
    int partition;
    int cap;
};

// Radix select in shared memory: the candidates of one query (<= 4096 keys, 32 KB) are staged once; the bits on which
// the smallest and the largest score word differ are resolved 8 at a time (histogram with shared-memory atomics, the
// bin that holds the rank-th largest found by one warp), so the work is proportional to the number of candidates and a
// pass costs three barriers.  Scores of one query's candidates share their upper bits (same sign, 1-2 exponents): 3
// passes are typical.
__global__ void __launch_bounds__(SEL_NT) tc_select_kernel(const TcSelectParams p) {
    extern __shared__ uint64_t s_key[];  // [cap]
    __shared__ int s_hist[256];
    __shared__ uint32_t s_part[3][SEL_NT / 32];
    __shared__ int s_misc[4];  // [0] digit, [1] remaining rank, [2] survivors written, [3] others written
    const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();  // the next round's GEMM may set up and prefetch while the last blocks of this kernel run
    unsigned raw = p.cnt[q];
    if (raw > unsigned(p.cap)) {  // candidates were lost: the query goes to the exact scan in the end
        if (tid == 0) atomicOr(p.flags + q, 1u);
        raw = unsigned(p.cap);
    }
    const int n = int(raw);
    uint64_t* base = p.cand + size_t(q) * p.cap;
    uint32_t lmin = 0xFFFFFFFFu, lmax = 0u, real = 0u;
    for (int i = tid; i < n; i += SEL_NT) {
        const uint64_t key = base[i];
        s_key[i] = key;
        if (key != 0ull) {
            const uint32_t h = uint32_t(key >> 32);
            lmin = h < lmin ? h : lmin;
            lmax = h > lmax ? h : lmax;
            ++real;
        }
    }
    lmin = __reduce_min_sync(FULL, lmin);
    lmax = __reduce_max_sync(FULL, lmax);
    real = __reduce_add_sync(FULL, real);
    if (lane == 0) {
        s_part[0][warp] = lmin;
        s_part[1][warp] = lmax;
        s_part[2][warp] = real;
    }
    if (tid == 0) {
        s_misc[2] = 0;
        s_misc[3] = 0;
    }
    __syncthreads();
    uint32_t gmin = 0xFFFFFFFFu, gmax = 0u;
    int nv = 0;
#pragma unroll
    for (int w = 0; w < SEL_NT / 32; ++w) {
        gmin = s_part[0][w] < gmin ? s_part[0][w] : gmin;
        gmax = s_part[1][w] > gmax ? s_part[1][w] : gmax;
        nv += int(s_part[2][w]);
    }
    const bool select = nv >= p.rank;
    uint32_t T = 0;  // rank-th largest score word
    if (select) {
        const uint32_t diff = gmin ^ gmax;
        if (diff == 0u) {
            T = gmin;
        } else {
            int lo = 32 - __clz(diff);           // bits [0, lo) differ somewhere; the bits above are common
            uint64_t pfx = uint64_t(gmax) >> lo;  // value of the bits resolved so far
            int need = p.rank;
            while (lo > 0) {  // block-uniform
                const int w = lo < 8 ? lo : 8;
                lo -= w;
                for (int b = tid; b < 256; b += SEL_NT) s_hist[b] = 0;
                __syncthreads();
                for (int i = tid; i < n; i += SEL_NT) {
                    const uint64_t key = s_key[i];
                    const uint32_t h = uint32_t(key >> 32);
                    if (key != 0ull && (uint64_t(h) >> (lo + w)) == pfx) atomicAdd(&s_hist[(h >> lo) & ((1u << w) - 1u)], 1);
                }
                __syncthreads();
                if (warp == 0) {
                    int sm = 0;
#pragma unroll
                    for (int b = 0; b < 8; ++b) sm += s_hist[lane * 8 + b];
                    int S = sm;  // keys in the bins of lanes >= this one
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_down_sync(FULL, S, o);
                        if (lane + o < 32) S += v;
                    }
                    const int above = S - sm;
                    if (above < need && need <= S) {  // exactly one lane
                        int acc = above;
                        for (int b = lane * 8 + 7; b >= lane * 8; --b) {
                            const int hcount = s_hist[b];
                            if (acc + hcount >= need) {
                                s_misc[0] = b;
                                s_misc[1] = need - acc;
                                break;
                            }
                            acc += hcount;
                        }
                    }
                }
                __syncthreads();
                pfx = (pfx << w) | uint64_t(s_misc[0]);
                need = s_misc[1];
            }
            T = uint32_t(pfx);
        }
    }
    // survivors (score word >= T; ties included) to the front; in partition mode the other real candidates behind them.
    // Every key is in shared memory by now, so the buffer can be rewritten in place.
    for (int i = tid; i < n; i += SEL_NT) {
        const uint64_t key = s_key[i];
        if (key != 0ull && (!select || uint32_t(key >> 32) >= T)) base[atomicAdd(&s_misc[2], 1)] = key;
    }
    __syncthreads();
    const int total_k = s_misc[2];
    if (p.partition && select) {
        for (int i = tid; i < n; i += SEL_NT) {
            const uint64_t key = s_key[i];
            if (key != 0ull && uint32_t(key >> 32) < T) base[total_k + atomicAdd(&s_misc[3], 1)] = key;
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (p.partition) {
            p.thr_sel[q] = select ? unorder_f32(T) : p.thr[q];
            p.cnt[q] = unsigned(total_k + s_misc[3]);
            p.sel_cnt[q] = unsigned(total_k);
        } else {
            if (select) p.thr[q] = unorder_f32(T);
            p.cnt[q] = unsigned(total_k);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4: exact fp32 rescore of the candidates, top-k, two-level certificate, device-side fallback queue
// ------------------------------------------------------------------------------------------------
struct TcRescoreParams {
    const float* rows;      // fp32 row-swizzled shard
    const float* qn;        // [nq][128] normalised queries
    const float* q_raw;     // [nq][128] queries as given (copied to the fallback queue)
    const uint64_t* cand;   // [nq][cap] approximate keys: [0, sel_cnt) the k' best, [sel_cnt, cnt) the rest
    int cap;
    const unsigned* cnt;
    const unsigned* sel_cnt;
    const float* eps;       // [nq] certificate slack
    const float* thr;       // [nq] threshold of the last GEMM round
    const float* thr_sel;   // [nq] k'-th best approximate score (== thr when fewer than k' candidates)
    unsigned* flags;
    unsigned* n_flagged;    // length of the fallback queue
    int* fb_list;           // [nq] queued query indices
    float* fb_q;            // [nq][128] their raw queries
    int nq, k;
    uint32_t id_base;
    uint64_t* out_keys;
    float* out_scores;
    int64_t* out_ids;
};

// offer candidates [c_begin, c_end) of the query, rescored exactly, to the warp's top-k list
__device__ __forceinline__ void rescore_range(const TcRescoreParams& p, const uint64_t* cand, const float4* q4, int c_begin, int c_end,
                                              WarpTopK<4>& tk, int lane) {
    constexpr int U = 8;  // candidate rows in flight per warp
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        uint64_t batch = 0;  // lane i holds the exact key of candidate (c0 + i)
        for (int u0 = 0; u0 < 32 && c0 + u0 < c_end; u0 += U) {
            float part[U];
            int64_t rowid[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int c = c0 + u0 + u;
                part[u] = 0.f;
                rowid[u] = -1;
                if (c < c_end) {
                    const int64_t row = key_id(cand[c]);
                    rowid[u] = row;
                    if (row >= 0) {
                        const float4 v = reinterpret_cast<const float4*>(p.rows + row * DIM)[lane];  // physical chunk `lane`
                        const float4 w = q4[swz_chunk(lane, row)];                                    // = logical chunk
                        part[u] = fmaf(v.w, w.w, fmaf(v.z, w.z, fmaf(v.y, w.y, v.x * w.x)));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float sacc = part[u];
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) sacc += __shfl_xor_sync(FULL, sacc, o);
                if (rowid[u] >= 0 && lane == u0 + u) batch = make_key(sacc, p.id_base + uint32_t(rowid[u]));
            }
        }
        tk.offer(batch, lane, p.k);
    }
}

// the WPQ warps of a query hold one sorted list each: its first warp ends up with the k best of the union
template <int WPQ>
__device__ __forceinline__ void rescore_merge_block(WarpTopK<4>& tk, uint64_t (*s_lists)[128], int warp, int lane, int k) {
    if (WPQ == 1) return;
    if (warp > 0) tk.store(s_lists[warp - 1], lane, 128);
    __syncthreads();
    if (warp == 0) {
#pragma unroll 1
        for (int w = 0; w < WPQ - 1; ++w) {
            WarpTopK<4> other;
            other.template load<false>(s_lists[w], lane, 128);
            tk.merge_sorted(other.key, lane, k);
        }
    }
    __syncthreads();
}

// WPQ = 4: one block (4 warps) per query -- small batches, where the chain of dependent gathers of one query sets the
// kernel's duration (512 queries: 53 us instead of 114 us); WPQ = 1: one warp per query, four queries per block -- large
// batches, where the extra merges and barriers cost more than the parallelism buys (4096 queries: 137 us vs 209 us).
template <int WPQ>
__global__ void __launch_bounds__(128) tc_rescore_kernel(const TcRescoreParams p) {
    __shared__ uint64_t s_lists[WPQ > 1 ? WPQ - 1 : 1][128];
    __shared__ int s_phase_b_all[4];
    const int lane = threadIdx.x & 31;
    const int warp = WPQ == 1 ? 0 : int(threadIdx.x >> 5);   // warp index within the query's team
    const int q = WPQ == 1 ? int(blockIdx.x) * 4 + int(threadIdx.x >> 5) : int(blockIdx.x);
    if (q >= p.nq) return;  // WPQ == 1 only: whole warps leave, no block barrier follows for them
    int& s_phase_b = s_phase_b_all[WPQ == 1 ? (threadIdx.x >> 5) : 0];
    const int n_all = int(p.cnt[q]);
    const int n_sel = int(p.sel_cnt[q]);
    const unsigned fl = p.flags[q];
    const uint64_t* cand = p.cand + size_t(q) * p.cap;
    const float4* q4 = reinterpret_cast<const float4*>(p.qn + size_t(q) * DIM);
    const float eps = p.eps[q];
    WarpTopK<4> tk;
    tk.init();
    // Phase A.  Every row that is NOT among the k' selected candidates has an approximate score <= t' and therefore an
    // exact score <= t' + eps: the k best of the candidates are the k best of the shard if the k-th exact score beats that.
    // t' = -inf means every row of the shard is a candidate (shards of at most 4096 rows with fewer than k' rows).
    {
        const int per = ((n_sel + WPQ - 1) / WPQ + 31) & ~31;
        const int lo = warp * per < n_sel ? warp * per : n_sel;
        const int hi = lo + per < n_sel ? lo + per : n_sel;
        rescore_range(p, cand, q4, lo, hi, tk, lane);
    }
    rescore_merge_block<WPQ>(tk, s_lists, warp, lane, p.k);
    bool ok = false;
    if (warp == 0) {
        if (!(fl & 1u)) {
            const float tprime = p.thr_sel[q];
            ok = (tprime == -INFINITY) || (key_score(tk.thr) > tprime + eps);
        }
        if (lane == 0) s_phase_b = (!ok && !(fl & 1u) && n_all > n_sel) ? 1 : 0;
    }
    if (WPQ == 1) __syncwarp(); else __syncthreads();
    if (s_phase_b) {
        // Phase B: all the rows above the last round's threshold are in the buffer; rescore the rest of them too.
        if (warp > 0) tk.init();
        const int n_rest = n_all - n_sel;
        const int per = ((n_rest + WPQ - 1) / WPQ + 31) & ~31;
        const int lo = n_sel + (warp * per < n_rest ? warp * per : n_rest);
        const int hi = lo + per < n_all ? lo + per : n_all;
        rescore_range(p, cand, q4, lo, hi, tk, lane);
        rescore_merge_block<WPQ>(tk, s_lists, warp, lane, p.k);
        if (warp == 0) {
            const float tround = p.thr[q];
            ok = (tround == -INFINITY) || (key_score(tk.thr) > tround + eps);
        }
    }
    if (warp != 0) return;
    if (!ok) {  // queue the query for the exact scan (device-side: the scan kernels read the queue length themselves)
        unsigned slot = 0;
        if (lane == 0) {
            slot = atomicAdd(p.n_flagged, 1u);
            p.fb_list[slot] = q;
            p.flags[q] = fl | 2u;
        }
        slot = __shfl_sync(FULL, slot, 0);
        reinterpret_cast<float4*>(p.fb_q + size_t(slot) * DIM)[lane] = reinterpret_cast<const float4*>(p.q_raw + size_t(q) * DIM)[lane];
    }
    const size_t base = size_t(q) * p.k;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = j * 32 + lane;
        if (r < p.k) {
            const uint64_t key = tk.key[j];
            p.out_keys[base + r] = key;
            if (p.out_scores) p.out_scores[base + r] = key_score(key);
            if (p.out_ids) p.out_ids[base + r] = key_id(key);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4 for k > 128 (up to FCS_MAX_K): one block per query.  The exact keys of the candidates are written to shared
// memory and sorted there (bitonic, descending); same two-phase certificate, same fallback queue.
// ------------------------------------------------------------------------------------------------
constexpr int RB_NT = 256;

__device__ __forceinline__ void rb_exact_keys(const TcRescoreParams& p, const uint64_t* cand, const float4* q4, int c_begin, int c_end,
                                              uint64_t* s_keys, int warp, int lane) {
    constexpr int U = 4;  // candidate rows in flight per warp
    for (int c0 = c_begin + warp * U; c0 < c_end; c0 += (RB_NT / 32) * U) {
        float part[U];
        int64_t rowid[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            part[u] = 0.f;
            rowid[u] = -1;
            if (c0 + u < c_end) {
                const int64_t row = key_id(cand[c0 + u]);
                rowid[u] = row;
                if (row >= 0) {
                    const float4 v = reinterpret_cast<const float4*>(p.rows + row * DIM)[lane];
                    const float4 w = q4[swz_chunk(lane, row)];
                    part[u] = fmaf(v.w, w.w, fmaf(v.z, w.z, fmaf(v.y, w.y, v.x * w.x)));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float sacc = part[u];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) sacc += __shfl_xor_sync(FULL, sacc, o);
            if (lane == 0 && c0 + u < c_end) s_keys[c0 + u] = rowid[u] >= 0 ? make_key(sacc, p.id_base + uint32_t(rowid[u])) : 0ull;
        }
    }
}

// sorts s_keys[0, n) descending (pads to a power of two with empty keys); all RB_NT threads call
__device__ __forceinline__ void rb_sort_desc(uint64_t* s_keys, int n, int tid) {
    int P = 2;
    while (P < n) P <<= 1;
    for (int i = n + tid; i < P; i += RB_NT) s_keys[i] = 0ull;
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (P >> 1); i += RB_NT) {
                const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const uint64_t a = s_keys[lo], b = s_keys[hi];
                if ((a < b) == desc) {
                    s_keys[lo] = b;
                    s_keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(RB_NT) tc_rescore_big_kernel(const TcRescoreParams p) {
    extern __shared__ uint64_t s_keys[];  // [cap]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = blockIdx.x;
    const int n_all = int(p.cnt[q]);
    const int n_sel = int(p.sel_cnt[q]);
    const unsigned fl = p.flags[q];
    const uint64_t* cand = p.cand + size_t(q) * p.cap;
    const float4* q4 = reinterpret_cast<const float4*>(p.qn + size_t(q) * DIM);
    const float eps = p.eps[q];
    // Phase A (see tc_rescore_kernel): the k' best approximate candidates
    rb_exact_keys(p, cand, q4, 0, n_sel, s_keys, warp, lane);
    __syncthreads();
    rb_sort_desc(s_keys, n_sel, tid);
    int n_have = n_sel;
    bool ok = false;
    if (!(fl & 1u)) {
        const float tprime = p.thr_sel[q];
        const float sk = n_have >= p.k ? key_score(s_keys[p.k - 1]) : -INFINITY;
        ok = (tprime == -INFINITY) || (sk > tprime + eps);
        if (!ok && n_all > n_sel) {  // Phase B: every row above the sweep threshold (block-uniform decision)
            __syncthreads();
            rb_exact_keys(p, cand, q4, n_sel, n_all, s_keys, warp, lane);
            __syncthreads();
            rb_sort_desc(s_keys, n_all, tid);
            n_have = n_all;
            const float tround = p.thr[q];
            const float sk2 = n_have >= p.k ? key_score(s_keys[p.k - 1]) : -INFINITY;
            ok = (tround == -INFINITY) || (sk2 > tround + eps);
        }
    }
    if (!ok && warp == 0) {
        unsigned slot = 0;
        if (lane == 0) {
            slot = atomicAdd(p.n_flagged, 1u);
            p.fb_list[slot] = q;
            p.flags[q] = fl | 2u;
        }
        slot = __shfl_sync(FULL, slot, 0);
        reinterpret_cast<float4*>(p.fb_q + size_t(slot) * DIM)[lane] = reinterpret_cast<const float4*>(p.q_raw + size_t(q) * DIM)[lane];
    }
    const size_t base = size_t(q) * p.k;
    for (int r = tid; r < p.k; r += RB_NT) {
        const uint64_t key = r < n_have ? s_keys[r] : 0ull;
        p.out_keys[base + r] = key;
        if (p.out_scores) p.out_scores[base + r] = key_score(key);
        if (p.out_ids) p.out_ids[base + r] = key_id(key);
    }
}

thread_local std::string g_tc_error;
int tc_fail(int code, const char* what, cudaError_t e) {
    g_tc_error = std::string(what) + ": " + cudaGetErrorString(e);
    (void)cudaGetLastError();
    return code;
}
#define TC_CUDA(call)                                                                                  \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) return tc_fail(e__ == cudaErrorMemoryAllocation ? FCS_ERR_NOMEM : FCS_ERR_CUDA, #call, e__); \
    } while (0)

double env_double(const char* name, double dflt, double lo, double hi) {
    if (const char* s = getenv(name)) {
        const double v = atof(s);
        if (v >= lo && v <= hi) return v;
    }
    return dflt;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// round plan (host)
// ------------------------------------------------------------------------------------------------
struct TcRound {
    int64_t n_idx, j0, stride, comp_T;
    int first;
    int rank;       // selection rank after the round
    int partition;  // last round: partition instead of compact
};

struct TcState {
    int device = 0, sm_count = 0;
    const float* rows = nullptr;
    int64_t n_rows = 0, n_tiles = 0;
    uint32_t id_base = 0;
    uint8_t* b_img = nullptr;
    float max_rhat = 1.f, max_dr = 0.f;
    // per-search workspace, grown on demand
    int nq_cap = 0;
    size_t cand_elems = 0;  // keys the candidate buffer holds
    float* qn = nullptr;
    uint8_t* a_img = nullptr;
    float* thr = nullptr;
    float* thr_sel = nullptr;
    unsigned* cnt = nullptr;
    unsigned* sel_cnt = nullptr;
    uint64_t* cand = nullptr;
    unsigned* flags = nullptr;
    float* eps = nullptr;
    int* fb_list = nullptr;
    float* fb_q = nullptr;
    unsigned* n_flagged = nullptr;    // device: [0] fallback queue length, [1..2] build statistics
    unsigned* h_n_flagged = nullptr;  // pinned copy of [0], refreshed at the end of every search
    // plan parameters (FCS_TC_* environment overrides, for sweeps)
    double cf_mult = 2.4, cf_min = 384.0, m_final = 24.0, c_sample = 128.0, m_min = 16.0;
    bool verbose = false;
    // timing of the dominant kernel: one event pair per K3 launch of the last search
    static constexpr int MAX_ROUNDS = 16;
    cudaEvent_t ev[2 * MAX_ROUNDS] = {};
    cudaEvent_t ev_done = nullptr;
    int last_rounds = 0;
    bool last_valid = false;
    bool last_timed = false;
    // FCS_TC_PHASES=1: an event in front of every kernel of a search, printed by tc_last_kernel_ms (diagnostics)
    bool phases = false;
    int r0_tiles = TC_R0_TILES;
    bool r0_auto = true;
    bool pdl_on = true;      // FCS_TC_PDL=0 turns the programmatic launches off (A/B)
    bool timing_on = false;  // event pairs around the GEMM launches (fcs_set_profiling): they cost the PDL overlap
    std::vector<cudaEvent_t> pev;
    std::vector<std::string> plabel;
    int n_pev = 0;
};

static void tc_phase(TcState* s, const char* label, cudaStream_t stream) {
    if (!s->phases) return;
    if (s->n_pev >= int(s->pev.size())) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        s->pev.push_back(e);
        s->plabel.emplace_back();
    }
    s->plabel[s->n_pev] = label;
    cudaEventRecord(s->pev[s->n_pev++], stream);
}

static std::vector<TcRound> tc_plan(const TcState* s, int kp, int n_qgroups) {
    std::vector<TcRound> plan;
    const int64_t nt = s->n_tiles;
    if (nt <= TC_R0_TILES) {  // the whole shard fits the candidate buffer: one round records everything
        plan.push_back({nt, 0, 1, 0, 1, kp, 1});
        return plan;
    }
    // round 0 records every score: keep it to one wave of CTAs (query groups x tiles <= SMs) for large batches
    int r0 = s->r0_tiles;
    if (s->r0_auto) r0 = n_qgroups <= 4 ? 32 : (n_qgroups <= 9 ? 16 : 8);
    const int64_t t0 = nt / 2 < r0 ? nt / 2 : r0;
    const double cf = std::fmax(s->cf_mult * kp, s->cf_min);  // rows expected above the sweep threshold
    int64_t t_last = int64_t(std::ceil(double(nt) * s->m_final / cf));
    if (t_last > nt / 2) t_last = nt / 2;
    if (t_last < t0) t_last = t0;
    const int64_t stride = nt / t_last;  // >= 2
    auto clamp_rank = [](double r) { return int(r < 1.0 ? 1.0 : (r > double(TC_MAX_RANK) ? double(TC_MAX_RANK) : std::ceil(r))); };
    // hits cost more than launches even for small batches (512 queries: 5 rounds of 128 hits beat 4 rounds of 512 hits)
    const double c_sample = s->c_sample;
    // nested samples t0 < T_1 < ... < t_last, each at most c_sample/m_min times the previous
    std::vector<int64_t> ts{t0};
    if (t_last > t0) {
        const double ratio = double(t_last) / double(t0);
        int ns = int(std::ceil(std::log(ratio) / std::log(c_sample / s->m_min) - 1e-9));
        if (ns < 1) ns = 1;
        for (int i = 1; i <= ns; ++i) {
            int64_t t = (i == ns) ? t_last : int64_t(std::llround(double(t0) * std::pow(ratio, double(i) / ns)));
            if (t <= ts.back()) t = ts.back() + 1;
            if (t > t_last) t = t_last;
            if (t > ts.back()) ts.push_back(t);
        }
    }
    for (size_t i = 0; i < ts.size(); ++i) {
        TcRound r = {};
        r.j0 = i == 0 ? 0 : ts[i - 1];
        r.n_idx = ts[i] - r.j0;
        r.stride = stride;
        r.first = i == 0 ? 1 : 0;
        const double seen = double(ts[i]);
        const double next = (i + 1 < ts.size()) ? double(ts[i + 1] - ts[i]) : double(nt - ts[i]);
        const double target = (i + 1 < ts.size()) ? c_sample : cf;
        r.rank = clamp_rank(target * seen / next);
        plan.push_back(r);
    }
    TcRound sweep = {};
    sweep.n_idx = nt - t_last;
    sweep.stride = stride;
    sweep.comp_T = t_last;
    sweep.rank = kp;
    sweep.partition = 1;
    plan.push_back(sweep);
    return plan;
}

void tc_phase_mark(TcState* s, const char* label, cudaStream_t stream) {
    if (s) tc_phase(s, label, stream);
}

const char* tc_last_error() { return g_tc_error.c_str(); }
int tc_min_batch() { return GEMV_MAX_NQ + 1; }  // one exact-scan pass always wins
int tc_max_k() { return FCS_MAX_K; }
int tc_last_rounds(const TcState* s) { return s ? s->last_rounds : 0; }
uint64_t tc_image_bytes(const TcState* s) { return s ? uint64_t(s->n_tiles) * B_TILE_BYTES : 0; }

void tc_set_timing(TcState* s, bool on) {
    if (s) s->timing_on = on;
}

float tc_last_kernel_ms(TcState* s) {
    if (!s || !s->last_valid || !s->last_timed) return 0.f;
    if (cudaEventSynchronize(s->ev_done) != cudaSuccess) return 0.f;
    float total = 0.f;
    if (s->phases) {
        for (int i = 0; i + 1 < s->n_pev; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, s->pev[i], s->pev[i + 1]);
            fprintf(stderr, "[fcs_tc] phase %-18s %8.1f us\n", s->plabel[i].c_str(), ms * 1e3f);
        }
    }
    for (int r = 0; r < s->last_rounds && r < TcState::MAX_ROUNDS; ++r) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s->ev[2 * r], s->ev[2 * r + 1]) == cudaSuccess) total += ms;
        if (s->verbose) fprintf(stderr, "[fcs_tc] round %d: gemm+filter %.3f ms\n", r, ms);
    }
    return total;
}

int tc_last_flagged(TcState* s) {
    if (!s || !s->last_valid) return 0;
    if (cudaEventSynchronize(s->ev_done) != cudaSuccess) return 0;
    return int(s->h_n_flagged[0]);
}

static void tc_free_workspace(TcState* s) {
    cudaFree(s->qn); cudaFree(s->a_img); cudaFree(s->thr); cudaFree(s->thr_sel); cudaFree(s->cnt); cudaFree(s->sel_cnt);
    cudaFree(s->cand); cudaFree(s->flags); cudaFree(s->eps); cudaFree(s->fb_list); cudaFree(s->fb_q);
    s->qn = nullptr; s->a_img = nullptr; s->thr = nullptr; s->thr_sel = nullptr; s->cnt = nullptr; s->sel_cnt = nullptr;
    s->cand = nullptr; s->flags = nullptr; s->eps = nullptr; s->fb_list = nullptr; s->fb_q = nullptr;
    s->nq_cap = 0;
    s->cand_elems = 0;
}

void tc_destroy(TcState* s) {
    if (!s) return;
    tc_free_workspace(s);
    for (cudaEvent_t e : s->ev)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : s->pev) cudaEventDestroy(e);
    if (s->ev_done) cudaEventDestroy(s->ev_done);
    cudaFree(s->b_img);
    cudaFree(s->n_flagged);
    if (s->h_n_flagged) cudaFreeHost(s->h_n_flagged);
    (void)cudaGetLastError();
    delete s;
}

int tc_create(TcState** out, int device, int sm_count, const float* rows, int64_t n_rows, uint32_t id_base, cudaStream_t stream) {
    *out = nullptr;
    TcState* s = new TcState();
    s->device = device;
    s->sm_count = sm_count;
    s->rows = rows;
    s->n_rows = n_rows;
    s->id_base = id_base;
    s->n_tiles = (n_rows + TC_N - 1) / TC_N;
    s->verbose = getenv("FCS_TC_VERBOSE") != nullptr;
    s->cf_mult = env_double("FCS_TC_CF_MULT", s->cf_mult, 1.0, 16.0);
    s->cf_min = env_double("FCS_TC_CF_MIN", s->cf_min, 32.0, 2048.0);
    s->m_final = env_double("FCS_TC_M_FINAL", s->m_final, 2.0, 256.0);
    s->c_sample = env_double("FCS_TC_C_SAMPLE", s->c_sample, 64.0, 2048.0);
    s->m_min = env_double("FCS_TC_M_MIN", s->m_min, 1.0, 64.0);
    s->r0_tiles = int(env_double("FCS_TC_R0_TILES", TC_R0_TILES, 1.0, TC_R0_TILES));
    s->r0_auto = getenv("FCS_TC_R0_TILES") == nullptr;
    if (const char* e = getenv("FCS_TC_PDL")) s->pdl_on = atoi(e) != 0;
    s->phases = getenv("FCS_TC_PHASES") != nullptr;
    auto run = [&]() -> int {
        TC_CUDA(cudaFuncSetAttribute(tc_gemm_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        // the small kernels between two GEMM launches ask for the same shared-memory carve-out as the GEMM kernel,
        // so the SMs are not reconfigured (drained) twice per round
        TC_CUDA(cudaFuncSetAttribute(tc_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_CAP_BIG * 8));
        TC_CUDA(cudaFuncSetAttribute(tc_rescore_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_CAP_BIG * 8));
        TC_CUDA(cudaFuncSetAttribute(tc_select_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TC_CUDA(cudaFuncSetAttribute(tc_rescore_big_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TC_CUDA(cudaFuncSetAttribute(tc_prep_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TC_CUDA(cudaFuncSetAttribute(tc_rescore_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TC_CUDA(cudaFuncSetAttribute(tc_rescore_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        for (cudaEvent_t& e : s->ev) TC_CUDA(cudaEventCreate(&e));
        TC_CUDA(cudaEventCreateWithFlags(&s->ev_done, cudaEventDisableTiming));
        TC_CUDA(cudaMalloc(&s->b_img, size_t(s->n_tiles) * B_TILE_BYTES));
        TC_CUDA(cudaMalloc(&s->n_flagged, 4 * sizeof(unsigned)));
        TC_CUDA(cudaMallocHost(&s->h_n_flagged, 4 * sizeof(unsigned)));
        TC_CUDA(cudaMemsetAsync(s->n_flagged, 0, 4 * sizeof(unsigned), stream));
        const int64_t total = s->n_tiles * TC_N * 16;
        int64_t g = (total + 255) / 256;
        if (g > sm_count * 16) g = sm_count * 16;
        tc_build_bimg_kernel<<<int(g), 256, 0, stream>>>(rows, n_rows, s->n_tiles, s->b_img, s->n_flagged + 1);
        TC_CUDA(cudaGetLastError());
        TC_CUDA(cudaMemcpyAsync(s->h_n_flagged, s->n_flagged, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
        TC_CUDA(cudaStreamSynchronize(stream));
        float n2, d2;
        memcpy(&n2, &s->h_n_flagged[1], 4);
        memcpy(&d2, &s->h_n_flagged[2], 4);
        s->max_rhat = sqrtf(n2) * 1.0001f;
        s->max_dr = sqrtf(d2) * 1.0001f;
        s->h_n_flagged[0] = 0;
        return FCS_OK;
    };
    const int rc = run();
    if (rc != FCS_OK) {
        tc_destroy(s);
        return rc;
    }
    *out = s;
    return FCS_OK;
}

static int tc_ensure_workspace(TcState* s, int nq, int cap) {
    const int nq_pad_want = ((nq + TC_QGROUP - 1) / TC_QGROUP) * TC_QGROUP;
    if (nq <= s->nq_cap && size_t(nq_pad_want) * size_t(cap) <= s->cand_elems) return FCS_OK;
    if (size_t(nq_pad_want) * size_t(cap) * 8 > (size_t(24) << 30)) {
        g_tc_error = "candidate buffers of this batch would exceed 24 GB: split the batch (large k with a very large batch)";
        return FCS_ERR_NOMEM;
    }
    if (nq <= s->nq_cap) {  // only the candidate buffer has to grow (a larger k than before)
        cudaFree(s->cand);
        s->cand = nullptr;
        s->cand_elems = 0;
        TC_CUDA(cudaMalloc(&s->cand, size_t(s->nq_cap) * size_t(cap) * 8));
        s->cand_elems = size_t(s->nq_cap) * size_t(cap);
        return FCS_OK;
    }
    tc_free_workspace(s);
    const int nq_pad = nq_pad_want;
    TC_CUDA(cudaMalloc(&s->qn, size_t(nq_pad) * DIM * 4));
    TC_CUDA(cudaMalloc(&s->a_img, size_t(nq_pad / TC_M) * A_TILE_BYTES));
    TC_CUDA(cudaMalloc(&s->thr, size_t(nq_pad) * 4));
    TC_CUDA(cudaMalloc(&s->thr_sel, size_t(nq_pad) * 4));
    TC_CUDA(cudaMalloc(&s->cnt, size_t(nq_pad) * 4));
    TC_CUDA(cudaMalloc(&s->sel_cnt, size_t(nq_pad) * 4));
    TC_CUDA(cudaMalloc(&s->flags, size_t(nq_pad) * 4));
    TC_CUDA(cudaMalloc(&s->eps, size_t(nq_pad) * 4));
    TC_CUDA(cudaMalloc(&s->fb_list, size_t(nq_pad) * 4));
    TC_CUDA(cudaMalloc(&s->fb_q, size_t(nq_pad) * DIM * 4));
    TC_CUDA(cudaMalloc(&s->cand, size_t(nq_pad) * size_t(cap) * 8));
    s->cand_elems = size_t(nq_pad) * size_t(cap);
    s->nq_cap = nq_pad;
    return FCS_OK;
}

int tc_default_kprime(int k) {
    // margin for the exactness certificate: eps (~0.0035 for unit vectors) covers a few dozen ranks at TED scale;
    // 60 % of k for small k, 256 + k/4 for large k (the ranks get denser further down the list)
    int margin = k * 6 / 10;
    if (margin > 256 + k / 4) margin = 256 + k / 4;
    if (margin < 32) margin = 32;
    int kp = (k + margin + 31) / 32 * 32;
    return kp > TC_MAX_KPRIME ? TC_MAX_KPRIME : kp;
}

static void tc_launch_prep(TcState* s, const float* q_dev, int nq, int nq_pad, int qnorm, unsigned first_slots, cudaStream_t stream) {
    TcPrepParams pp = {};
    pp.q_raw = q_dev; pp.nq = nq; pp.nq_pad = nq_pad; pp.qnorm = qnorm;
    pp.qn = s->qn; pp.a_img = s->a_img; pp.thr = s->thr; pp.thr_sel = s->thr_sel; pp.cnt = s->cnt; pp.sel_cnt = s->sel_cnt;
    pp.flags = s->flags; pp.eps = s->eps; pp.n_flagged = s->n_flagged; pp.first_round_slots = first_slots;
    pp.max_rhat = s->max_rhat; pp.max_dr = s->max_dr;
    tc_prep_kernel<<<(nq_pad + 3) / 4, 128, 0, stream>>>(pp);
}

static void tc_launch_gemm(TcState* s, const TcRound& r, int nq, int n_qgroups, int cap, cudaStream_t stream, int trace_on, bool pdl) {
    TcGemmParams gp = {};
    gp.a_img = s->a_img; gp.b_img = s->b_img; gp.n_rows = s->n_rows; gp.nq = nq; gp.n_qgroups = n_qgroups;
    gp.n_idx = r.n_idx; gp.j0 = r.j0; gp.stride = r.stride; gp.comp_T = r.comp_T;
    gp.thr = s->thr; gp.cnt = s->cnt; gp.cand = s->cand; gp.cap = cap; gp.first_round = r.first; gp.trace_on = trace_on;
    const int64_t steps = int64_t(n_qgroups) * r.n_idx;
    const int grid = int(steps < s->sm_count ? steps : s->sm_count);
    // every thread that appends at all rounds its reservation up to res_block slots: keep the waste of the
    // ~grid/n_qgroups segments that share a query below ~512 slots
    const int segs = (grid + n_qgroups - 1) / n_qgroups + 1;
    const int rb = 512 / segs;
    gp.res_block = rb < 2 ? 2 : (rb > TC_RES_MAX ? TC_RES_MAX : rb);
    gp.pdl = pdl ? 1 : 0;
    if (!pdl) {
        tc_gemm_filter_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(gp);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = TC_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    (void)cudaLaunchKernelEx(&cfg, tc_gemm_filter_kernel, gp);
}

// Enqueues the whole batched search on `stream` and returns without synchronising.  Queries whose certificate failed
// are queued on the device (fallback queue: *fb_count_dev entries of fb_list_dev / fb_q_dev); the caller launches the
// exact scan over that queue.
int tc_search(TcState* s, const float* q_dev, int nq, int k, int kprime, int qnorm, float* out_scores, int64_t* out_ids,
              uint64_t* out_keys, cudaStream_t stream, int* launches, TcFallbackQueue* fbq) {
    if (k > tc_max_k()) {
        g_tc_error = "k too large for the tensor-core path";
        return FCS_ERR_UNSUPPORTED;
    }
    int kp = kprime > 0 ? kprime : tc_default_kprime(k);
    if (kp < k) kp = k;
    if (kp > TC_MAX_KPRIME) kp = TC_MAX_KPRIME;
    // candidate slots per query: the sweep leaves about cf rows above its threshold (+- 50 %), plus the kept sample ranks
    // and the unused ends of slot reservations
    const double cf = std::fmax(s->cf_mult * kp, s->cf_min);
    const int cap = (1.6 * cf + 600.0 > double(TC_CAP)) ? TC_CAP_BIG : TC_CAP;
    int rc = tc_ensure_workspace(s, nq, cap);
    if (rc != FCS_OK) return rc;
    const int n_qgroups = (nq + TC_QGROUP - 1) / TC_QGROUP;
    const int nq_pad = n_qgroups * TC_QGROUP;
    const std::vector<TcRound> plan = tc_plan(s, kp, n_qgroups);

    if (s->phases && s->n_pev > 0) {  // keep the previous search's last event: the gap between two searches is a phase too
        std::swap(s->pev[0], s->pev[s->n_pev - 1]);
        s->plabel[0] = "(between searches)";
        s->n_pev = 1;
    }
    tc_phase(s, "prep", stream);
    tc_launch_prep(s, q_dev, nq, nq_pad, qnorm, unsigned(plan[0].n_idx * TC_N), stream);
    TC_CUDA(cudaGetLastError());
    ++*launches;

    static const int trace_round = getenv("FCS_TC_TRACE_ROUND") ? atoi(getenv("FCS_TC_TRACE_ROUND")) : -1;
    int rounds = 0;
    for (const TcRound& r : plan) {
        const bool timed = s->timing_on && rounds < TcState::MAX_ROUNDS;
        tc_phase(s, r.first ? "gemm-dump" : (r.comp_T ? "gemm-sweep" : "gemm-sample"), stream);
        if (timed) TC_CUDA(cudaEventRecord(s->ev[2 * rounds], stream));
        // rounds >= 1 follow a selection kernel and overlap its tail (programmatic dependent launch) -- unless events are
        // recorded in between (profiling, phase diagnostics), which serialises the two kernels again
        tc_launch_gemm(s, r, nq, n_qgroups, cap, stream, trace_round == rounds ? 1 : 0, rounds > 0 && s->pdl_on && !s->timing_on && !s->phases);
        TC_CUDA(cudaGetLastError());
        if (timed) TC_CUDA(cudaEventRecord(s->ev[2 * rounds + 1], stream));
        tc_phase(s, "select", stream);
        TcSelectParams sp = {s->cand, s->cnt, s->thr, s->thr_sel, s->sel_cnt, s->flags, r.rank, r.partition, cap};
        tc_select_kernel<<<nq, SEL_NT, size_t(cap) * 8, stream>>>(sp);
        TC_CUDA(cudaGetLastError());
        *launches += 2;
        if (s->verbose)
            fprintf(stderr, "[fcs_tc] round %d: %s tiles=%lld j0=%lld stride=%lld comp_T=%lld rank=%d%s\n", rounds,
                    r.first ? "dump" : (r.comp_T ? "sweep" : "sample"), (long long)r.n_idx, (long long)r.j0, (long long)r.stride,
                    (long long)r.comp_T, r.rank, r.partition ? " (partition)" : "");
        ++rounds;
    }

    TcRescoreParams rp = {};
    rp.rows = s->rows; rp.qn = s->qn; rp.q_raw = q_dev; rp.cand = s->cand; rp.cnt = s->cnt; rp.sel_cnt = s->sel_cnt;
    rp.eps = s->eps; rp.thr = s->thr; rp.thr_sel = s->thr_sel; rp.flags = s->flags; rp.n_flagged = s->n_flagged;
    rp.fb_list = s->fb_list; rp.fb_q = s->fb_q; rp.nq = nq; rp.k = k; rp.id_base = s->id_base; rp.cap = cap;
    rp.out_keys = out_keys; rp.out_scores = out_scores; rp.out_ids = out_ids;
    tc_phase(s, "rescore", stream);
    if (k > TC_WARP_K) tc_rescore_big_kernel<<<nq, RB_NT, size_t(cap) * 8, stream>>>(rp);
    else if (nq <= 2048) tc_rescore_kernel<4><<<nq, 128, 0, stream>>>(rp);
    else tc_rescore_kernel<1><<<(nq + 3) / 4, 128, 0, stream>>>(rp);
    TC_CUDA(cudaGetLastError());
    ++*launches;
    tc_phase(s, "tail", stream);
    TC_CUDA(cudaMemcpyAsync(s->h_n_flagged, s->n_flagged, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    TC_CUDA(cudaEventRecord(s->ev_done, stream));
    s->last_rounds = rounds;
    s->last_valid = true;
    s->last_timed = s->timing_on;
    fbq->count_dev = s->n_flagged;
    fbq->list_dev = s->fb_list;
    fbq->q_dev = s->fb_q;
    fbq->count_host = s->h_n_flagged;
#ifdef FCS_TC_TRACE
    {
        TC_CUDA(cudaStreamSynchronize(stream));
        static long long h[8][256];
        cudaMemcpyFromSymbol(h, g_trace, sizeof h);
        fprintf(stderr, "[fcs_tc trace, CTA 0, tile 0; cycles relative to iteration 100]\n"
                        " it | producer: stage free | MMA: B tile landed, accumulator free, 8 MMAs issued | epilogue (warp 5): accumulator full, drained, after appends\n");
        for (int i = 100; i < 124; ++i)
            fprintf(stderr, "%3d | %8lld | %8lld %8lld %8lld | %8lld %8lld %8lld\n", i, h[5][i] - h[0][100], h[6][i] - h[0][100], h[0][i] - h[0][100],
                    h[1][i] - h[0][100], h[2][i] - h[0][100], h[3][i] - h[0][100], h[4][i] - h[0][100]);
    }
#endif
    return FCS_OK;
}

// Test hook: the round plan for a shard of n_rows rows and a given k' (what tc_search would launch).
int tc_debug_plan(int64_t n_rows, int kprime, int n_qgroups, int64_t* out, int max_rounds) {
    TcState s;
    s.n_rows = n_rows;
    s.n_tiles = (n_rows + TC_N - 1) / TC_N;
    const std::vector<TcRound> plan = tc_plan(&s, kprime, n_qgroups);
    int n = 0;
    for (const TcRound& r : plan) {
        if (n >= max_rounds) break;
        int64_t* o = out + size_t(n) * 7;
        o[0] = r.n_idx; o[1] = r.j0; o[2] = r.stride; o[3] = r.comp_T; o[4] = r.first; o[5] = r.rank; o[6] = r.partition;
        ++n;
    }
    return n;
}

// Test hook: the DB tile a round visits at position idx (host mirror of tile_of, for the plan tests).
int64_t tc_debug_tile_of(int64_t j0, int64_t stride, int64_t comp_T, int64_t idx) {
    TcGemmParams p = {};
    p.j0 = j0; p.stride = stride; p.comp_T = comp_T;
    if (p.comp_T == 0) return (p.j0 + idx) * p.stride;
    const int64_t gap = p.stride - 1, body = p.comp_T * gap;
    if (idx < body) {
        const int64_t b = idx / gap;
        return b * p.stride + 1 + (idx - b * gap);
    }
    return idx + p.comp_T;
}

// Test hook: approximate (bf16 tensor-core) scores of every (query, row) pair, for shards of at most
// TC_CAP rows.  One K3 launch in round-0 mode over all tiles records every score.
int tc_debug_approx(TcState* s, const float* q_dev, int nq, int qnorm, float* out_host, cudaStream_t stream) {
    if (s->n_rows > TC_CAP) {
        g_tc_error = "tc_debug_approx: shard larger than the candidate buffer";
        return FCS_ERR_UNSUPPORTED;
    }
    int rc = tc_ensure_workspace(s, nq, TC_CAP);
    if (rc != FCS_OK) return rc;
    const int n_qgroups = (nq + TC_QGROUP - 1) / TC_QGROUP;
    const int nq_pad = n_qgroups * TC_QGROUP;
    tc_launch_prep(s, q_dev, nq, nq_pad, qnorm, unsigned(s->n_tiles * TC_N), stream);
    TC_CUDA(cudaGetLastError());
    const TcRound r0 = {s->n_tiles, 0, 1, 0, 1, 1, 0};
    tc_launch_gemm(s, r0, nq, n_qgroups, TC_CAP, stream, 0, false);
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaStreamSynchronize(stream));
    std::string keys(size_t(nq) * TC_CAP * 8, '\0');
    TC_CUDA(cudaMemcpy(&keys[0], s->cand, keys.size(), cudaMemcpyDeviceToHost));
    const uint64_t* kk = reinterpret_cast<const uint64_t*>(keys.data());
    for (size_t i = 0; i < size_t(nq) * s->n_rows; ++i) out_host[i] = NAN;
    for (int q = 0; q < nq; ++q) {
        for (int64_t i = 0; i < s->n_tiles * TC_N; ++i) {
            const uint64_t key = kk[size_t(q) * TC_CAP + i];
            const int64_t row = key_id(key);
            if (row >= 0 && row < s->n_rows) out_host[size_t(q) * s->n_rows + row] = key_score(key);
        }
    }
    return FCS_OK;
}

}  // namespace fcs
