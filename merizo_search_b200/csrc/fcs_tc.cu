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
//          umma_two_issuers.cu, and one issuer per tile keeps a slow epilogue from stalling the other tiles).  Measured on B200 (scripts/micro/umma_rate.cu): one tcgen05.mma costs
//          >= 70 cycles whatever its N, so N=64 caps the tensor pipe at 46 % and N=128 reaches 90 %
//          (N=256: 100 %); A from TMEM instead of shared memory changes nothing.  Accumulators: fp32 in
//          TMEM, one 128-column buffer per query tile = all 512 columns; the epilogue of tile t drains
//          its buffer while the MMAs of the other three tiles run.
//        * warps 5..20: epilogue, one THREAD PER QUERY (tcgen05.ld 32x32b: lane = accumulator row).
//          Each thread reduces its scores 32 at a time with 3-input max, compares the group maximum
//          with the query's current threshold and only on a hit scans the 8-column sub-groups and
//          appends (score,row) keys to the query's candidate buffer in HBM (slots are reserved 8 at a
//          time with one atomic; the first round, where every row qualifies, indexes by row instead).
//        The threshold is a per-query constant during a launch; it is the k'-th best approximate
//        score over the rows seen so far.  The DB is therefore swept in ROUNDS of geometrically
//        growing size with a small selection kernel (tc_select_warp_kernel) in between, which keeps the
//        number of appends at about k' per round and the epilogue on its fast path.
//   K4  tc_rescore_kernel -- one warp per query gathers the k' candidate rows (fp32), recomputes the
//        inner products exactly, keeps the k best, and checks the exactness certificate
//        (k-th exact score) > (k'-th approximate score) + eps, eps = rigorous bf16 rounding bound.
//        Queries that fail it (or overflowed their buffer) are re-run on the exact fp32 scan (K2).
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <string>

#include "fcs_common.cuh"
#include "fcs_internal.h"
#include "fcs_tc.h"

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
constexpr int TC_CAP = 4096;                     // candidate slots per query
constexpr int TC_SMEM = TC_QT * A_TILE_BYTES + TC_STAGES * B_TILE_BYTES + 32 * 8 + 16;
constexpr int TC_MAX_KPRIME = 512;

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

// ------------------------------------------------------------------------------------------------
// one-time: fp32 (row-swizzled) rows -> bf16 operand images; also max row norm (for the certificate)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tc_build_bimg_kernel(const float* __restrict__ rows, int64_t n_rows, int64_t n_tiles,
                                                            uint8_t* __restrict__ b_img, unsigned* __restrict__ max_norm2_bits) {
    const int64_t total = n_tiles * TC_N * 16;  // one thread per (row, 8-element k-chunk)
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int k8 = int(t & 15);
        const int64_t row = t >> 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (row < n_rows) {
            const float4* src = reinterpret_cast<const float4*>(rows + row * DIM);
            a = src[swz_chunk(2 * k8, row)];
            b = src[swz_chunk(2 * k8 + 1, row)];
        }
        float ss = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) ss += __shfl_xor_sync(FULL, ss, o);  // 16 threads = one row
        if (k8 == 0) atomicMax(max_norm2_bits, __float_as_uint(ss));          // ss >= 0: uint order == float order
        __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
        o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
        const int64_t tile = row / TC_N;
        const int rit = int(row - tile * TC_N);
        *reinterpret_cast<uint4*>(b_img + tile * B_TILE_BYTES + image_offset(rit, k8)) = o;
    }
}

// ------------------------------------------------------------------------------------------------
// per search: normalise queries, write fp32 copy + bf16 operand images, reset per-query state
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tc_prep_kernel(const float* __restrict__ q_raw, int nq, int nq_pad, int qnorm,
                                                      float* __restrict__ qn, uint8_t* __restrict__ a_img,
                                                      float* __restrict__ thr, unsigned* __restrict__ cnt,
                                                      unsigned* __restrict__ flags, float* __restrict__ q_norm,
                                                      unsigned* __restrict__ n_flagged, unsigned first_round_rows) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (q == 0 && lane == 0) *n_flagged = 0u;
    if (q >= nq_pad) return;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    float nrm = 0.f;
    if (q < nq) {
        v = reinterpret_cast<const float4*>(q_raw + size_t(q) * DIM)[lane];
        float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(FULL, ss, o);
        nrm = sqrtf(ss);
        if (qnorm != FCS_QNORM_NONE) {
            const float d = fmaxf(nrm, (qnorm == FCS_QNORM_COSINE) ? 1e-8f : 1e-12f);
            v.x = v.x / d; v.y = v.y / d; v.z = v.z / d; v.w = v.w / d;
            nrm = nrm / d;
        }
        reinterpret_cast<float4*>(qn + size_t(q) * DIM)[lane] = v;
        if (lane == 0) {
            thr[q] = -INFINITY;
            cnt[q] = first_round_rows;  // the first round indexes its candidates by row
            flags[q] = 0u;
            q_norm[q] = nrm;
        }
    }
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&p0);
    o.y = *reinterpret_cast<uint32_t*>(&p1);
    const int tile = q / TC_M, m = q % TC_M;
    *reinterpret_cast<uint2*>(a_img + size_t(tile) * A_TILE_BYTES + image_offset(m, lane >> 1) + (lane & 1) * 8) = o;
}

// ------------------------------------------------------------------------------------------------
// K3
// ------------------------------------------------------------------------------------------------
struct TcGemmParams {
    const uint8_t* a_img;  // [n_qgroups*4][32 KB]
    const uint8_t* b_img;  // [n_tiles][16 KB]
    int64_t n_rows;
    int nq;
    int n_qgroups;
    int64_t tile0, tile1;  // DB tiles of this round
    const float* thr;      // [nq] approximate-score threshold (strict >)
    unsigned* cnt;         // [nq] append counters
    uint64_t* cand;        // [nq][TC_CAP] approximate keys (score, LOCAL row); 0 = unused reserved slot
    int first_round;       // thresholds are all -inf and tile0 == 0: slot = row, no atomics
    int trace_on;          // debug builds (FCS_TC_TRACE): record cycle stamps in this launch
    int res_block;         // slots reserved per atomic: small when many CTAs share a query group (unused slots are waste)
};

#ifdef FCS_TC_TRACE
// debug build only: cycle stamps of CTA 0 (MMA thread: row 0/1; epilogue warp of tile 0, quadrant 2: rows 2..4)
__device__ long long g_trace[5][256];
#define TRACE(row, idx, cond) do { if (p.trace_on && (cond) && (idx) < 256) g_trace[row][idx] = clock64(); } while (0)
#else
#define TRACE(row, idx, cond) do { } while (0)
#endif

// Per-thread slot reservation: one atomic buys p.res_block slots of the query's buffer.
struct SlotRes {
    unsigned base = 0, used = ~0u;  // ~0u: nothing reserved yet
};
__device__ __forceinline__ void tc_append(const TcGemmParams& p, SlotRes& res, int q, float v, int64_t row) {
    if (row >= p.n_rows) return;  // zero padding rows of the last DB tile
    unsigned slot;
    if (p.first_round) {
        slot = unsigned(row);
    } else {
        if (res.used >= unsigned(p.res_block)) {
            res.base = atomicAdd(p.cnt + q, unsigned(p.res_block));
            res.used = 0;
        }
        slot = res.base + res.used++;
    }
    if (slot < unsigned(TC_CAP)) p.cand[size_t(q) * TC_CAP + slot] = make_key(v, uint32_t(row));
}
// unused slots of the last reservation must read as empty
__device__ __forceinline__ void tc_close_reservation(const TcGemmParams& p, SlotRes& res, int q) {
    if (res.used == ~0u) return;  // this thread never appended
    for (; res.used < unsigned(p.res_block); ++res.used) {
        const unsigned slot = res.base + res.used;
        if (slot < unsigned(TC_CAP)) p.cand[size_t(q) * TC_CAP + slot] = 0ull;
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

    // contiguous share of the linearised (query group, DB tile) steps of this round
    const int64_t n_tiles = p.tile1 - p.tile0;
    const int64_t steps = int64_t(p.n_qgroups) * n_tiles;
    const int64_t s_begin = steps * blockIdx.x / gridDim.x;
    const int64_t s_end = steps * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (lane == 0) {
            const uint64_t pol_stream = policy_evict_first();
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
                        bulk_prefetch_l2(p.b_img + size_t(p.tile0 + ti + j + TC_PREFETCH) * B_TILE_BYTES, B_TILE_BYTES);
                    mbar_wait(&empty_b[st], ph ^ 1u);
                    mbar_arrive_expect_tx(&full_b[st], B_TILE_BYTES);
                    bulk_g2s(sB + st * B_TILE_BYTES, p.b_img + size_t(p.tile0 + ti + j) * B_TILE_BYTES, B_TILE_BYTES, &full_b[st], pol_stream);
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
                    tc_commit(&empty_b[st]);  // B stage free once both issuers' MMAs have read it
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
        for (int64_t s = s_begin; s < s_end;) {
            const int64_t qg = s / n_tiles, ti = s - qg * n_tiles;
            const int64_t seg_len = (s_end - s < n_tiles - ti) ? (s_end - s) : (n_tiles - ti);
            const int q = int(qg) * TC_QGROUP + t * TC_M + quad * 32 + lane;
            const float thr = (q < p.nq) ? p.thr[q] : INFINITY;
            SlotRes res;
            for (int64_t j = 0; j < seg_len; ++j, ++it) {
                const uint32_t tph = it & 1u;
                mbar_wait(&tmem_full[t], tph);
                tc_fence_after();
                TRACE(2, it, blockIdx.x == 0 && warp == 5 && lane == 0);
                const int64_t row_base = (p.tile0 + ti + j) * TC_N;
                int pend_n = 0;
                uint32_t pend_v0 = 0, pend_v1 = 0;
                int64_t pend_r0 = 0, pend_r1 = 0;
                // Fast path: 32 scores per tcgen05.ld, reduced with 3-input max, one compare against the
                // threshold.  Slow path (some lane of the warp has a hit in an 8-column group): the group is
                // re-read from TMEM into 8 fixed registers and handled by ONE compact, warp-uniform loop.  Keep
                // it small: the first version unrolled an append site per column (~140 KB of SASS) and every hit
                // ran through cold code -- instruction-cache misses cost 3-5 k cycles per append.
#pragma unroll 1
                for (int part = 0; part < TC_N / 32; ++part) {
                    const uint32_t taddr = tmem_base + lane_base + uint32_t(t * TC_N + part * 32);
                    float m[4];
                    {
                        uint32_t r[32];
                        tc_ld32(taddr, r);
                        tc_wait_ld();
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float* f = reinterpret_cast<const float*>(&r[g * 8]);
                            m[g] = fmaxf(max3(f[0], f[1], f[2]), max3(max3(f[3], f[4], f[5]), f[6], f[7]));
                        }
                    }
                    const float mx = fmaxf(max3(m[0], m[1], m[2]), m[3]);
                    if (__any_sync(FULL, p.first_round || mx > thr)) {  // rare once the threshold is warm
                        unsigned wgm = 0;  // groups in which some lane has a hit (warp-uniform)
#pragma unroll
                        for (int g = 0; g < 4; ++g) wgm |= __any_sync(FULL, p.first_round || m[g] > thr) ? (1u << g) : 0u;
                        while (wgm) {
                            const int g = __ffs(wgm) - 1;
                            wgm &= wgm - 1;
                            uint32_t v8[8];
                            __syncwarp();  // lanes left the divergent append loop below at different times
                            tc_ld8(taddr + uint32_t(g * 8), v8);
                            tc_wait_ld();
                            unsigned hits = 0;
#pragma unroll
                            for (int c = 0; c < 8; ++c) hits |= (p.first_round || __uint_as_float(v8[c]) > thr) ? (1u << c) : 0u;
                            while (hits) {
                                const int c = __ffs(hits) - 1;
                                hits &= hits - 1;
                                uint32_t bits = v8[0];
#pragma unroll
                                for (int cc = 1; cc < 8; ++cc) bits = (c == cc) ? v8[cc] : bits;
                                const int64_t hit_row = row_base + part * 32 + g * 8 + c;
                                // park up to two hits in registers: they are appended after the accumulator has been
                                // handed back to the MMA issuer (the append's atomics/stores are off the critical path)
                                if (pend_n == 0) {
                                    pend_v0 = bits;
                                    pend_r0 = hit_row;
                                    pend_n = 1;
                                } else if (pend_n == 1) {
                                    pend_v1 = bits;
                                    pend_r1 = hit_row;
                                    pend_n = 2;
                                } else {
                                    tc_append(p, res, q, __uint_as_float(bits), hit_row);
                                }
                            }
                        }
                    }
                }
                TRACE(3, it, blockIdx.x == 0 && warp == 5 && lane == 0);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[t]);
                if (pend_n > 0) tc_append(p, res, q, __uint_as_float(pend_v0), pend_r0);
                if (pend_n > 1) tc_append(p, res, q, __uint_as_float(pend_v1), pend_r1);
                TRACE(4, it, blockIdx.x == 0 && warp == 5 && lane == 0);
            }
            if (!p.first_round) tc_close_reservation(p, res, q);
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
// between rounds: keep each query's k' best candidates (sorted), publish the k'-th score as threshold
// ------------------------------------------------------------------------------------------------
// Fast path of the selection: one WARP per query, candidates in registers (<= 32 per lane), the k'-th largest
// score found by bisection on the 32-bit order-preserving score word (32 vote rounds), survivors compacted to the
// front of the buffer.  No sort: the rounds only need the k' best as a set and the k'-th score as threshold.
constexpr int SEL_WARP_MAX = 1024;
__global__ void __launch_bounds__(128) tc_select_warp_kernel(uint64_t* __restrict__ cand, unsigned* __restrict__ cnt,
                                                             float* __restrict__ thr, unsigned* __restrict__ flags, int nq,
                                                             int kprime) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (q >= nq) return;
    const unsigned raw = cnt[q];
    const unsigned fl = flags[q];
    if ((fl & 4u) && raw == (fl >> 8)) return;  // nothing appended since the last truncated state
    if (raw > unsigned(SEL_WARP_MAX)) return;    // tc_select_kernel's share
    uint64_t* base = cand + size_t(q) * TC_CAP;
    const int n = int(raw);
    uint64_t key[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int i = j * 32 + lane;
        key[j] = (i < n) ? base[i] : 0ull;
    }
    int nv = 0;  // real candidates (unused reserved slots hold key 0)
#pragma unroll
    for (int j = 0; j < 32; ++j) nv += (key[j] != 0ull) ? 1 : 0;
    nv = __reduce_add_sync(FULL, nv);
    const bool select = nv >= kprime;  // otherwise every real candidate is kept and the threshold stays
    uint32_t T = 0;                     // k'-th largest score word
    if (select) {
#pragma unroll 1
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t c = T | (1u << bit);
            int ge = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) ge += (uint32_t(key[j] >> 32) >= c) ? 1 : 0;  // empty slots have score word 0 < c
            if (__reduce_add_sync(FULL, ge) >= kprime) T = c;
        }
    }
    // survivors: score word > T always; score word == T (ties at the threshold) until k' are kept
    int above = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) above += (key[j] != 0ull && uint32_t(key[j] >> 32) > T) ? 1 : 0;
    above = __reduce_add_sync(FULL, above);
    int ties_left = select ? (kprime - above) : 0;
    int out = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const uint32_t hi = uint32_t(key[j] >> 32);
        const bool real = key[j] != 0ull;
        const bool is_tie = select && real && hi == T;
        const unsigned tie_mask = __ballot_sync(FULL, is_tie);
        const bool keep = real && (!select || hi > T || (is_tie && __popc(tie_mask & ((1u << lane) - 1u)) < ties_left));
        const unsigned keep_mask = __ballot_sync(FULL, keep);
        if (keep) base[out + __popc(keep_mask & ((1u << lane) - 1u))] = key[j];
        out += __popc(keep_mask);
        const int ties_taken = __popc(tie_mask);
        ties_left -= (ties_taken < ties_left) ? ties_taken : ties_left;
    }
    if (lane == 0) {
        cnt[q] = unsigned(out);
        if (select) thr[q] = unorder_f32(T);  // k'-th best approximate score over the rows seen so far
        flags[q] = (flags[q] & 0xFFu) | 4u | (unsigned(out) << 8);
    }
}

__global__ void __launch_bounds__(256) tc_select_kernel(uint64_t* __restrict__ cand, unsigned* __restrict__ cnt,
                                                        float* __restrict__ thr, unsigned* __restrict__ flags,
                                                        unsigned* __restrict__ n_flagged, int kprime) {
    extern __shared__ uint64_t s_keys[];
    __shared__ int s_valid;
    const int q = blockIdx.x, tid = threadIdx.x;
    const unsigned raw = cnt[q];
    const unsigned fl = flags[q];
    if ((fl & 4u) && raw == (fl >> 8)) return;  // nothing appended since the last truncated state
    if (raw <= unsigned(SEL_WARP_MAX)) return;   // tc_select_warp_kernel's share
    if (raw > unsigned(TC_CAP)) {  // lost candidates: the query falls back to the exact scan
        if (tid == 0 && (atomicOr(flags + q, 1u) & 3u) == 0u) atomicAdd(n_flagged, 1u);
    }
    const int n = int(raw < unsigned(TC_CAP) ? raw : unsigned(TC_CAP));
    int P = 2, logP = 1;
    while (P < n) { P <<= 1; ++logP; }
    uint64_t* base = cand + size_t(q) * TC_CAP;
    for (int i = tid; i < P; i += 256) s_keys[i] = (i < n) ? base[i] : 0ull;
    if (tid == 0) s_valid = 0;
    __syncthreads();
    // bitonic sort, descending; all strides are powers of two
    for (int ls = 1; ls <= logP; ++ls) {
        const int size = 1 << ls;
        for (int lt = ls - 1; lt >= 0; --lt) {
            const int stride = 1 << lt;
            for (int i = tid; i < (P >> 1); i += 256) {
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
    // unused reserved slots hold key 0 and sorted to the end: count the real candidates
    for (int i = tid; i < P; i += 256)
        if (s_keys[i] != 0ull && (i + 1 == P || s_keys[i + 1] == 0ull)) s_valid = i + 1;
    __syncthreads();
    const int nv = s_valid;
    const int keep = nv < kprime ? nv : kprime;
    for (int i = tid; i < keep; i += 256) base[i] = s_keys[i];
    if (tid == 0) {
        cnt[q] = unsigned(keep);
        if (nv >= kprime) thr[q] = key_score(s_keys[kprime - 1]);
        flags[q] = (flags[q] & 0xFFu) | 4u | (unsigned(keep) << 8);
    }
}

// ------------------------------------------------------------------------------------------------
// K4: exact fp32 rescore of the k' candidates, top-k, certificate
// ------------------------------------------------------------------------------------------------
struct TcRescoreParams {
    const float* rows;      // fp32 row-swizzled shard
    const float* qn;        // [nq][128] normalised queries
    const uint64_t* cand;   // [nq][TC_CAP] sorted approximate keys
    const unsigned* cnt;    // [nq] <= kprime
    const float* q_norm;    // [nq] |q| as used
    const float* thr;       // [nq] final approximate-score thresholds
    unsigned* flags;
    unsigned* n_flagged;
    int nq, k, kprime;
    uint32_t id_base;
    float eps_rel;          // 0.004 * max |row|
    uint64_t* out_keys;
    float* out_scores;
    int64_t* out_ids;
};

__global__ void __launch_bounds__(128) tc_rescore_kernel(const TcRescoreParams p) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (q >= p.nq) return;
    const int n = int(p.cnt[q]);
    const uint64_t* cand = p.cand + size_t(q) * TC_CAP;
    const float4* q4 = reinterpret_cast<const float4*>(p.qn + size_t(q) * DIM);
    WarpTopK<4> tk;
    tk.init();
    constexpr int U = 8;  // candidate rows in flight per warp
    uint64_t batch = 0;   // lane i holds the exact key of candidate (c0 + i) of the current 32-batch
    for (int c0 = 0; c0 < n; c0 += 32) {
        batch = 0;
        for (int u0 = 0; u0 < 32 && c0 + u0 < n; u0 += U) {
            float part[U];
            int64_t rowid[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int c = c0 + u0 + u;
                part[u] = 0.f;
                rowid[u] = -1;
                if (c < n) {
                    const int64_t row = key_id(cand[c]);
                    rowid[u] = row;
                    const float4 v = reinterpret_cast<const float4*>(p.rows + row * DIM)[lane];  // physical chunk `lane`
                    const float4 w = q4[swz_chunk(lane, row)];                                    // = logical chunk
                    part[u] = fmaf(v.w, w.w, fmaf(v.z, w.z, fmaf(v.y, w.y, v.x * w.x)));
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
    // certificate: every row that is NOT a candidate has approximate score <= t' (the k'-th approximate
    // score) and therefore exact score <= t' + eps; the result is exact if the k-th exact score beats that.
    bool ok = true;
    if (n >= p.kprime) {
        const float tprime = p.thr[q];  // k'-th best approximate score over all rows (last selection)
        const float sk = key_score(tk.thr);  // k-th best exact (thr == 0 -> -inf if fewer than k candidates)
        const float eps = p.eps_rel * p.q_norm[q] + 2e-5f;
        ok = sk > tprime + eps;
    }
    if (!ok && lane == 0) {
        if ((atomicOr(p.flags + q, 2u) & 3u) == 0u) atomicAdd(p.n_flagged, 1u);
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

}  // namespace

struct TcState {
    int device = 0, sm_count = 0;
    const float* rows = nullptr;
    int64_t n_rows = 0, n_tiles = 0;
    uint32_t id_base = 0;
    uint8_t* b_img = nullptr;
    float max_norm = 1.f;
    // per-search workspace, grown on demand
    int nq_cap = 0;
    float* qn = nullptr;
    uint8_t* a_img = nullptr;
    float* thr = nullptr;
    unsigned* cnt = nullptr;
    uint64_t* cand = nullptr;
    unsigned* flags = nullptr;
    float* q_norm = nullptr;
    unsigned* n_flagged = nullptr;
    unsigned* h_n_flagged = nullptr;  // pinned
    unsigned* h_flags = nullptr;      // pinned, nq_cap
    double growth = 2.0;  // rows seen grow x3 per round; measured best on B200 (profiles/r01_ncu_summary.md)
    bool verbose = false;
    // timing of the dominant kernel: one event pair per K3 launch of the last search
    static constexpr int MAX_ROUNDS = 48;
    cudaEvent_t ev[2 * MAX_ROUNDS] = {};
    float last_k3_ms = 0.f;
    int last_rounds = 0;
};

const char* tc_last_error() { return g_tc_error.c_str(); }
int tc_min_batch() { return 32; }
int tc_max_k() { return 128; }
float tc_last_kernel_ms(const TcState* s) { return s ? s->last_k3_ms : 0.f; }
int tc_last_rounds(const TcState* s) { return s ? s->last_rounds : 0; }
uint64_t tc_image_bytes(const TcState* s) { return s ? uint64_t(s->n_tiles) * B_TILE_BYTES : 0; }

static void tc_free_workspace(TcState* s) {
    cudaFree(s->qn); cudaFree(s->a_img); cudaFree(s->thr); cudaFree(s->cnt); cudaFree(s->cand);
    cudaFree(s->flags); cudaFree(s->q_norm);
    if (s->h_flags) cudaFreeHost(s->h_flags);
    s->qn = nullptr; s->a_img = nullptr; s->thr = nullptr; s->cnt = nullptr; s->cand = nullptr;
    s->flags = nullptr; s->q_norm = nullptr; s->h_flags = nullptr;
    s->nq_cap = 0;
}

void tc_destroy(TcState* s) {
    if (!s) return;
    tc_free_workspace(s);
    for (cudaEvent_t e : s->ev)
        if (e) cudaEventDestroy(e);
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
    if (const char* g = getenv("FCS_TC_GROWTH")) {
        const double v = atof(g);
        if (v >= 1.25 && v <= 16.0) s->growth = v;
    }
    auto run = [&]() -> int {
        TC_CUDA(cudaFuncSetAttribute(tc_gemm_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        TC_CUDA(cudaFuncSetAttribute(tc_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_CAP * 8));
        // the small kernels between two GEMM launches ask for the same shared-memory carve-out as the GEMM kernel,
        // so the SMs are not reconfigured (drained) twice per round
        TC_CUDA(cudaFuncSetAttribute(tc_select_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TC_CUDA(cudaFuncSetAttribute(tc_select_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TC_CUDA(cudaFuncSetAttribute(tc_prep_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TC_CUDA(cudaFuncSetAttribute(tc_rescore_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        for (cudaEvent_t& e : s->ev) TC_CUDA(cudaEventCreate(&e));
        TC_CUDA(cudaMalloc(&s->b_img, size_t(s->n_tiles) * B_TILE_BYTES));
        TC_CUDA(cudaMalloc(&s->n_flagged, 2 * sizeof(unsigned)));
        TC_CUDA(cudaMallocHost(&s->h_n_flagged, 2 * sizeof(unsigned)));
        TC_CUDA(cudaMemsetAsync(s->n_flagged, 0, 2 * sizeof(unsigned), stream));
        const int64_t total = s->n_tiles * TC_N * 16;
        int64_t g = (total + 255) / 256;
        if (g > sm_count * 16) g = sm_count * 16;
        tc_build_bimg_kernel<<<int(g), 256, 0, stream>>>(rows, n_rows, s->n_tiles, s->b_img, s->n_flagged + 1);
        TC_CUDA(cudaGetLastError());
        TC_CUDA(cudaMemcpyAsync(s->h_n_flagged, s->n_flagged, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
        TC_CUDA(cudaStreamSynchronize(stream));
        float n2;
        memcpy(&n2, &s->h_n_flagged[1], 4);
        s->max_norm = sqrtf(n2);
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

static int tc_ensure_workspace(TcState* s, int nq) {
    if (nq <= s->nq_cap) return FCS_OK;
    tc_free_workspace(s);
    const int nq_pad = ((nq + TC_QGROUP - 1) / TC_QGROUP) * TC_QGROUP;
    TC_CUDA(cudaMalloc(&s->qn, size_t(nq_pad) * DIM * 4));
    TC_CUDA(cudaMalloc(&s->a_img, size_t(nq_pad / TC_M) * A_TILE_BYTES));
    TC_CUDA(cudaMalloc(&s->thr, size_t(nq_pad) * 4));
    TC_CUDA(cudaMalloc(&s->cnt, size_t(nq_pad) * 4));
    TC_CUDA(cudaMalloc(&s->flags, size_t(nq_pad) * 4));
    TC_CUDA(cudaMalloc(&s->q_norm, size_t(nq_pad) * 4));
    TC_CUDA(cudaMalloc(&s->cand, size_t(nq_pad) * TC_CAP * 8));
    TC_CUDA(cudaMallocHost(&s->h_flags, size_t(nq_pad) * 4));
    s->nq_cap = nq_pad;
    return FCS_OK;
}

int tc_default_kprime(int k) {
    // margin for the exactness certificate: the bf16 rounding bound eps covers a few dozen ranks at TED scale
    int kp = k + (k * 6 / 10 > 32 ? k * 6 / 10 : 32);
    kp = (kp + 31) / 32 * 32;
    return kp > TC_MAX_KPRIME ? TC_MAX_KPRIME : kp;
}

int tc_search(TcState* s, const float* q_dev, int nq, int k, int kprime, int qnorm, float* out_scores, int64_t* out_ids,
              uint64_t* out_keys, cudaStream_t stream, int* launches, const unsigned** flagged_host, int* n_flagged_out) {
    *n_flagged_out = 0;
    *flagged_host = nullptr;
    if (k > tc_max_k()) {
        g_tc_error = "k too large for the tensor-core path";
        return FCS_ERR_UNSUPPORTED;
    }
    int kp = kprime > 0 ? kprime : tc_default_kprime(k);
    if (kp < k) kp = k;
    if (kp > TC_MAX_KPRIME) kp = TC_MAX_KPRIME;
    int rc = tc_ensure_workspace(s, nq);
    if (rc != FCS_OK) return rc;
    const int n_qgroups = (nq + TC_QGROUP - 1) / TC_QGROUP;
    const int nq_pad = n_qgroups * TC_QGROUP;

    // the first round (threshold -inf: every row is a candidate) covers at most CAP/2 rows and indexes by row
    const int64_t first_tiles = SEL_WARP_MAX / TC_N;
    const int64_t first_rows = (first_tiles * TC_N < s->n_rows) ? first_tiles * TC_N : s->n_rows;
    tc_prep_kernel<<<(nq_pad + 3) / 4, 128, 0, stream>>>(q_dev, nq, nq_pad, qnorm, s->qn, s->a_img, s->thr, s->cnt, s->flags,
                                                         s->q_norm, s->n_flagged, unsigned(first_rows));
    TC_CUDA(cudaGetLastError());
    ++*launches;

    TcGemmParams gp = {};
    gp.a_img = s->a_img;
    gp.b_img = s->b_img;
    gp.n_rows = s->n_rows;
    gp.nq = nq;
    gp.n_qgroups = n_qgroups;
    gp.thr = s->thr;
    gp.cnt = s->cnt;
    gp.cand = s->cand;
    // rounds: a round over R rows appends about k' * R / seen keys per query
    int64_t seen = 0;
    int rounds = 0;
    int64_t round_tiles = first_tiles;
    while (seen < s->n_tiles) {
        const int64_t t1 = (seen + round_tiles < s->n_tiles) ? (seen + round_tiles) : s->n_tiles;
        gp.tile0 = seen;
        gp.tile1 = t1;
        gp.first_round = rounds == 0 ? 1 : 0;
        gp.trace_on = (getenv("FCS_TC_TRACE_ROUND") ? atoi(getenv("FCS_TC_TRACE_ROUND")) : 7) == rounds;
        const int64_t steps = int64_t(n_qgroups) * (t1 - seen);
        const int grid = int(steps < s->sm_count ? steps : s->sm_count);
        {   // every thread that appends at all rounds its reservation up to res_block slots: keep the waste of the
            // ~grid/n_qgroups segments that share a query below ~512 slots so the buffers stay in the warp-select range
            const int segs = (grid + n_qgroups - 1) / n_qgroups + 1;
            int rb = 512 / segs;
            gp.res_block = rb < 2 ? 2 : (rb > TC_RES_MAX ? TC_RES_MAX : rb);
        }
        const bool timed = rounds < TcState::MAX_ROUNDS;
        if (timed) TC_CUDA(cudaEventRecord(s->ev[2 * rounds], stream));
        tc_gemm_filter_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(gp);
        TC_CUDA(cudaGetLastError());
        if (timed) TC_CUDA(cudaEventRecord(s->ev[2 * rounds + 1], stream));
        ++rounds;
        tc_select_warp_kernel<<<(nq + 3) / 4, 128, 0, stream>>>(s->cand, s->cnt, s->thr, s->flags, nq, kp);
        TC_CUDA(cudaGetLastError());
        tc_select_kernel<<<nq, 256, TC_CAP * 8, stream>>>(s->cand, s->cnt, s->thr, s->flags, s->n_flagged, kp);
        TC_CUDA(cudaGetLastError());
        *launches += 3;
        seen = t1;
        double g = double(TC_CAP - kp) / (2.0 * kp);
        if (g > s->growth) g = s->growth;
        if (g < 0.25) g = 0.25;
        round_tiles = int64_t(double(seen) * g);
        if (round_tiles < 1) round_tiles = 1;
    }

    TcRescoreParams rp = {};
    rp.rows = s->rows;
    rp.qn = s->qn;
    rp.cand = s->cand;
    rp.cnt = s->cnt;
    rp.q_norm = s->q_norm;
    rp.thr = s->thr;
    rp.flags = s->flags;
    rp.n_flagged = s->n_flagged;
    rp.nq = nq;
    rp.k = k;
    rp.kprime = kp;
    rp.id_base = s->id_base;
    rp.eps_rel = 0.004f * s->max_norm;
    rp.out_keys = out_keys;
    rp.out_scores = out_scores;
    rp.out_ids = out_ids;
    tc_rescore_kernel<<<(nq + 3) / 4, 128, 0, stream>>>(rp);
    TC_CUDA(cudaGetLastError());
    ++*launches;

    // which queries need the exact fallback?  (host decision: one small synchronous read-back)
    TC_CUDA(cudaMemcpyAsync(s->h_n_flagged, s->n_flagged, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    TC_CUDA(cudaStreamSynchronize(stream));
#ifdef FCS_TC_TRACE
    {
        static long long h[5][256];
        cudaMemcpyFromSymbol(h, g_trace, sizeof h);
        fprintf(stderr, "[fcs_tc trace, last round, CTA 0, cycles relative to first stamp]\n it  mma_ready mma_issued | epi_full epi_done epi_arrived\n");
        for (int i = 100; i < 116; ++i)
            fprintf(stderr, "%3d  %9lld %9lld | %9lld %9lld %9lld\n", i, h[0][i] - h[0][100], h[1][i] - h[0][100], h[2][i] - h[0][100],
                    h[3][i] - h[0][100], h[4][i] - h[0][100]);
    }
#endif
    s->last_k3_ms = 0.f;
    s->last_rounds = rounds;
    for (int r = 0; r < rounds && r < TcState::MAX_ROUNDS; ++r) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s->ev[2 * r], s->ev[2 * r + 1]) == cudaSuccess) s->last_k3_ms += ms;
        if (s->verbose) fprintf(stderr, "[fcs_tc] round %d: gemm+filter %.3f ms\n", r, ms);
    }
    if (s->h_n_flagged[0] > 0) {
        TC_CUDA(cudaMemcpyAsync(s->h_flags, s->flags, size_t(nq) * 4, cudaMemcpyDeviceToHost, stream));
        TC_CUDA(cudaStreamSynchronize(stream));
        *n_flagged_out = int(s->h_n_flagged[0]);
        *flagged_host = s->h_flags;
    }
    return FCS_OK;
}

// Test hook: approximate (bf16 tensor-core) scores of every (query, row) pair, for shards of at most
// TC_CAP rows.  One K3 launch over all tiles with the threshold at -inf appends every row.
int tc_debug_approx(TcState* s, const float* q_dev, int nq, int qnorm, float* out_host, cudaStream_t stream) {
    if (s->n_rows > TC_CAP) {
        g_tc_error = "tc_debug_approx: shard larger than the candidate buffer";
        return FCS_ERR_UNSUPPORTED;
    }
    int rc = tc_ensure_workspace(s, nq);
    if (rc != FCS_OK) return rc;
    const int n_qgroups = (nq + TC_QGROUP - 1) / TC_QGROUP;
    const int nq_pad = n_qgroups * TC_QGROUP;
    tc_prep_kernel<<<(nq_pad + 3) / 4, 128, 0, stream>>>(q_dev, nq, nq_pad, qnorm, s->qn, s->a_img, s->thr, s->cnt, s->flags,
                                                         s->q_norm, s->n_flagged, unsigned(s->n_rows));
    TC_CUDA(cudaGetLastError());
    TcGemmParams gp = {};
    gp.first_round = 1;
    gp.res_block = TC_RES_MAX;
    gp.a_img = s->a_img; gp.b_img = s->b_img; gp.n_rows = s->n_rows; gp.nq = nq; gp.n_qgroups = n_qgroups;
    gp.thr = s->thr; gp.cnt = s->cnt; gp.cand = s->cand; gp.tile0 = 0; gp.tile1 = s->n_tiles;
    const int64_t steps = int64_t(n_qgroups) * s->n_tiles;
    tc_gemm_filter_kernel<<<int(steps < s->sm_count ? steps : s->sm_count), TC_THREADS, TC_SMEM, stream>>>(gp);
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaStreamSynchronize(stream));
    std::string keys(size_t(nq) * TC_CAP * 8, '\0');
    std::string cnts(size_t(nq) * 4, '\0');
    TC_CUDA(cudaMemcpy(&keys[0], s->cand, keys.size(), cudaMemcpyDeviceToHost));
    TC_CUDA(cudaMemcpy(&cnts[0], s->cnt, cnts.size(), cudaMemcpyDeviceToHost));
    const uint64_t* kk = reinterpret_cast<const uint64_t*>(keys.data());
    const unsigned* cc = reinterpret_cast<const unsigned*>(cnts.data());
    for (size_t i = 0; i < size_t(nq) * s->n_rows; ++i) out_host[i] = NAN;
    for (int q = 0; q < nq; ++q) {
        const unsigned n = cc[q] < unsigned(TC_CAP) ? cc[q] : unsigned(TC_CAP);
        for (unsigned i = 0; i < n; ++i) {
            const uint64_t key = kk[size_t(q) * TC_CAP + i];
            const int64_t row = key_id(key);
            if (row >= 0 && row < s->n_rows) out_host[size_t(q) * s->n_rows + row] = key_score(key);
        }
    }
    return FCS_OK;
}

}  // namespace fcs
