// fcs_gemv.cu -- K2: exact fp32 cosine / inner-product scan of a row shard with a fused top-k.
//
// Replaces the arithmetic of search_query_against_db (reference dbsearch.py:75-81: cosine * coverage
// mask -> topk) and, for small batches, of knn_exact_faiss (dbsearch.py:213-248).  HBM-bound: the
// shard is read exactly once per launch (512 B per row, + 2 B per row when the coverage mask is on)
// for up to 8 queries.
//
// Shape of the kernel (one persistent CTA per SM, 12 warps):
//   * warp 11          -- producer: one lane streams 16 KB chunks (32 rows) into an 11-stage
//     shared-memory ring with 1-D bulk asynchronous copies (cp.async.bulk -> SASS UBLKCP, the TMA
//     engine), one mbarrier pair (full/empty) per stage; 176 KB in flight per SM.  It starts
//     streaming before the consumers have finished preparing the queries.
//   * warps 0..10      -- consumers: warp w owns stage w and scores one ROW PER LANE.  Rows are
//     stored chunk-swizzled (fcs_common.cuh) so that the 32 lanes read their 32 different rows with
//     conflict-free LDS.128; the query chunks are shared-memory broadcasts; packed FFMA2 (64 per row
//     and query); no cross-lane reduction.  The score is masked, packed into a sortable 64-bit key and
//     offered to a warp-private, register-resident sorted top-k list (WarpTopK).
//   * epilogue: the 11 warp lists are tree-merged (bitonic merges through shared memory) into one
//     CTA list -> global scratch; the last CTA to finish (atomic ticket) merges the <=148 CTA lists
//     the same way and writes scores / ids / keys.
// Chunks are dealt round-robin to CTAs (chunk c -> CTA c % grid), so every SM streams the same
// number of bytes +-16 KB.
#include "fcs_common.cuh"
#include "fcs_internal.h"

namespace fcs {

namespace {

constexpr int STAGE_ROWS = GEMV_STAGE_ROWS;
constexpr int STAGE_BYTES = STAGE_ROWS * ROW_BYTES;  // 16 KB
constexpr int NWARPS = GEMV_WARPS;                   // consumer warps
constexpr int NTHREADS = (NWARPS + 1) * 32;
constexpr int RING_BYTES = NWARPS * STAGE_BYTES;  // 176 KB
constexpr int SMEM_BYTES = RING_BYTES + 2 * NWARPS * 8 + GEMV_MAX_NQ * DIM * 4 + 16;

// two fp32 FMAs per instruction (SASS FFMA2): halves the issue slots of the dot products
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

// Tree-merge `n` sorted lists held in shared memory (list i at lists + i*stride, k keys each); the
// warps of the CTA cooperate; on return list 0 holds the k best.  All NTHREADS threads must call.
template <int KPL>
__device__ __forceinline__ void tree_merge_smem(uint64_t* lists, size_t stride, int n, int k, int warp, int lane) {
    for (int half = 8; half >= 1; half >>= 1) {  // n <= 16
        if (warp < half && warp + half < n) {
            WarpTopK<KPL> a, b;
            a.template load<false>(lists + size_t(warp) * stride, lane, k);
            b.template load<false>(lists + size_t(warp + half) * stride, lane, k);
            a.merge_sorted(b.key, lane, k);
            a.store(lists + size_t(warp) * stride, lane, k);
        }
        if (n > half) n = half;
        __syncthreads();
    }
}

template <int NQ, int KPL>
__global__ void __launch_bounds__(NTHREADS, 1) gemv_topk_kernel(const GemvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RING_BYTES);
    uint64_t* empty_bar = full_bar + NWARPS;
    float* qs = reinterpret_cast<float*>(empty_bar + NWARPS);  // [NQ][128] normalised queries
    __shared__ int s_is_last;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int k = p.k;
    const int64_t n_chunks = (p.n_rows + STAGE_ROWS - 1) / STAGE_ROWS;
    int nq = p.nq;
    if (p.nq_dev) {
        // indirect launch: the queue length was written by a predecessor on the stream.  Every thread waits for it
        // here (the producer may not prefetch before the CTA knows whether it has work: a CTA must not exit with
        // bulk copies in flight).
        pdl_wait();
        const unsigned total = *reinterpret_cast<const volatile unsigned*>(p.nq_dev);
        const int rem = total > unsigned(p.nq_off) ? int(total - unsigned(p.nq_off)) : 0;
        nq = rem < NQ ? rem : NQ;
        if (nq == 0) return;
    }

    WarpTopK<KPL> tk[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) tk[q].init();

    if (warp == NWARPS) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            for (int s = 0; s < NWARPS; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            mbar_fence_init();
        }
        __syncwarp();
        named_bar_arrive(1, NTHREADS);  // barriers are live; consumers wait for this
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            int64_t it = 0;
            for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
                const int s = int(it % NWARPS);
                const uint32_t j = uint32_t(it / NWARPS);
                mbar_wait(&empty_bar[s], (j & 1u) ^ 1u);
                const int64_t row0 = c * STAGE_ROWS;
                const int64_t rem = p.n_rows - row0;
                const uint32_t bytes = uint32_t(rem < STAGE_ROWS ? rem : STAGE_ROWS) * ROW_BYTES;
                mbar_arrive_expect_tx(&full_bar[s], bytes);
                bulk_g2s(ring + s * STAGE_BYTES, p.rows + row0 * DIM, bytes, &full_bar[s], pol);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ consumers
        // Everything a predecessor kernel on the stream may have written (the queries, the previous pass's keys,
        // and -- at the end -- scratch/ticket/outputs) is touched only after this point; the producer warp is
        // already prefetching database rows (immutable after finalize) into the ring meanwhile.
        pdl_wait();
        // query normalisation fused here (dbsearch.py:78 cosine eps 1e-8 / dbsearch.py:304 F.normalize eps 1e-12)
        if (warp < NQ) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (warp < nq) v = reinterpret_cast<const float4*>(p.q + size_t(warp) * DIM)[lane];
            if (p.qnorm != FCS_QNORM_NONE) {
                float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(FULL, ss, o);
                const float d = fmaxf(sqrtf(ss), (p.qnorm == FCS_QNORM_COSINE) ? 1e-8f : 1e-12f);
                v.x = v.x / d; v.y = v.y / d; v.z = v.z / d; v.w = v.w / d;
            }
            reinterpret_cast<float4*>(qs + warp * DIM)[lane] = v;
        }
        named_bar_sync(1, NTHREADS);
        const float4* qs4 = reinterpret_cast<const float4*>(qs);

        uint64_t ub[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            ub[q] = ~0ull;
            if (p.bounded && q < nq) ub[q] = p.out_keys[size_t(p.out_index ? p.out_index[q] : q) * p.out_stride + p.out_off - 1];
        }
        // this lane's row inside the stage, and its 8 swizzled chunk offsets (bytes)
        const uint8_t* my_row = ring + warp * STAGE_BYTES + lane * ROW_BYTES;
        uint32_t swz[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) swz[c] = uint32_t((c ^ (lane & 7)) * 16);

        uint32_t j = 0;
        for (int64_t it = warp;; it += NWARPS, ++j) {
            const int64_t c = blockIdx.x + it * gridDim.x;
            if (c >= n_chunks) break;
            const int64_t row = c * STAGE_ROWS + lane;
            const bool valid = row < p.n_rows;
            float lenf = 0.f;
            if (p.use_mask && valid) lenf = float(__ldg(p.lens + row));
            mbar_wait(&full_bar[warp], j & 1u);

            float2 acc0[NQ], acc1[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) acc0[q] = acc1[q] = make_float2(0.f, 0.f);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const float4 v = *reinterpret_cast<const float4*>(my_row + g * 128 + swz[cc]);
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const float4 w = qs4[q * 32 + g * 8 + cc];  // warp-uniform address: broadcast
                        acc0[q] = fma2(make_float2(v.x, v.y), make_float2(w.x, w.y), acc0[q]);
                        acc1[q] = fma2(make_float2(v.z, v.w), make_float2(w.z, w.w), acc1[q]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[warp]);  // stage goes back to the producer

#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                float score = (acc0[q].x + acc0[q].y) + (acc1[q].x + acc1[q].y);
                if (p.use_mask) {
                    // dbsearch.py:76: (qlen >= lengths * mincov).float(), an fp32 product
                    const float need = __fmul_rn(lenf, p.mincov);
                    score = score * ((p.qlen[q] >= need) ? 1.0f : 0.0f);
                }
                uint64_t cand = make_key(score, p.id_base + uint32_t(row));
                cand = (valid && cand < ub[q]) ? cand : 0ull;
                tk[q].offer(cand, lane, k);
            }
        }
    }

    // ---------------------------------------------------------------------- CTA-level merge
    if (warp == NWARPS) pdl_wait();  // the producer warp joins the writes below (consumers waited above)
    __syncthreads();  // every bulk copy has landed and been consumed: the ring can be reused
    pdl_launch_dependents();  // the next search on this stream may start prefetching while we merge
    uint64_t* lists = reinterpret_cast<uint64_t*>(ring);  // [NQ][NWARPS][k]
    if (warp < NWARPS) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) tk[q].store(lists + (size_t(q) * NWARPS + warp) * k, lane, k);
    }
    __syncthreads();
    for (int q = 0; q < NQ; ++q) tree_merge_smem<KPL>(lists + size_t(q) * NWARPS * k, size_t(k), NWARPS, k, warp, lane);
    // list q*NWARPS holds query q's CTA result -> scratch[q][cta][k]
    for (int i = tid; i < NQ * k; i += NTHREADS) {
        const int q = i / k, r = i - q * k;
        p.scratch[(size_t(q) * gridDim.x + blockIdx.x) * k + r] = lists[size_t(q) * NWARPS * k + r];
    }
    // ---------------------------------------------------------------------- last CTA merges the grid
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(p.ticket, 1u);
        s_is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();
    const int G = int(gridDim.x);
    for (int q = 0; q < NQ; ++q) {
        if (q >= nq) break;
        const uint64_t* src = p.scratch + size_t(q) * G * k;
        uint64_t* mine = lists + size_t(warp) * k;  // one smem list per warp (12 warps incl. the producer's)
        {
            // this warp's share of the CTA lists: fetched in chunks with all loads in flight (the lists sit in
            // L2; one dependent load per list would cost a full L2 round trip each), then merged
            WarpTopK<KPL> a;
            a.init();
            constexpr int CH = 16 / KPL;
            for (int b0 = warp; b0 < G; b0 += CH * (NWARPS + 1)) {
                WarpTopK<KPL> o[CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int b = b0 + c * (NWARPS + 1);
                    if (b < G) o[c].template load<true>(src + size_t(b) * k, lane, k);
                    else o[c].init();
                }
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    if (b0 + c * (NWARPS + 1) < G) a.merge_sorted(o[c].key, lane, k);
            }
            a.store(mine, lane, k);
        }
        __syncthreads();
        {   // 12 -> 1 (tree over warps 0..11; tree_merge_smem handles n <= 16)
            int n = NWARPS + 1;
            for (int half = 8; half >= 1; half >>= 1) {
                if (warp < half && warp + half < n) {
                    WarpTopK<KPL> a, b;
                    a.template load<false>(lists + size_t(warp) * k, lane, k);
                    b.template load<false>(lists + size_t(warp + half) * k, lane, k);
                    a.merge_sorted(b.key, lane, k);
                    a.store(lists + size_t(warp) * k, lane, k);
                }
                if (n > half) n = half;
                __syncthreads();
            }
        }
        const size_t base = size_t(p.out_index ? p.out_index[q] : q) * p.out_stride + p.out_off;
        for (int r = tid; r < k; r += NTHREADS) {
            const uint64_t key = lists[r];
            p.out_keys[base + r] = key;
            if (p.out_scores) p.out_scores[base + r] = key_score(key);
            if (p.out_ids) p.out_ids[base + r] = key_id(key);
        }
        __syncthreads();
    }
    if (tid == 0) *p.ticket = 0u;  // ready for the next launch on this stream
}

template <int NQ, int KPL>
cudaError_t launch_inst(const GemvParams& p, int grid, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gemv_topk_kernel<NQ, KPL>, p);
}

template <int NQ, int KPL>
cudaError_t configure_inst() {
    return cudaFuncSetAttribute(gemv_topk_kernel<NQ, KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
}

}  // namespace

size_t gemv_scratch_bytes(int max_grid) { return size_t(max_grid) * GEMV_MAX_NQ * GEMV_MAX_K * sizeof(uint64_t); }

#define FCS_FOR_ALL_INST(X) \
    X(1, 1) X(1, 2) X(1, 4) X(2, 1) X(2, 2) X(2, 4) X(4, 1) X(4, 2) X(4, 4) X(8, 1) X(8, 2) X(8, 4)

cudaError_t gemv_configure() {
    cudaError_t e;
#define FCS_CFG(NQ, KPL) \
    if ((e = configure_inst<NQ, KPL>()) != cudaSuccess) return e;
    FCS_FOR_ALL_INST(FCS_CFG)
#undef FCS_CFG
    return cudaSuccess;
}

cudaError_t gemv_launch(const GemvParams& p, int sm_count, cudaStream_t stream) {
    if (p.nq < 1 || p.nq > GEMV_MAX_NQ || p.k < 1 || p.k > GEMV_MAX_K || p.n_rows < 1) return cudaErrorInvalidValue;
    if (p.nq_dev && p.nq != GEMV_MAX_NQ) return cudaErrorInvalidValue;  // indirect launches use the 8-query instantiation
    const int64_t n_chunks = (p.n_rows + GEMV_STAGE_ROWS - 1) / GEMV_STAGE_ROWS;
    const int grid = int(n_chunks < sm_count ? n_chunks : sm_count);
    const int nqt = p.nq == 1 ? 1 : (p.nq == 2 ? 2 : (p.nq <= 4 ? 4 : 8));
    const int kpl = p.k <= 32 ? 1 : (p.k <= 64 ? 2 : 4);
#define FCS_GO(NQ, KPL) \
    if (nqt == NQ && kpl == KPL) return launch_inst<NQ, KPL>(p, grid, stream);
    FCS_FOR_ALL_INST(FCS_GO)
#undef FCS_GO
    return cudaErrorInvalidValue;
}

}  // namespace fcs
