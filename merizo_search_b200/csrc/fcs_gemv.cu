// fcs_gemv.cu -- K2: exact fp32 cosine / inner-product scan of a row shard with a fused top-k.
//
// Replaces the arithmetic of search_query_against_db (reference dbsearch.py:75-81: cosine * coverage
// mask -> topk) and, for small batches, of knn_exact_faiss (dbsearch.py:213-248).  HBM-bound: the
// shard is read exactly once per launch (512 B per row, + 2 B per row when the coverage mask is on)
// for up to 4 queries.
//
// Shape of the kernel (one persistent CTA per SM, 13 warps):
//   * warp 12, lane 0  -- producer: streams 16 KB chunks (32 rows) into a 12-stage shared-memory
//     ring with 1-D bulk asynchronous copies (cp.async.bulk -> SASS UBLKCP, the TMA engine),
//     one mbarrier pair (full/empty) per stage; 192 KB in flight per SM.
//   * warps 0..11      -- consumers: warp w owns stage w.  Lane l reads 16 B chunk l of each of the
//     32 rows (conflict-free LDS.128), forms the 4-element partial dot, releases the stage, then a
//     transposing butterfly (31 shuffles per 32 rows) leaves lane r with the full dot of row r.
//     The score is masked, packed into a sortable 64-bit key and offered to a warp-private,
//     register-resident sorted top-k list (WarpTopK): one compare + ballot per 32 rows once warm.
//   * epilogue: 12 warp lists -> one CTA list (shared memory) -> global scratch; the last CTA to
//     finish (atomic ticket) merges the <=148 CTA lists and writes scores / ids / keys.
// Chunks are dealt round-robin to CTAs (chunk c -> CTA c % grid), so every SM streams the same
// number of bytes +-16 KB.
#include "fcs_common.cuh"
#include "fcs_internal.h"

namespace fcs {

namespace {

constexpr int STAGE_ROWS = GEMV_STAGE_ROWS;
constexpr int STAGE_BYTES = STAGE_ROWS * ROW_BYTES;  // 16 KB
constexpr int NWARPS = GEMV_WARPS;
constexpr int NTHREADS = (NWARPS + 1) * 32;
constexpr int RING_BYTES = NWARPS * STAGE_BYTES;  // 192 KB
constexpr int SMEM_BYTES = RING_BYTES + 2 * NWARPS * 8 + GEMV_MAX_NQ * DIM * 4 + 16;

// Transposing butterfly: in: acc[r] = this lane's partial of row r (r < 32);
// out: acc[0] in lane l = sum over lanes of the partials of row l.
__device__ __forceinline__ void transpose_reduce32(float (&acc)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? acc[i] : acc[i + half];
            const float keep = up ? acc[i + half] : acc[i];
            acc[i] = keep + __shfl_xor_sync(FULL, send, half);
        }
    }
}

// merge `n_lists` sorted lists of length k (stride `stride` keys apart) into `m` (one warp)
template <int KPL, bool CG>
__device__ __forceinline__ void merge_lists(WarpTopK<KPL>& m, const uint64_t* lists, int first, int step, int n_lists,
                                            size_t stride, int k, int lane) {
    for (int s = first; s < n_lists; s += step) {
        const uint64_t* src = lists + size_t(s) * stride;
        for (int i = 0; i < k; i += 32) {
            uint64_t cand = 0;
            if (i + lane < k) cand = CG ? __ldcg(reinterpret_cast<const unsigned long long*>(src + i + lane)) : src[i + lane];
            if (__ballot_sync(FULL, cand > m.thr) == 0) break;  // lists are sorted: nothing further can enter
            m.offer(cand, lane, k);
        }
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

    if (tid == 0) {
        for (int s = 0; s < NWARPS; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_fence_init();
    }
    // query normalisation fused here (dbsearch.py:78 cosine eps 1e-8 / dbsearch.py:304 F.normalize eps 1e-12)
    if (warp < NQ) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (warp < p.nq) v = reinterpret_cast<const float4*>(p.q + size_t(warp) * DIM)[lane];
        if (p.qnorm != FCS_QNORM_NONE) {
            float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(FULL, ss, o);
            const float eps = (p.qnorm == FCS_QNORM_COSINE) ? 1e-8f : 1e-12f;
            const float d = fmaxf(sqrtf(ss), eps);
            v.x = v.x / d; v.y = v.y / d; v.z = v.z / d; v.w = v.w / d;
        }
        reinterpret_cast<float4*>(qs + warp * DIM)[lane] = v;
    }
    __syncthreads();

    const int64_t n_chunks = (p.n_rows + STAGE_ROWS - 1) / STAGE_ROWS;

    WarpTopK<KPL> tk[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) tk[q].init();

    if (warp == NWARPS) {
        // ------------------------------------------------------------------ producer
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
        float4 qv[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) qv[q] = reinterpret_cast<const float4*>(qs + q * DIM)[lane];
        uint64_t ub[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            ub[q] = ~0ull;
            if (p.bounded && q < p.nq) ub[q] = p.out_keys[size_t(q) * p.out_stride + p.out_off - 1];
        }
        const float4* st = reinterpret_cast<const float4*>(ring + warp * STAGE_BYTES);
        uint32_t j = 0;
        for (int64_t it = warp;; it += NWARPS, ++j) {
            const int64_t c = blockIdx.x + it * gridDim.x;
            if (c >= n_chunks) break;
            const int64_t my_row = c * STAGE_ROWS + lane;
            const bool valid = my_row < p.n_rows;
            float lenf = 0.f;
            if (p.use_mask && valid) lenf = float(__ldg(p.lens + my_row));
            mbar_wait(&full_bar[warp], j & 1u);

            constexpr int QP = NQ < 2 ? NQ : 2;  // queries per pass over the stage
#pragma unroll
            for (int q0 = 0; q0 < NQ; q0 += QP) {
                float acc[QP][32];
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const float4 v = st[r * 32 + lane];
#pragma unroll
                    for (int qq = 0; qq < QP; ++qq) {
                        const float4 w = qv[q0 + qq];
                        acc[qq][r] = fmaf(v.w, w.w, fmaf(v.z, w.z, fmaf(v.y, w.y, v.x * w.x)));
                    }
                }
                if (q0 + QP >= NQ) {  // last pass over this stage: hand it back to the producer
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[warp]);
                }
#pragma unroll
                for (int qq = 0; qq < QP; ++qq) {
                    transpose_reduce32(acc[qq], lane);
                    float score = acc[qq][0];
                    if (p.use_mask) {
                        // dbsearch.py:76: (qlen >= lengths * mincov).float(), an fp32 product
                        const float need = __fmul_rn(lenf, p.mincov);
                        score = score * ((p.qlen[q0 + qq] >= need) ? 1.0f : 0.0f);
                    }
                    uint64_t cand = make_key(score, p.id_base + uint32_t(my_row));
                    cand = (valid && cand < ub[q0 + qq]) ? cand : 0ull;
                    tk[q0 + qq].offer(cand, lane, k);
                }
            }
        }
    }

    // ---------------------------------------------------------------------- CTA-level merge
    __syncthreads();  // every bulk copy has landed and been consumed: the ring can be reused
    uint64_t* lists = reinterpret_cast<uint64_t*>(ring);  // [NWARPS][NQ][k]
    if (warp < NWARPS) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) tk[q].store(lists + (size_t(warp) * NQ + q) * k, lane, k);
    }
    __syncthreads();
    if (warp < NQ) {
        WarpTopK<KPL> m;
        m.init();
        merge_lists<KPL, false>(m, lists + size_t(warp) * k, 0, 1, NWARPS, size_t(NQ) * k, k, lane);
        m.store(p.scratch + (size_t(blockIdx.x) * NQ + warp) * k, lane, k);
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
    if (warp < NWARPS) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            WarpTopK<KPL> m;
            m.init();
            merge_lists<KPL, true>(m, p.scratch + size_t(q) * k, warp, NWARPS, int(gridDim.x), size_t(NQ) * k, k, lane);
            m.store(lists + (size_t(warp) * NQ + q) * k, lane, k);
        }
    }
    __syncthreads();
    if (warp < NQ && warp < p.nq) {
        WarpTopK<KPL> m;
        m.init();
        merge_lists<KPL, false>(m, lists + size_t(warp) * k, 0, 1, NWARPS, size_t(NQ) * k, k, lane);
        const size_t base = size_t(warp) * p.out_stride + p.out_off;
#pragma unroll
        for (int jj = 0; jj < KPL; ++jj) {
            const int r = jj * 32 + lane;
            if (r < k) {
                const uint64_t key = m.key[jj];
                p.out_keys[base + r] = key;
                if (p.out_scores) p.out_scores[base + r] = key_score(key);
                if (p.out_ids) p.out_ids[base + r] = key_id(key);
            }
        }
    }
    if (tid == 0) *p.ticket = 0u;  // ready for the next launch on this stream
}

template <int NQ, int KPL>
cudaError_t launch_inst(const GemvParams& p, int grid, cudaStream_t stream) {
    gemv_topk_kernel<NQ, KPL><<<grid, NTHREADS, SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

template <int NQ, int KPL>
cudaError_t configure_inst() {
    return cudaFuncSetAttribute(gemv_topk_kernel<NQ, KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
}

}  // namespace

size_t gemv_scratch_bytes(int max_grid) { return size_t(max_grid) * GEMV_MAX_NQ * GEMV_MAX_K * sizeof(uint64_t); }

cudaError_t gemv_configure() {
    cudaError_t e;
#define FCS_CFG(NQ, KPL) \
    if ((e = configure_inst<NQ, KPL>()) != cudaSuccess) return e;
    FCS_CFG(1, 1) FCS_CFG(1, 2) FCS_CFG(1, 4)
    FCS_CFG(2, 1) FCS_CFG(2, 2) FCS_CFG(2, 4)
    FCS_CFG(4, 1) FCS_CFG(4, 2) FCS_CFG(4, 4)
#undef FCS_CFG
    return cudaSuccess;
}

cudaError_t gemv_launch(const GemvParams& p, int sm_count, cudaStream_t stream) {
    if (p.nq < 1 || p.nq > GEMV_MAX_NQ || p.k < 1 || p.k > GEMV_MAX_K || p.n_rows < 1) return cudaErrorInvalidValue;
    const int64_t n_chunks = (p.n_rows + GEMV_STAGE_ROWS - 1) / GEMV_STAGE_ROWS;
    const int grid = int(n_chunks < sm_count ? n_chunks : sm_count);
    const int nqt = p.nq == 1 ? 1 : (p.nq == 2 ? 2 : 4);
    const int kpl = p.k <= 32 ? 1 : (p.k <= 64 ? 2 : 4);
#define FCS_GO(NQ, KPL) \
    if (nqt == NQ && kpl == KPL) return launch_inst<NQ, KPL>(p, grid, stream);
    FCS_GO(1, 1) FCS_GO(1, 2) FCS_GO(1, 4)
    FCS_GO(2, 1) FCS_GO(2, 2) FCS_GO(2, 4)
    FCS_GO(4, 1) FCS_GO(4, 2) FCS_GO(4, 4)
#undef FCS_GO
    return cudaErrorInvalidValue;
}

}  // namespace fcs
