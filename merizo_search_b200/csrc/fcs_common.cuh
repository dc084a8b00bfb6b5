// fcs_common.cuh -- shared device helpers: packed sort keys, warp-resident top-k lists,
// mbarrier / bulk-copy (TMA) PTX wrappers.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fcs {

constexpr int DIM = 128;
constexpr int ROW_BYTES = DIM * 4;
constexpr unsigned FULL = 0xffffffffu;

// --------------------------------------------------------------------------------------
// Packed sort key: high 32 bits = order-preserving image of the fp32 score, low 32 bits =
// 0xFFFFFFFF - row id.  Bigger key == better hit (higher score; ties -> lower id), so one
// unsigned 64-bit compare orders hits deterministically.  Key 0 == empty slot: it is below
// every real key (the low word of a real key is >= 1 because ids are < 0xFFFFFFFF) and
// decodes to (-inf, -1), faiss's padding convention.
// --------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t order_f32(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b;
    memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float unorder_f32(uint32_t u) {
    uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t id) {
    return (uint64_t(order_f32(score)) << 32) | uint64_t(0xFFFFFFFFu - id);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key) {
    return key == 0 ? -INFINITY : unorder_f32(uint32_t(key >> 32));
}
__host__ __device__ __forceinline__ int64_t key_id(uint64_t key) {
    return key == 0 ? int64_t(-1) : int64_t(0xFFFFFFFFu - uint32_t(key));
}

// --------------------------------------------------------------------------------------
// Row storage ("row-swizzled fp32"): a row is 32 chunks of 16 B.  Chunk c of local row r is
// stored at chunk position c ^ (r & 7) of the same row, i.e. the 8 chunks of every aligned
// 128 B group are permuted by the row's low 3 bits.  A 32-row tile copied verbatim into shared
// memory can then be read one ROW PER LANE with conflict-free 128-bit loads (8 consecutive lanes
// hit 8 different 16 B bank groups) -- no cross-lane reduction is needed for the dot product.
// --------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int swz_chunk(int chunk, int64_t row) { return chunk ^ int(row & 7); }

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(FULL, uint32_t(v), src);
    uint32_t hi = __shfl_sync(FULL, uint32_t(v >> 32), src);
    return (uint64_t(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int d) {
    uint32_t lo = __shfl_up_sync(FULL, uint32_t(v), d);
    uint32_t hi = __shfl_up_sync(FULL, uint32_t(v >> 32), d);
    return (uint64_t(hi) << 32) | lo;
}

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    uint32_t lo = __shfl_xor_sync(FULL, uint32_t(v), m);
    uint32_t hi = __shfl_xor_sync(FULL, uint32_t(v >> 32), m);
    return (uint64_t(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t umax64(uint64_t a, uint64_t b) { return a > b ? a : b; }
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a > b ? b : a; }

// Bitonic sort of 32 keys (one per lane) into descending order by lane: 15 compare-exchange stages.
__device__ __forceinline__ uint64_t warp_sort_desc(uint64_t v, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride >= 1; stride >>= 1) {
            const uint64_t o = shfl_xor_u64(v, stride);
            const bool desc_block = (lane & size) == 0;  // size == 32: always true
            const bool lower = (lane & stride) == 0;
            v = (lower == desc_block) ? umax64(v, o) : umin64(v, o);
        }
    }
    return v;
}

// --------------------------------------------------------------------------------------
// Warp-resident sorted top-k list.  Rank r (0 = best) lives in lane (r % 32), register
// slot (r / 32); capacity 32*KPL >= k.  `thr` is the key at rank k-1 (0 while the list
// holds fewer than k hits): a candidate enters only if its key beats thr, which after a
// short warm-up is rare (about k*ln(n/k) times over n rows), so the streaming loop pays
// one compare + one ballot per 32 rows.  One or two entrants are inserted by shifting;
// more are sorted across the warp and bitonic-merged into the list, and sorted lists are
// merged the same way (max(A[e], B[n-1-e]) is bitonic and holds the n best of the union).
// --------------------------------------------------------------------------------------
template <int KPL>
struct WarpTopK {
    uint64_t key[KPL];
    uint64_t thr;

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < KPL; ++j) key[j] = 0;
        thr = 0;
    }

    __device__ __forceinline__ void update_thr(int k) {
        const int slot = (k - 1) >> 5;
        uint64_t sel = key[0];
#pragma unroll
        for (int j = 1; j < KPL; ++j) sel = (slot == j) ? key[j] : sel;
        thr = shfl_u64(sel, (k - 1) & 31);
    }

    // All 32 lanes call with the same nk.  Inserts nk at its sorted position; the last rank falls off.
    __device__ __forceinline__ void insert(uint64_t nk, int lane, int k) {
        int p = 0;
#pragma unroll
        for (int j = 0; j < KPL; ++j) p += __popc(__ballot_sync(FULL, key[j] > nk));
#pragma unroll
        for (int j = KPL - 1; j >= 0; --j) {
            uint64_t up = shfl_up_u64(key[j], 1);
            if (j > 0) {
                uint64_t wrap = shfl_u64(key[j - 1], 31);
                if (lane == 0) up = wrap;
            }
            const int r = j * 32 + lane;
            key[j] = (r > p) ? up : ((r == p) ? nk : key[j]);
        }
        update_thr(k);
    }

    // Merge a descending-sorted list (same rank layout; slots >= n_slots are empty) into this one.
    template <int OSLOTS>
    __device__ __forceinline__ void merge_sorted(const uint64_t (&other)[OSLOTS], int lane, int k) {
        static_assert(OSLOTS <= KPL, "other list larger than this one");
        // C[e] = max(A[e], B[n-1-e]); B's slot (KPL-1-j) is empty for KPL-1-j >= OSLOTS
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            if (KPL - 1 - j < OSLOTS) {
                const uint64_t rev = shfl_u64(other[KPL - 1 - j], 31 - lane);
                key[j] = umax64(key[j], rev);
            }
        }
        // bitonic merge, descending, over n = 32*KPL elements; strides >= 32 are register-to-register
#pragma unroll
        for (int sj = KPL / 2; sj >= 1; sj >>= 1) {
#pragma unroll
            for (int j = 0; j < KPL; ++j) {
                if ((j & sj) == 0) {
                    const uint64_t a = key[j], b = key[j + sj];
                    key[j] = umax64(a, b);
                    key[j + sj] = umin64(a, b);
                }
            }
        }
#pragma unroll
        for (int stride = 16; stride >= 1; stride >>= 1) {
#pragma unroll
            for (int j = 0; j < KPL; ++j) {
                const uint64_t o = shfl_xor_u64(key[j], stride);
                key[j] = ((lane & stride) == 0) ? umax64(key[j], o) : umin64(key[j], o);
            }
        }
        update_thr(k);
    }

    // Each lane offers one candidate (cand == 0 -> none).  Warp-uniform control flow.
    __device__ __forceinline__ void offer(uint64_t cand, int lane, int k) {
        unsigned m = __ballot_sync(FULL, cand > thr);
        if (m == 0) return;
        if (__popc(m) <= 2) {
            while (m) {
                const int src = __ffs(m) - 1;
                const uint64_t nk = shfl_u64(cand, src);
                insert(nk, lane, k);
                m &= m - 1;
                m &= __ballot_sync(FULL, cand > thr);
            }
        } else {
            uint64_t one[1];
            one[0] = warp_sort_desc(cand > thr ? cand : 0ull, lane);
            merge_sorted<1>(one, lane, k);
        }
    }

    // store the first k ranks
    __device__ __forceinline__ void store(uint64_t* dst, int lane, int k) const {
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            const int r = j * 32 + lane;
            if (r < k) dst[r] = key[j];
        }
    }
    // load a stored list of k ranks (ranks >= k read as empty); CG: bypass L1 (data written by other CTAs)
    template <bool CG>
    __device__ __forceinline__ void load(const uint64_t* src, int lane, int k) {
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            const int r = j * 32 + lane;
            uint64_t v = 0;
            if (r < k) v = CG ? __ldcg(reinterpret_cast<const unsigned long long*>(src + r)) : src[r];
            key[j] = v;
        }
    }
};

// ---------------------------------------------------------------------------- mbarrier / TMA
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk asynchronous copy global -> shared (the TMA engine's non-tensor form; SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.  Completion is signalled on `bar` (complete_tx).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                         uint64_t cache_policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(cache_policy)
        : "memory");
}
// Programmatic dependent launch (PDL): a kernel launched with programmaticStreamSerialization may start while
// its predecessor in the stream is still finishing; it must not touch anything the predecessor writes before
// pdl_wait(), and the predecessor lets dependents start early with pdl_launch_dependents().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// named barriers (id 1..15): arrive does not block, sync does
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// L2 prefetch of a contiguous range (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#endif  // __CUDACC__

}  // namespace fcs
