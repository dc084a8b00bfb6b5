// fcs_merge.cu -- K5: merge of per-shard top-k lists (after the NVLink all-gather).
//
// Replaces faiss.ResultHeap.add_result / finalize (reference dbsearch.py:224, 239, 245): the exact
// top-k of a union is the top-k of the per-part top-k lists.  Input: n_lists sorted key lists per
// query ([n_lists][nq][k] packed 64-bit keys with GLOBAL ids, unique across lists).  Merge by
// ranking: the final rank of an element is its position in its own list plus, for every other list,
// the number of keys greater than it (a binary search, lists are sorted).  Fully parallel, no
// shared memory, any n_lists and k; the whole problem is a few hundred KB at most.
#include "fcs_common.cuh"
#include "fcs_internal.h"

namespace fcs {

namespace {

// number of keys in sorted-descending list[0..k) that are > key
__device__ __forceinline__ int count_greater(const uint64_t* __restrict__ list, int k, uint64_t key) {
    int lo = 0, hi = k;  // first position whose key <= `key`
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (list[mid] > key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) merge_topk_kernel(const uint64_t* __restrict__ keys, int n_lists, int nq, int k,
                                                         float* __restrict__ out_scores, int64_t* __restrict__ out_ids,
                                                         uint64_t* __restrict__ out_keys) {
    const int64_t total = int64_t(n_lists) * nq * k;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int pos = int(t % k);
        const int q = int((t / k) % nq);
        const int a = int(t / (int64_t(k) * nq));
        const uint64_t key = keys[t];
        int rank;
        if (key != 0) {
            rank = pos;
            for (int b = 0; b < n_lists; ++b)
                if (b != a) rank += count_greater(keys + (int64_t(b) * nq + q) * k, k, key);
        } else {
            // empty slot: ranks after every real key; order among empties by (list, pos)
            int n_valid = 0, before = 0;
            for (int b = 0; b < n_lists; ++b) {
                const int nv = count_greater(keys + (int64_t(b) * nq + q) * k, k, 0ull);
                n_valid += nv;
                if (b < a) before += k - nv;
                if (b == a) before += pos - nv;
            }
            rank = n_valid + before;
        }
        if (rank < k) {
            const int64_t o = int64_t(q) * k + rank;
            if (out_keys) out_keys[o] = key;
            if (out_scores) out_scores[o] = key_score(key);
            if (out_ids) out_ids[o] = key_id(key);
        }
    }
}

}  // namespace

cudaError_t merge_topk_launch(const uint64_t* keys, int n_lists, int nq, int k, float* out_scores, int64_t* out_ids,
                              uint64_t* out_keys, cudaStream_t stream) {
    const int64_t total = int64_t(n_lists) * nq * k;
    if (total <= 0) return cudaErrorInvalidValue;
    int64_t g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    merge_topk_kernel<<<int(g), 256, 0, stream>>>(keys, n_lists, nq, k, out_scores, out_ids, out_keys);
    return cudaGetLastError();
}

}  // namespace fcs
