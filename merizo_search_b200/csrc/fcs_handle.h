// fcs_handle.h -- the database handle behind the C ABI (struct fcs_db) and the internal entry points shared by
// fcs_api.cu (one shard on one device) and fcs_group.cu (several shards driven from one host thread).
#pragma once

#include <string>
#include <utility>
#include <vector>

#include "fcs_internal.h"
#include "fcs_tc.h"

namespace fcs {

// error code + thread-local message (fcs_last_error); never throws
int api_fail(int code, const char* fmt, ...);

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace fcs

#define FCS_FAIL(...) fcs::api_fail(__VA_ARGS__)
#define FCS_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            const int code__ = (e__ == cudaErrorMemoryAllocation) ? FCS_ERR_NOMEM : FCS_ERR_CUDA;        \
            (void)cudaGetLastError();                                                                    \
            return fcs::api_fail(code__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                                \
    } while (0)

// ------------------------------------------------------------------------------------ handle
struct fcs_db {
    int device = 0;
    int64_t n_rows = 0;
    int64_t id_offset = 0;
    uint32_t flags = 0;
    bool finalized = false;
    int sm_count = 0;

    float* rows = nullptr;         // [n_rows,128] fp32, row-swizzled after finalize
    uint16_t* lens = nullptr;      // [n_rows] (optional)

    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t out_ev = nullptr;  // fcs_search: scores have arrived on the host (their copy to the caller overlaps the ids' D2H)
    bool ev_valid = false;
    bool profiling = false;  // record ev0/ev1 around every search (off: event records between two scan kernels
                             // would break their programmatic-dependent-launch overlap)

    uint64_t* gemv_scratch = nullptr;
    unsigned* ticket = nullptr;
    int* d_bad = nullptr;

    // buffers behind the host-pointer API (grown on demand)
    float* d_q = nullptr;
    size_t d_q_cap = 0;  // queries
    uint64_t* d_keys = nullptr;
    float* d_scores = nullptr;
    int64_t* d_ids = nullptr;
    size_t d_out_cap = 0;  // entries
    float* h_q = nullptr;
    size_t h_q_cap = 0;
    float* h_scores = nullptr;
    int64_t* h_ids = nullptr;
    uint64_t* h_keys = nullptr;  // only for the zero-copy small-result path
    size_t h_out_cap = 0;

    void* h_stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    int stage_next = 0;

    fcs::TcState* tc = nullptr;
    fcs_timing timing = {};

    // Tensor-core searches queue the queries that failed their certificate on the device; FB_ASYNC_PASSES indirect
    // scan launches (8 queries each) follow every search.  What is needed to finish a longer queue later:
    struct Pending {
        bool valid = false;
        fcs::TcFallbackQueue q;
        int k = 0, qnorm = 0;
        float* out_scores = nullptr;
        int64_t* out_ids = nullptr;
        uint64_t* out_keys = nullptr;
    } pending;

    // rows uploaded so far, as disjoint sorted [lo, hi) ranges (finalize checks that they cover the shard)
    std::vector<std::pair<int64_t, int64_t>> covered;
};

namespace fcs {

constexpr int FB_ASYNC_PASSES = 4;  // exact-scan launches enqueued behind every tensor-core search (32 queries)

int api_ensure_query_bufs(fcs_db* db, size_t nq, bool host);
int api_ensure_out_bufs(fcs_db* db, size_t entries, bool host);
int api_check_search(const fcs_db* db, const void* q, int nq, int k, int qnorm, int mode, const char* fn);
// enqueue one search of the shard on `stream` (device pointers; no synchronisation)
int api_search_core(fcs_db* db, const float* q_dev, int nq, const int32_t* qlen, float mincov, int k, int qnorm, int mode,
                    int kprime, float* out_scores, int64_t* out_ids, uint64_t* out_keys, cudaStream_t stream);
// `stream` is idle: complete a fallback queue longer than the passes enqueued behind the last tensor-core search
int api_finish_pending(fcs_db* db, cudaStream_t stream, int* n_queued);
bool api_auto_prefers_tc(const fcs_db* db, int nq, int k, bool mask_on);

}  // namespace fcs
