// fcs_group.cu -- several row shards behind ONE handle, driven by ONE host thread (the single-process
// multi-GPU deployment: what `merizo.py search -d cuda` gets on a multi-GPU node).
//
// Replaces faiss.index_cpu_to_all_gpus + ResultHeap (reference dbsearch.py:228-245): the shards are contiguous
// row ranges (global id = shard offset + local row, the reference's `I += i0`, dbsearch.py:238), every shard
// searches the replicated queries on its own GPU and stream, and the per-shard sorted key lists ([nq,k] packed
// 64-bit keys, 8 bytes per entry) travel device-to-device over NVLink (cudaMemcpyPeerAsync issued on the producing
// stream) into one [shards][nq][k] buffer on the first GPU, where K5 (merge_topk_kernel) ranks them.  Nothing in a
// search waits for the host until the final result copy: the per-shard searches are asynchronous (fcs_api.cu), so
// one thread keeps all GPUs busy.  The loader feeds every shard from its own host thread (one PCIe link per GPU).
//
// Several shards may live on the same device (`devices` may repeat an ordinal): that is how the exchange + merge
// path is exercised on a single-GPU box.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "fcs_common.cuh"
#include "fcs_handle.h"

using namespace fcs;

struct fcs_group {
    int n_shards = 0;
    int64_t n_rows = 0;
    uint32_t flags = 0;
    std::vector<fcs_db*> shards;
    std::vector<int> devices;
    std::vector<int64_t> bounds;  // n_shards + 1 row offsets
    std::vector<cudaEvent_t> ev;  // one per shard, created on the shard's device
    // gather + merge buffers on devices[0], result staging in pinned host memory
    uint64_t* g_keys = nullptr;
    float* g_scores = nullptr;
    int64_t* g_ids = nullptr;
    size_t g_cap = 0;  // entries (nq*k) the buffers hold
    float* h_q = nullptr;
    size_t h_q_cap = 0;
    float* h_scores = nullptr;
    int64_t* h_ids = nullptr;
    size_t h_out_cap = 0;
    int last_queued = 0;  // fallback-queue length summed over the shards of the last search
};

namespace {

int ensure_group_bufs(fcs_group* g, size_t nq, size_t entries) {
    DeviceGuard guard(g->devices[0]);
    if (g->g_cap < entries) {
        cudaFree(g->g_keys); cudaFree(g->g_scores); cudaFree(g->g_ids);
        g->g_keys = nullptr; g->g_scores = nullptr; g->g_ids = nullptr;
        g->g_cap = 0;
        FCS_CUDA(cudaMalloc(&g->g_keys, size_t(g->n_shards) * entries * sizeof(uint64_t)));
        FCS_CUDA(cudaMalloc(&g->g_scores, entries * sizeof(float)));
        FCS_CUDA(cudaMalloc(&g->g_ids, entries * sizeof(int64_t)));
        g->g_cap = entries;
    }
    if (g->h_out_cap < entries) {
        if (g->h_scores) cudaFreeHost(g->h_scores);
        if (g->h_ids) cudaFreeHost(g->h_ids);
        g->h_scores = nullptr; g->h_ids = nullptr;
        g->h_out_cap = 0;
        FCS_CUDA(cudaMallocHost(&g->h_scores, entries * sizeof(float)));
        FCS_CUDA(cudaMallocHost(&g->h_ids, entries * sizeof(int64_t)));
        g->h_out_cap = entries;
    }
    if (g->h_q_cap < nq) {
        if (g->h_q) cudaFreeHost(g->h_q);
        g->h_q = nullptr;
        g->h_q_cap = 0;
        FCS_CUDA(cudaMallocHost(&g->h_q, nq * DIM * sizeof(float)));
        g->h_q_cap = nq;
    }
    return FCS_OK;
}

// shard s: its key list -> slot s of the gather buffer on the first device, on the shard's own stream
int gather_keys(fcs_group* g, int s, size_t entries) {
    fcs_db* sh = g->shards[s];
    DeviceGuard guard(sh->device);
    FCS_CUDA(cudaMemcpyPeerAsync(g->g_keys + size_t(s) * entries, g->devices[0], sh->d_keys, sh->device, entries * sizeof(uint64_t),
                                 sh->stream));
    FCS_CUDA(cudaEventRecord(g->ev[s], sh->stream));
    return FCS_OK;
}

int merge_and_fetch(fcs_group* g, int nq, int k) {
    const size_t entries = size_t(nq) * k;
    DeviceGuard guard(g->devices[0]);
    cudaStream_t st0 = g->shards[0]->stream;
    for (int s = 1; s < g->n_shards; ++s) FCS_CUDA(cudaStreamWaitEvent(st0, g->ev[s], 0));
    FCS_CUDA(merge_topk_launch(g->g_keys, g->n_shards, nq, k, g->g_scores, g->g_ids, nullptr, st0));
    FCS_CUDA(cudaMemcpyAsync(g->h_scores, g->g_scores, entries * sizeof(float), cudaMemcpyDeviceToHost, st0));
    FCS_CUDA(cudaMemcpyAsync(g->h_ids, g->g_ids, entries * sizeof(int64_t), cudaMemcpyDeviceToHost, st0));
    FCS_CUDA(cudaStreamSynchronize(st0));
    return FCS_OK;
}

}  // namespace

extern "C" int fcs_group_create(const int* devices, int n_shards, int64_t n_rows, int dim, uint32_t flags, fcs_group** out) {
    if (!out) return FCS_FAIL(FCS_ERR_INVALID, "fcs_group_create: out is NULL");
    *out = nullptr;
    if (!devices || n_shards < 1) return FCS_FAIL(FCS_ERR_INVALID, "fcs_group_create: at least one shard device is required");
    if (n_rows < n_shards) return FCS_FAIL(FCS_ERR_INVALID, "fcs_group_create: %lld rows cannot fill %d shards", (long long)n_rows, n_shards);
    fcs_group* g = new (std::nothrow) fcs_group();
    if (!g) return FCS_FAIL(FCS_ERR_NOMEM, "fcs_group_create: out of host memory");
    g->n_shards = n_shards;
    g->n_rows = n_rows;
    g->flags = flags;
    g->devices.assign(devices, devices + n_shards);
    // contiguous ranges: shard s holds rows [s*ceil(N/G), min(N, (s+1)*ceil(N/G)))  
    const int64_t per = (n_rows + n_shards - 1) / n_shards;
    g->bounds.resize(n_shards + 1);
    for (int s = 0; s <= n_shards; ++s) g->bounds[s] = (int64_t(s) * per < n_rows) ? int64_t(s) * per : n_rows;
    if (g->bounds[n_shards - 1] >= n_rows) {
        delete g;
        return FCS_FAIL(FCS_ERR_INVALID, "fcs_group_create: %lld rows leave shard %d of %d empty", (long long)n_rows, n_shards - 1, n_shards);
    }
    g->ev.assign(n_shards, nullptr);
    for (int s = 0; s < n_shards; ++s) {
        fcs_db* sh = nullptr;
        int rc = fcs_db_create(devices[s], g->bounds[s + 1] - g->bounds[s], dim, g->bounds[s], flags, &sh);
        if (rc == FCS_OK) {
            g->shards.push_back(sh);
            DeviceGuard guard(devices[s]);
            if (cudaEventCreateWithFlags(&g->ev[s], cudaEventDisableTiming) != cudaSuccess) rc = FCS_FAIL(FCS_ERR_CUDA, "fcs_group_create: event creation failed");
        }
        if (rc != FCS_OK) {
            const std::string keep = fcs_last_error();
            fcs_group_destroy(g);
            set_last_error(keep.c_str());
            return rc;
        }
    }
    // direct NVLink paths between the first device (where the lists are merged) and the others, where available
    for (int s = 1; s < n_shards; ++s) {
        if (devices[s] == devices[0]) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[s], devices[0]) == cudaSuccess && can) {
            DeviceGuard guard(devices[s]);
            (void)cudaDeviceEnablePeerAccess(devices[0], 0);
        }
        if (cudaDeviceCanAccessPeer(&can, devices[0], devices[s]) == cudaSuccess && can) {
            DeviceGuard guard(devices[0]);
            (void)cudaDeviceEnablePeerAccess(devices[s], 0);
        }
        (void)cudaGetLastError();  // "already enabled" is fine
    }
    *out = g;
    return FCS_OK;
}

extern "C" int fcs_group_destroy(fcs_group* g) {
    if (!g) return FCS_OK;
    for (fcs_db* sh : g->shards) {
        if (sh && sh->stream) {
            DeviceGuard guard(sh->device);
            cudaStreamSynchronize(sh->stream);
        }
    }
    for (size_t s = 0; s < g->ev.size(); ++s) {
        if (g->ev[s]) {
            DeviceGuard guard(g->devices[s]);
            cudaEventDestroy(g->ev[s]);
        }
    }
    if (!g->devices.empty()) {
        DeviceGuard guard(g->devices[0]);
        cudaFree(g->g_keys); cudaFree(g->g_scores); cudaFree(g->g_ids);
        if (g->h_q) cudaFreeHost(g->h_q);
        if (g->h_scores) cudaFreeHost(g->h_scores);
        if (g->h_ids) cudaFreeHost(g->h_ids);
    }
    for (fcs_db* sh : g->shards) fcs_db_destroy(sh);
    (void)cudaGetLastError();
    delete g;
    return FCS_OK;
}

extern "C" int fcs_group_get_info(const fcs_group* g, int* out_n_shards, int64_t* out_bounds, int* out_devices, int max_shards) {
    if (!g) return FCS_FAIL(FCS_ERR_INVALID, "fcs_group_get_info: group is NULL");
    if (out_n_shards) *out_n_shards = g->n_shards;
    for (int s = 0; s < g->n_shards && s < max_shards; ++s) {
        if (out_bounds) {
            out_bounds[s] = g->bounds[s];
            out_bounds[s + 1] = g->bounds[s + 1];
        }
        if (out_devices) out_devices[s] = g->devices[s];
    }
    return FCS_OK;
}

extern "C" int fcs_group_shard(fcs_group* g, int index, fcs_db** out) {
    if (!g || !out || index < 0 || index >= g->n_shards) return FCS_FAIL(FCS_ERR_INVALID, "fcs_group_shard: bad argument");
    *out = g->shards[index];
    return FCS_OK;
}

// ------------------------------------------------------------------------------------ loader
namespace {
struct UploadTask {
    int shard;
    int64_t local_row0, n, src_row;  // rows [src_row, src_row+n) of the caller's block go to shard rows [local_row0, ...)
    int rc = FCS_OK;
    std::string err;
};

std::vector<UploadTask> split_upload(const fcs_group* g, int64_t row0, int64_t n) {
    std::vector<UploadTask> tasks;
    for (int s = 0; s < g->n_shards; ++s) {
        const int64_t lo = row0 > g->bounds[s] ? row0 : g->bounds[s];
        const int64_t hi = row0 + n < g->bounds[s + 1] ? row0 + n : g->bounds[s + 1];
        if (lo < hi) tasks.push_back({s, lo - g->bounds[s], hi - lo, lo - row0});
    }
    return tasks;
}

template <typename F>
int run_tasks(std::vector<UploadTask>& tasks, F&& body) {
    if (tasks.size() == 1) {
        tasks[0].rc = body(tasks[0]);
        return tasks[0].rc;  // the thread-local error message is already this thread's
    }
    std::vector<std::thread> pool;
    for (UploadTask& t : tasks)
        pool.emplace_back([&t, &body] {
            t.rc = body(t);
            if (t.rc != FCS_OK) t.err = fcs_last_error();  // thread-local: carry it over to the caller's thread
        });
    for (auto& th : pool) th.join();
    for (UploadTask& t : tasks)
        if (t.rc != FCS_OK) {
            set_last_error(t.err.c_str());
            return t.rc;
        }
    return FCS_OK;
}
}  // namespace

extern "C" int fcs_group_upload(fcs_group* g, int64_t row0, int64_t n, const float* host_rows, const int32_t* host_lengths) {
    if (!g) return FCS_FAIL(FCS_ERR_INVALID, "fcs_group_upload: group is NULL");
    if (!host_rows) return FCS_FAIL(FCS_ERR_INVALID, "fcs_group_upload: rows is NULL");
    if (row0 < 0 || n < 0 || row0 + n > g->n_rows)
        return FCS_FAIL(FCS_ERR_INVALID, "fcs_group_upload: rows [%lld, %lld) outside [0, %lld)", (long long)row0, (long long)(row0 + n), (long long)g->n_rows);
    std::vector<UploadTask> tasks = split_upload(g, row0, n);
    if (tasks.empty()) return FCS_OK;
    // one host thread per shard: staging copies and H2D DMAs of different GPUs overlap (one PCIe link each)
    return run_tasks(tasks, [&](UploadTask& t) {
        return fcs_db_upload(g->shards[t.shard], t.local_row0, t.n, host_rows + t.src_row * DIM, host_lengths ? host_lengths + t.src_row : nullptr);
    });
}

extern "C" int fcs_group_upload_file(fcs_group* g, const char* path, int64_t file_offset, int64_t row0, int64_t n) {
    if (!g || !path) return FCS_FAIL(FCS_ERR_INVALID, "fcs_group_upload_file: NULL argument");
    if (row0 < 0 || n < 0 || row0 + n > g->n_rows)
        return FCS_FAIL(FCS_ERR_INVALID, "fcs_group_upload_file: rows [%lld, %lld) outside [0, %lld)", (long long)row0, (long long)(row0 + n), (long long)g->n_rows);
    std::vector<UploadTask> tasks = split_upload(g, row0, n);
    if (tasks.empty()) return FCS_OK;
    return run_tasks(tasks, [&](UploadTask& t) {
        return fcs_db_upload_file(g->shards[t.shard], t.local_row0, t.n, path, file_offset + t.src_row * int64_t(ROW_BYTES));
    });
}

extern "C" int fcs_group_finalize(fcs_group* g) {
    if (!g) return FCS_FAIL(FCS_ERR_INVALID, "fcs_group_finalize: group is NULL");
    std::vector<UploadTask> tasks;
    for (int s = 0; s < g->n_shards; ++s) tasks.push_back({s, 0, 0, 0});
    return run_tasks(tasks, [&](UploadTask& t) { return fcs_db_finalize(g->shards[t.shard]); });
}

// ------------------------------------------------------------------------------------ search
extern "C" int fcs_group_search(fcs_group* g, const float* q, int nq, const int32_t* qlen, float mincov, int k, int qnorm, int mode,
                                int kprime, float* out_scores, int64_t* out_ids) {
    if (!g) return FCS_FAIL(FCS_ERR_INVALID, "fcs_group_search: group is NULL");
    if (g->n_shards == 1) {
        const int rc = fcs_search(g->shards[0], q, nq, qlen, mincov, k, qnorm, mode, kprime, out_scores, out_ids);
        g->last_queued = g->shards[0]->timing.last_tc_fallbacks;
        return rc;
    }
    int rc;
    for (fcs_db* sh : g->shards)
        if ((rc = api_check_search(sh, q, nq, k, qnorm, mode, "fcs_group_search")) != FCS_OK) return rc;
    if (!out_scores || !out_ids) return FCS_FAIL(FCS_ERR_INVALID, "fcs_group_search: output buffer is NULL");
    const size_t entries = size_t(nq) * k;
    if ((rc = ensure_group_bufs(g, size_t(nq), entries)) != FCS_OK) return rc;
    memcpy(g->h_q, q, size_t(nq) * DIM * sizeof(float));
    // enqueue everything: queries to every shard, shard search, key list to the first device
    for (int s = 0; s < g->n_shards; ++s) {
        fcs_db* sh = g->shards[s];
        DeviceGuard guard(sh->device);
        if ((rc = api_ensure_query_bufs(sh, size_t(nq), false)) != FCS_OK) return rc;
        if ((rc = api_ensure_out_bufs(sh, entries, false)) != FCS_OK) return rc;
        FCS_CUDA(cudaMemcpyAsync(sh->d_q, g->h_q, size_t(nq) * DIM * sizeof(float), cudaMemcpyHostToDevice, sh->stream));
        if ((rc = api_search_core(sh, sh->d_q, nq, qlen, mincov, k, qnorm, mode, kprime, nullptr, nullptr, sh->d_keys, sh->stream)) != FCS_OK) return rc;
        if ((rc = gather_keys(g, s, entries)) != FCS_OK) return rc;
    }
    if ((rc = merge_and_fetch(g, nq, k)) != FCS_OK) return rc;
    // every shard's stream work is complete (the merge waited for all of them).  A shard whose fallback queue was longer
    // than the passes enqueued behind its search finishes it now; its list is gathered and everything merged again.
    bool redo = false;
    g->last_queued = 0;
    for (int s = 0; s < g->n_shards; ++s) {
        fcs_db* sh = g->shards[s];
        DeviceGuard guard(sh->device);
        int queued = 0;
        if ((rc = api_finish_pending(sh, sh->stream, &queued)) != FCS_OK) return rc;
        g->last_queued += queued;
        if (queued > FB_ASYNC_PASSES * GEMV_MAX_NQ) {
            if ((rc = gather_keys(g, s, entries)) != FCS_OK) return rc;
            redo = true;
        }
    }
    if (redo && (rc = merge_and_fetch(g, nq, k)) != FCS_OK) return rc;
    memcpy(out_scores, g->h_scores, entries * sizeof(float));
    memcpy(out_ids, g->h_ids, entries * sizeof(int64_t));
    return FCS_OK;
}

extern "C" int fcs_group_last_fallbacks(const fcs_group* g) { return g ? g->last_queued : 0; }
