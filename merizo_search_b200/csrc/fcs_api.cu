// fcs_api.cu -- the C ABI declared in include/fcsearch.h: handle, loader, search dispatch.
// No torch types, no exceptions across the boundary, no exit(): errors are codes + fcs_last_error().
#include <fcntl.h>
#include <unistd.h>

#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <algorithm>
#include <thread>
#include <utility>
#include <vector>

#include "fcs_common.cuh"
#include "fcs_handle.h"

using namespace fcs;

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

int fcs::api_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
void fcs::set_last_error(const char* msg) { g_last_error = msg ? msg : ""; }
namespace {
constexpr size_t STAGE_BYTES_HOST = size_t(32) << 20;  // pinned staging buffers for fcs_db_upload

// pageable (or memory-mapped) -> pinned staging copy; one thread tops out near 10 GB/s, well under a PCIe 5 x16 link
void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    unsigned hw = std::thread::hardware_concurrency();
    const size_t nthreads = bytes < (size_t(4) << 20) ? 1 : (hw >= 8 ? 4 : (hw >= 4 ? 2 : 1));
    if (nthreads == 1) {
        memcpy(dst, src, bytes);
        return;
    }
    const size_t per = ((bytes / nthreads) + 4095) & ~size_t(4095);
    std::vector<std::thread> pool;
    for (size_t t = 1; t < nthreads; ++t) {
        const size_t off = t * per;
        if (off >= bytes) break;
        const size_t len = (off + per < bytes) ? per : (bytes - off);
        pool.emplace_back([=] { memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len); });
    }
    memcpy(dst, src, per < bytes ? per : bytes);
    for (auto& th : pool) th.join();
}

// file -> pinned staging: positional reads split over a few threads (page-cache copies or direct I/O both scale with
// the number of readers); returns 0 or an errno
int parallel_pread(int fd, void* dst, size_t bytes, int64_t offset) {
    unsigned hw = std::thread::hardware_concurrency();
    const size_t nthreads = bytes < (size_t(4) << 20) ? 1 : (hw >= 16 ? 8 : (hw >= 8 ? 4 : (hw >= 4 ? 2 : 1)));
    const size_t per = ((bytes / nthreads) + 4095) & ~size_t(4095);
    std::vector<int> err(nthreads, 0);
    auto body = [&](size_t t) {
        size_t off = t * per;
        if (off >= bytes) return;
        size_t end = (off + per < bytes) ? off + per : bytes;
        while (off < end) {
            const ssize_t got = pread(fd, static_cast<char*>(dst) + off, end - off, offset + int64_t(off));
            if (got < 0) {
                if (errno == EINTR) continue;
                err[t] = errno;
                return;
            }
            if (got == 0) {
                err[t] = ENODATA;  // file shorter than the rows asked for
                return;
            }
            off += size_t(got);
        }
    };
    std::vector<std::thread> pool;
    for (size_t t = 1; t < nthreads; ++t) pool.emplace_back(body, t);
    body(0);
    for (auto& th : pool) th.join();
    for (int e : err)
        if (e) return e;
    return 0;
}

}  // namespace

static void mark_covered(fcs_db* db, int64_t lo, int64_t hi) {
    auto& v = db->covered;
    v.emplace_back(lo, hi);
    std::sort(v.begin(), v.end());
    size_t w = 0;
    for (size_t i = 1; i < v.size(); ++i) {
        if (v[i].first <= v[w].second) v[w].second = std::max(v[w].second, v[i].second);
        else v[++w] = v[i];
    }
    v.resize(w + 1);
}

int fcs::api_ensure_query_bufs(fcs_db* db, size_t nq, bool host) {
    if (db->d_q_cap < nq) {
        if (db->d_q) cudaFree(db->d_q);
        db->d_q = nullptr;
        db->d_q_cap = 0;
        FCS_CUDA(cudaMalloc(&db->d_q, nq * DIM * sizeof(float)));
        db->d_q_cap = nq;
    }
    if (host && db->h_q_cap < nq) {
        if (db->h_q) cudaFreeHost(db->h_q);
        db->h_q = nullptr;
        db->h_q_cap = 0;
        FCS_CUDA(cudaMallocHost(&db->h_q, nq * DIM * sizeof(float)));
        db->h_q_cap = nq;
    }
    return FCS_OK;
}

int fcs::api_ensure_out_bufs(fcs_db* db, size_t entries, bool host) {
    if (db->d_out_cap < entries) {
        if (db->d_keys) cudaFree(db->d_keys);
        if (db->d_scores) cudaFree(db->d_scores);
        if (db->d_ids) cudaFree(db->d_ids);
        db->d_keys = nullptr; db->d_scores = nullptr; db->d_ids = nullptr;
        db->d_out_cap = 0;
        FCS_CUDA(cudaMalloc(&db->d_keys, entries * sizeof(uint64_t)));
        FCS_CUDA(cudaMalloc(&db->d_scores, entries * sizeof(float)));
        FCS_CUDA(cudaMalloc(&db->d_ids, entries * sizeof(int64_t)));
        db->d_out_cap = entries;
    }
    if (host && db->h_out_cap < entries) {
        if (db->h_scores) cudaFreeHost(db->h_scores);
        if (db->h_ids) cudaFreeHost(db->h_ids);
        if (db->h_keys) cudaFreeHost(db->h_keys);
        db->h_scores = nullptr; db->h_ids = nullptr; db->h_keys = nullptr;
        db->h_out_cap = 0;
        FCS_CUDA(cudaMallocHost(&db->h_scores, entries * sizeof(float)));
        FCS_CUDA(cudaMallocHost(&db->h_ids, entries * sizeof(int64_t)));
        FCS_CUDA(cudaMallocHost(&db->h_keys, entries * sizeof(uint64_t)));
        db->h_out_cap = entries;
    }
    return FCS_OK;
}

// ------------------------------------------------------------------------------------ library
extern "C" int fcs_version(void) { return 100; }
extern "C" const char* fcs_last_error(void) { return g_last_error.c_str(); }

extern "C" int fcs_device_count(int* out_count) {
    if (!out_count) return FCS_FAIL(FCS_ERR_INVALID, "fcs_device_count: out_count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        *out_count = 0;
        return FCS_FAIL(FCS_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    *out_count = n;
    return FCS_OK;
}

// ------------------------------------------------------------------------------------ create / destroy
extern "C" int fcs_db_create(int device, int64_t n_rows, int dim, int64_t id_offset, uint32_t flags, fcs_db** out) {
    if (!out) return FCS_FAIL(FCS_ERR_INVALID, "fcs_db_create: out is NULL");
    *out = nullptr;
    if (dim != DIM) return FCS_FAIL(FCS_ERR_INVALID, "fcs_db_create: dim must be %d (got %d)", DIM, dim);
    if (n_rows < 1) return FCS_FAIL(FCS_ERR_INVALID, "fcs_db_create: n_rows must be >= 1 (got %lld)", (long long)n_rows);
    if (id_offset < 0 || id_offset + n_rows >= int64_t(0xFFFFFFFFll))
        return FCS_FAIL(FCS_ERR_INVALID, "fcs_db_create: global ids must stay below 2^32-1 (offset %lld + rows %lld)",
                    (long long)id_offset, (long long)n_rows);
    if (flags & ~(FCS_DB_NORMALISE_ROWS | FCS_DB_KEEP_BF16 | FCS_DB_HAS_LENGTHS))
        return FCS_FAIL(FCS_ERR_INVALID, "fcs_db_create: unknown flag bits 0x%x", flags);
    int ndev = 0;
    FCS_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return FCS_FAIL(FCS_ERR_INVALID, "fcs_db_create: device %d out of range (%d GPUs)", device, ndev);
    DeviceGuard guard(device);
    if (!guard.ok) return FCS_FAIL(FCS_ERR_CUDA, "fcs_db_create: cudaSetDevice(%d) failed", device);
    cudaDeviceProp prop;
    FCS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return FCS_FAIL(FCS_ERR_UNSUPPORTED, "fcs_db_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    fcs_db* db = new (std::nothrow) fcs_db();
    if (!db) return FCS_FAIL(FCS_ERR_NOMEM, "fcs_db_create: out of host memory");
    db->device = device;
    db->n_rows = n_rows;
    db->id_offset = id_offset;
    db->flags = flags;
    db->sm_count = prop.multiProcessorCount;
    int rc = FCS_OK;
    auto run = [&]() -> int {
        FCS_CUDA(gemv_configure());
        FCS_CUDA(cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking));
        FCS_CUDA(cudaEventCreate(&db->ev0));
        FCS_CUDA(cudaEventCreate(&db->ev1));
        FCS_CUDA(cudaEventCreateWithFlags(&db->out_ev, cudaEventDisableTiming));
        FCS_CUDA(cudaMalloc(&db->rows, size_t(n_rows) * ROW_BYTES));
        if (flags & FCS_DB_HAS_LENGTHS) FCS_CUDA(cudaMalloc(&db->lens, size_t(n_rows) * sizeof(uint16_t)));
        FCS_CUDA(cudaMalloc(&db->gemv_scratch, gemv_scratch_bytes(db->sm_count)));
        FCS_CUDA(cudaMalloc(&db->ticket, sizeof(unsigned)));
        FCS_CUDA(cudaMemsetAsync(db->ticket, 0, sizeof(unsigned), db->stream));
        FCS_CUDA(cudaMalloc(&db->d_bad, sizeof(int)));
        FCS_CUDA(cudaMemsetAsync(db->d_bad, 0, sizeof(int), db->stream));
        FCS_CUDA(cudaStreamSynchronize(db->stream));
        return FCS_OK;
    };
    rc = run();
    if (rc != FCS_OK) {
        std::string keep = g_last_error;
        fcs_db_destroy(db);
        g_last_error = keep;
        return rc;
    }
    *out = db;
    return FCS_OK;
}

extern "C" int fcs_db_destroy(fcs_db* db) {
    if (!db) return FCS_OK;
    DeviceGuard guard(db->device);
    if (db->stream) cudaStreamSynchronize(db->stream);
    if (db->tc) tc_destroy(db->tc);
    cudaFree(db->rows);
    cudaFree(db->lens);
    cudaFree(db->gemv_scratch);
    cudaFree(db->ticket);
    cudaFree(db->d_bad);
    cudaFree(db->d_q);
    cudaFree(db->d_keys);
    cudaFree(db->d_scores);
    cudaFree(db->d_ids);
    if (db->h_q) cudaFreeHost(db->h_q);
    if (db->h_scores) cudaFreeHost(db->h_scores);
    if (db->h_ids) cudaFreeHost(db->h_ids);
    if (db->h_keys) cudaFreeHost(db->h_keys);
    for (int i = 0; i < 2; ++i) {
        if (db->h_stage[i]) cudaFreeHost(db->h_stage[i]);
        if (db->stage_ev[i]) cudaEventDestroy(db->stage_ev[i]);
    }
    if (db->ev0) cudaEventDestroy(db->ev0);
    if (db->ev1) cudaEventDestroy(db->ev1);
    if (db->out_ev) cudaEventDestroy(db->out_ev);
    if (db->stream) cudaStreamDestroy(db->stream);
    (void)cudaGetLastError();
    delete db;
    return FCS_OK;
}

extern "C" int fcs_db_get_info(const fcs_db* db, fcs_info* out) {
    if (!db || !out) return FCS_FAIL(FCS_ERR_INVALID, "fcs_db_get_info: NULL argument");
    out->n_rows = db->n_rows;
    out->id_offset = db->id_offset;
    out->device = db->device;
    out->flags = db->flags;
    out->finalized = db->finalized ? 1 : 0;
    out->sm_count = db->sm_count;
    out->bytes_fp32 = uint64_t(db->n_rows) * ROW_BYTES;
    out->bytes_bf16 = db->tc ? tc_image_bytes(db->tc) : 0;
    return FCS_OK;
}

// ------------------------------------------------------------------------------------ upload
static int check_upload(fcs_db* db, int64_t row0, int64_t n, const void* rows, const void* lengths, const char* fn) {
    if (!db) return FCS_FAIL(FCS_ERR_INVALID, "%s: db is NULL", fn);
    if (db->finalized) return FCS_FAIL(FCS_ERR_STATE, "%s: database already finalized", fn);
    if (!rows) return FCS_FAIL(FCS_ERR_INVALID, "%s: rows is NULL", fn);
    if (row0 < 0 || n < 0 || row0 + n > db->n_rows)
        return FCS_FAIL(FCS_ERR_INVALID, "%s: rows [%lld, %lld) outside [0, %lld)", fn, (long long)row0, (long long)(row0 + n),
                    (long long)db->n_rows);
    if ((db->flags & FCS_DB_HAS_LENGTHS) && !lengths) return FCS_FAIL(FCS_ERR_INVALID, "%s: lengths required (FCS_DB_HAS_LENGTHS)", fn);
    return FCS_OK;
}

extern "C" int fcs_db_upload(fcs_db* db, int64_t row0, int64_t n, const float* host_rows, const int32_t* host_lengths) {
    int rc = check_upload(db, row0, n, host_rows, host_lengths, "fcs_db_upload");
    if (rc != FCS_OK) return rc;
    if (n == 0) return FCS_OK;
    DeviceGuard guard(db->device);
    for (int i = 0; i < 2; ++i) {
        if (!db->h_stage[i]) {
            FCS_CUDA(cudaMallocHost(&db->h_stage[i], STAGE_BYTES_HOST));
            FCS_CUDA(cudaEventCreateWithFlags(&db->stage_ev[i], cudaEventDisableTiming));
        }
    }
    // rows: pageable -> pinned (CPU copy) -> device (async DMA), double-buffered
    const int64_t rows_per_stage = int64_t(STAGE_BYTES_HOST / ROW_BYTES);
    for (int64_t r = 0; r < n; r += rows_per_stage) {
        const int64_t cnt = (n - r < rows_per_stage) ? (n - r) : rows_per_stage;
        const int b = db->stage_next;
        db->stage_next ^= 1;
        FCS_CUDA(cudaEventSynchronize(db->stage_ev[b]));
        parallel_memcpy(db->h_stage[b], host_rows + r * DIM, size_t(cnt) * ROW_BYTES);
        FCS_CUDA(cudaMemcpyAsync(db->rows + (row0 + r) * DIM, db->h_stage[b], size_t(cnt) * ROW_BYTES, cudaMemcpyHostToDevice,
                                 db->stream));
        FCS_CUDA(cudaEventRecord(db->stage_ev[b], db->stream));
    }
    if (db->flags & FCS_DB_HAS_LENGTHS) {
        const int64_t per_stage = int64_t(STAGE_BYTES_HOST / sizeof(uint16_t));
        for (int64_t r = 0; r < n; r += per_stage) {
            const int64_t cnt = (n - r < per_stage) ? (n - r) : per_stage;
            const int b = db->stage_next;
            db->stage_next ^= 1;
            FCS_CUDA(cudaEventSynchronize(db->stage_ev[b]));
            uint16_t* dst = static_cast<uint16_t*>(db->h_stage[b]);
            for (int64_t i = 0; i < cnt; ++i) {
                const int32_t v = host_lengths[r + i];
                if (v < 0 || v > 65535)
                    return FCS_FAIL(FCS_ERR_INVALID, "fcs_db_upload: length %d of row %lld outside [0, 65535]", v, (long long)(row0 + r + i));
                dst[i] = uint16_t(v);
            }
            FCS_CUDA(cudaMemcpyAsync(db->lens + row0 + r, dst, size_t(cnt) * sizeof(uint16_t), cudaMemcpyHostToDevice, db->stream));
            FCS_CUDA(cudaEventRecord(db->stage_ev[b], db->stream));
        }
    }
    mark_covered(db, row0, row0 + n);
    return FCS_OK;
}

extern "C" int fcs_db_upload_file(fcs_db* db, int64_t row0, int64_t n, const char* path, int64_t file_offset) {
    int rc = check_upload(db, row0, n, path, nullptr, "fcs_db_upload_file");
    if (rc != FCS_OK) return rc;
    if (db->flags & FCS_DB_HAS_LENGTHS) return FCS_FAIL(FCS_ERR_INVALID, "fcs_db_upload_file: databases with domain lengths are fed with fcs_db_upload");
    if (file_offset < 0) return FCS_FAIL(FCS_ERR_INVALID, "fcs_db_upload_file: negative file offset");
    if (n == 0) return FCS_OK;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return FCS_FAIL(FCS_ERR_INVALID, "fcs_db_upload_file: cannot open %s: %s", path, strerror(errno));
    DeviceGuard guard(db->device);
    auto run = [&]() -> int {
        for (int i = 0; i < 2; ++i) {
            if (!db->h_stage[i]) {
                FCS_CUDA(cudaMallocHost(&db->h_stage[i], STAGE_BYTES_HOST));
                FCS_CUDA(cudaEventCreateWithFlags(&db->stage_ev[i], cudaEventDisableTiming));
            }
        }
        // file -> pinned (positional reads) -> device (async DMA), double-buffered: the reads of block i+1 overlap the DMA of block i
        const int64_t rows_per_stage = int64_t(STAGE_BYTES_HOST / ROW_BYTES);
        for (int64_t r = 0; r < n; r += rows_per_stage) {
            const int64_t cnt = (n - r < rows_per_stage) ? (n - r) : rows_per_stage;
            const int b = db->stage_next;
            db->stage_next ^= 1;
            FCS_CUDA(cudaEventSynchronize(db->stage_ev[b]));
            const int e = parallel_pread(fd, db->h_stage[b], size_t(cnt) * ROW_BYTES, file_offset + r * int64_t(ROW_BYTES));
            if (e) return FCS_FAIL(FCS_ERR_INVALID, "fcs_db_upload_file: reading %s failed: %s", path, e == ENODATA ? "file too short" : strerror(e));
            FCS_CUDA(cudaMemcpyAsync(db->rows + (row0 + r) * DIM, db->h_stage[b], size_t(cnt) * ROW_BYTES, cudaMemcpyHostToDevice, db->stream));
            FCS_CUDA(cudaEventRecord(db->stage_ev[b], db->stream));
        }
        return FCS_OK;
    };
    rc = run();
    close(fd);
    if (rc != FCS_OK) return rc;
    mark_covered(db, row0, row0 + n);
    return FCS_OK;
}

extern "C" int fcs_db_upload_device(fcs_db* db, int64_t row0, int64_t n, const float* dev_rows, const int32_t* dev_lengths) {
    int rc = check_upload(db, row0, n, dev_rows, dev_lengths, "fcs_db_upload_device");
    if (rc != FCS_OK) return rc;
    if (n == 0) return FCS_OK;
    DeviceGuard guard(db->device);
    // the source may have been produced on another stream (e.g. torch's): make it visible first
    FCS_CUDA(cudaDeviceSynchronize());
    FCS_CUDA(cudaMemcpyAsync(db->rows + row0 * DIM, dev_rows, size_t(n) * ROW_BYTES, cudaMemcpyDeviceToDevice, db->stream));
    if (db->flags & FCS_DB_HAS_LENGTHS)
        FCS_CUDA(lengths_to_u16_launch(dev_lengths, db->lens + row0, n, db->d_bad, db->stream));
    FCS_CUDA(cudaStreamSynchronize(db->stream));  // the caller may free dev_rows right after
    mark_covered(db, row0, row0 + n);
    return FCS_OK;
}

extern "C" int fcs_db_finalize(fcs_db* db) {
    if (!db) return FCS_FAIL(FCS_ERR_INVALID, "fcs_db_finalize: db is NULL");
    if (db->finalized) return FCS_OK;
    if (db->covered.size() != 1 || db->covered[0].first != 0 || db->covered[0].second != db->n_rows) {
        int64_t have = 0;
        for (const auto& r : db->covered) have += r.second - r.first;
        return FCS_FAIL(FCS_ERR_STATE, "fcs_db_finalize: only %lld of %lld rows uploaded (%zu disjoint ranges)", (long long)have,
                    (long long)db->n_rows, db->covered.size());
    }
    DeviceGuard guard(db->device);
    FCS_CUDA(finalize_rows_launch(db->rows, db->n_rows, (db->flags & FCS_DB_NORMALISE_ROWS) ? 1 : 0, 1e-8f, db->stream));
    int bad = 0;
    FCS_CUDA(cudaMemcpyAsync(&bad, db->d_bad, sizeof(int), cudaMemcpyDeviceToHost, db->stream));
    FCS_CUDA(cudaStreamSynchronize(db->stream));
    if (bad) return FCS_FAIL(FCS_ERR_INVALID, "fcs_db_finalize: a domain length was outside [0, 65535]");
    for (int i = 0; i < 2; ++i) {  // upload staging is no longer needed
        if (db->h_stage[i]) cudaFreeHost(db->h_stage[i]);
        db->h_stage[i] = nullptr;
    }
    if (db->flags & FCS_DB_KEEP_BF16) {
        int rc = tc_create(&db->tc, db->device, db->sm_count, db->rows, db->n_rows, uint32_t(db->id_offset), db->stream);
        if (rc != FCS_OK) return FCS_FAIL(rc, "fcs_db_finalize: tensor-core path setup failed: %s", tc_last_error());
        tc_set_timing(db->tc, db->profiling);
    }
    db->finalized = true;
    return FCS_OK;
}

// ------------------------------------------------------------------------------------ search
int fcs::api_check_search(const fcs_db* db, const void* q, int nq, int k, int qnorm, int mode, const char* fn) {
    if (!db) return FCS_FAIL(FCS_ERR_INVALID, "%s: db is NULL", fn);
    if (!db->finalized) return FCS_FAIL(FCS_ERR_STATE, "%s: database not finalized", fn);
    if (!q) return FCS_FAIL(FCS_ERR_INVALID, "%s: q is NULL", fn);
    if (nq < 1) return FCS_FAIL(FCS_ERR_INVALID, "%s: nq must be >= 1 (got %d)", fn, nq);
    if (k < 1 || k > FCS_MAX_K) return FCS_FAIL(FCS_ERR_UNSUPPORTED, "%s: k must be in [1, %d] (got %d)", fn, FCS_MAX_K, k);
    if (qnorm < FCS_QNORM_NONE || qnorm > FCS_QNORM_L2) return FCS_FAIL(FCS_ERR_INVALID, "%s: bad qnorm %d", fn, qnorm);
    if (mode < FCS_MODE_AUTO || mode > FCS_MODE_TC) return FCS_FAIL(FCS_ERR_INVALID, "%s: bad mode %d", fn, mode);
    if (mode == FCS_MODE_TC && !db->tc) return FCS_FAIL(FCS_ERR_STATE, "%s: FCS_MODE_TC needs a database created with FCS_DB_KEEP_BF16", fn);
    return FCS_OK;
}

// GEMV path for nq queries (device pointers), all passes enqueued on `stream`.
static int gemv_search(fcs_db* db, const float* q, int nq, const int32_t* qlen, float mincov, int k,
                       int qnorm, float* out_scores, int64_t* out_ids, uint64_t* out_keys, cudaStream_t stream, int* launches) {
    const bool use_mask = qlen != nullptr && db->lens != nullptr;
    for (int q0 = 0; q0 < nq; q0 += GEMV_MAX_NQ) {
        const int nqg = (nq - q0 < GEMV_MAX_NQ) ? (nq - q0) : GEMV_MAX_NQ;
        for (int off = 0; off < k; off += GEMV_MAX_K) {
            GemvParams p = {};
            p.rows = db->rows;
            p.lens = db->lens;
            p.n_rows = db->n_rows;
            p.id_base = uint32_t(db->id_offset);
            p.q = q + size_t(q0) * DIM;
            p.nq = nqg;
            p.qnorm = qnorm;
            p.use_mask = use_mask ? 1 : 0;
            p.mincov = mincov;
            for (int i = 0; i < nqg; ++i) p.qlen[i] = use_mask ? float(qlen[q0 + i]) : 0.f;
            p.k = (k - off < GEMV_MAX_K) ? (k - off) : GEMV_MAX_K;
            p.out_stride = k;
            p.out_off = off;
            p.bounded = off > 0 ? 1 : 0;
            p.scratch = db->gemv_scratch;
            p.ticket = db->ticket;
            p.out_keys = out_keys + size_t(q0) * k;
            p.out_scores = out_scores ? out_scores + size_t(q0) * k : nullptr;
            p.out_ids = out_ids ? out_ids + size_t(q0) * k : nullptr;
            FCS_CUDA(gemv_launch(p, db->sm_count, stream));
            ++*launches;
        }
    }
    return FCS_OK;
}

// Exact-scan passes over the tensor-core path's device-side fallback queue: pass p handles queued queries
// [8p, 8p+8) and writes their rows of the search's output buffers.  The kernels read the queue length on the
// device; a pass beyond the end of the queue exits at once.
static int fallback_passes(fcs_db* db, const fcs_db::Pending& pd, int pass0, int pass1, cudaStream_t stream, int* launches) {
    for (int pass = pass0; pass < pass1; ++pass) {
        for (int off = 0; off < pd.k; off += GEMV_MAX_K) {
            GemvParams p = {};
            p.rows = db->rows;
            p.lens = nullptr;
            p.n_rows = db->n_rows;
            p.id_base = uint32_t(db->id_offset);
            p.q = pd.q.q_dev + size_t(pass) * GEMV_MAX_NQ * DIM;
            p.nq = GEMV_MAX_NQ;
            p.qnorm = pd.qnorm;
            p.use_mask = 0;
            p.k = (pd.k - off < GEMV_MAX_K) ? (pd.k - off) : GEMV_MAX_K;
            p.out_stride = pd.k;
            p.out_off = off;
            p.bounded = off > 0 ? 1 : 0;
            p.scratch = db->gemv_scratch;
            p.ticket = db->ticket;
            p.out_keys = pd.out_keys;
            p.out_scores = pd.out_scores;
            p.out_ids = pd.out_ids;
            p.nq_dev = pd.q.count_dev;
            p.nq_off = pass * GEMV_MAX_NQ;
            p.out_index = pd.q.list_dev + size_t(pass) * GEMV_MAX_NQ;
            FCS_CUDA(gemv_launch(p, db->sm_count, stream));
            ++*launches;
        }
    }
    return FCS_OK;
}

// The stream's work is complete: if the last tensor-core search queued more queries than the passes enqueued behind it
// cover, scan the rest now (same stream) and wait.  Returns the queue length through *n_queued.
int fcs::api_finish_pending(fcs_db* db, cudaStream_t stream, int* n_queued) {
    *n_queued = 0;
    if (!db->pending.valid) return FCS_OK;
    const int queued = int(*db->pending.q.count_host);
    *n_queued = queued;
    db->timing.last_tc_fallbacks = queued;
    const int passes = (queued + GEMV_MAX_NQ - 1) / GEMV_MAX_NQ;
    if (passes > FB_ASYNC_PASSES) {
        int launches = 0;
        int rc = fallback_passes(db, db->pending, FB_ASYNC_PASSES, passes, stream, &launches);
        if (rc != FCS_OK) return rc;
        db->timing.last_launches += launches;
        FCS_CUDA(cudaStreamSynchronize(stream));
    }
    db->pending.valid = false;
    return FCS_OK;
}

// AUTO picks the cheaper of two fitted cost models (profiles/r02_auto_model.json: 5 shard sizes x 10 batch sizes on B200,
// no wrong choice on that grid):
//   exact scan    ceil(nq/8) * ceil(k/128) passes, each 50 us + 82 ns per 1000 rows (8 queries per pass: issue-bound FFMA2);
//   tensor cores  125 us of sampling rounds / selections / rescore + 100 ns per 1000 rows per 512-query group.
// One scan pass (nq <= 8) always wins; with the coverage mask on only the scan applies.
bool fcs::api_auto_prefers_tc(const fcs_db* db, int nq, int k, bool mask_on) {
    if (!db->tc || mask_on || k > tc_max_k() || nq < tc_min_batch()) return false;
    const double rows = double(db->n_rows);
    const double passes = double((nq + GEMV_MAX_NQ - 1) / GEMV_MAX_NQ) * double((k + GEMV_MAX_K - 1) / GEMV_MAX_K);
    const double t_gemv = passes * (5.0e-5 + rows * 8.2e-11);
    const double t_tc = 1.25e-4 + double((nq + 511) / 512) * rows * 1.0e-10;
    return t_tc < t_gemv;
}

int fcs::api_search_core(fcs_db* db, const float* q_dev, int nq, const int32_t* qlen, float mincov, int k,
                       int qnorm, int mode, int kprime, float* out_scores, int64_t* out_ids, uint64_t* out_keys,
                       cudaStream_t stream) {
    int use_mode = mode;
    const bool mask_on = qlen != nullptr && db->lens != nullptr;
    if (use_mode == FCS_MODE_AUTO) use_mode = api_auto_prefers_tc(db, nq, k, mask_on) ? FCS_MODE_TC : FCS_MODE_GEMV;
    if (use_mode == FCS_MODE_TC && mask_on)
        return FCS_FAIL(FCS_ERR_UNSUPPORTED, "FCS_MODE_TC does not apply the coverage mask (the faiss flavour has none, dbsearch.py:307-310)");
    int launches = 0;
    if (db->profiling) FCS_CUDA(cudaEventRecord(db->ev0, stream));
    int rc = FCS_OK;
    db->pending.valid = false;
    if (use_mode == FCS_MODE_GEMV) {
        rc = gemv_search(db, q_dev, nq, qlen, mincov, k, qnorm, out_scores, out_ids, out_keys, stream, &launches);
    } else {
        fcs_db::Pending pd;
        rc = tc_search(db->tc, q_dev, nq, k, kprime, qnorm, out_scores, out_ids, out_keys, stream, &launches, &pd.q);
        if (rc != FCS_OK) return FCS_FAIL(rc, "tensor-core search failed: %s", tc_last_error());
        // queries whose exactness certificate failed (or whose candidate buffer overflowed) sit in a device-side queue:
        // the exact scan runs over it, 8 queries per pass, with the queue length read by the kernels themselves
        pd.valid = true;
        pd.k = k;
        pd.qnorm = qnorm;
        pd.out_scores = out_scores;
        pd.out_ids = out_ids;
        pd.out_keys = out_keys;
        const int max_passes = (nq + GEMV_MAX_NQ - 1) / GEMV_MAX_NQ;
        tc_phase_mark(db->tc, "exact-scan queue", stream);
        rc = fallback_passes(db, pd, 0, max_passes < FB_ASYNC_PASSES ? max_passes : FB_ASYNC_PASSES, stream, &launches);
        tc_phase_mark(db->tc, "end", stream);
        db->pending = pd;
    }
    if (rc != FCS_OK) return rc;
    if (db->profiling) FCS_CUDA(cudaEventRecord(db->ev1, stream));
    db->ev_valid = db->profiling;
    db->timing.last_mode = use_mode;
    db->timing.last_launches = launches;
    db->timing.last_tc_fallbacks = 0;
    return FCS_OK;
}

extern "C" int fcs_search_device(fcs_db* db, const float* q_dev, int nq, const int32_t* qlen, float mincov, int k, int qnorm,
                                 int mode, int kprime, float* out_scores_dev, int64_t* out_ids_dev, uint64_t* out_keys_dev,
                                 void* stream) {
    int rc = api_check_search(db, q_dev, nq, k, qnorm, mode, "fcs_search_device");
    if (rc != FCS_OK) return rc;
    DeviceGuard guard(db->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : db->stream;
    uint64_t* keys = out_keys_dev;
    if (!keys) {
        rc = api_ensure_out_bufs(db, size_t(nq) * k, false);
        if (rc != FCS_OK) return rc;
        keys = db->d_keys;
    }
    return api_search_core(db, q_dev, nq, qlen, mincov, k, qnorm, mode, kprime, out_scores_dev, out_ids_dev, keys, st);
}

extern "C" int fcs_search_queue_len_to(fcs_db* db, void* dst_dev_u32, void* stream) {
    if (!db || !dst_dev_u32) return FCS_FAIL(FCS_ERR_INVALID, "fcs_search_queue_len_to: NULL argument");
    DeviceGuard guard(db->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : db->stream;
    if (db->pending.valid)
        FCS_CUDA(cudaMemcpyAsync(dst_dev_u32, db->pending.q.count_dev, sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
    else
        FCS_CUDA(cudaMemsetAsync(dst_dev_u32, 0, sizeof(unsigned), st));
    return FCS_OK;
}

extern "C" int fcs_search_finish(fcs_db* db, void* stream, int* out_queued) {
    if (!db) return FCS_FAIL(FCS_ERR_INVALID, "fcs_search_finish: db is NULL");
    DeviceGuard guard(db->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : db->stream;
    FCS_CUDA(cudaStreamSynchronize(st));
    int queued = 0;
    int rc = api_finish_pending(db, st, &queued);
    if (out_queued) *out_queued = queued;
    return rc;
}

extern "C" int fcs_search(fcs_db* db, const float* q, int nq, const int32_t* qlen, float mincov, int k, int qnorm, int mode,
                          int kprime, float* out_scores, int64_t* out_ids) {
    int rc = api_check_search(db, q, nq, k, qnorm, mode, "fcs_search");
    if (rc != FCS_OK) return rc;
    if (!out_scores || !out_ids) return FCS_FAIL(FCS_ERR_INVALID, "fcs_search: output buffer is NULL");
    DeviceGuard guard(db->device);
    const size_t entries = size_t(nq) * k;
    if ((rc = api_ensure_query_bufs(db, size_t(nq), true)) != FCS_OK) return rc;
    if ((rc = api_ensure_out_bufs(db, entries, true)) != FCS_OK) return rc;
    memcpy(db->h_q, q, size_t(nq) * DIM * sizeof(float));
    // Small exact-scan searches (the per-query loop of the reference's torch flavour) run zero-copy: the kernel
    // reads the queries from, and writes the k results to, pinned host memory (UVA), so the call is one launch
    // plus one stream synchronisation instead of launch + three copies.
    const bool gemv = mode == FCS_MODE_GEMV || (mode == FCS_MODE_AUTO && !api_auto_prefers_tc(db, nq, k, qlen != nullptr && db->lens != nullptr));
    if (gemv && k <= GEMV_MAX_K && nq <= 64) {
        rc = api_search_core(db, db->h_q, nq, qlen, mincov, k, qnorm, FCS_MODE_GEMV, kprime, db->h_scores, db->h_ids, db->h_keys, db->stream);
        if (rc != FCS_OK) return rc;
        FCS_CUDA(cudaStreamSynchronize(db->stream));
    } else {
        FCS_CUDA(cudaMemcpyAsync(db->d_q, db->h_q, size_t(nq) * DIM * sizeof(float), cudaMemcpyHostToDevice, db->stream));
        rc = api_search_core(db, db->d_q, nq, qlen, mincov, k, qnorm, mode, kprime, db->d_scores, db->d_ids, db->d_keys, db->stream);
        if (rc != FCS_OK) return rc;
        for (int attempt = 0; attempt < 2; ++attempt) {
            FCS_CUDA(cudaMemcpyAsync(db->h_scores, db->d_scores, entries * sizeof(float), cudaMemcpyDeviceToHost, db->stream));
            FCS_CUDA(cudaEventRecord(db->out_ev, db->stream));
            FCS_CUDA(cudaMemcpyAsync(db->h_ids, db->d_ids, entries * sizeof(int64_t), cudaMemcpyDeviceToHost, db->stream));
            // the scores go to the caller's buffer while the (twice as large) ids are still crossing PCIe
            FCS_CUDA(cudaEventSynchronize(db->out_ev));
            memcpy(out_scores, db->h_scores, entries * sizeof(float));
            FCS_CUDA(cudaStreamSynchronize(db->stream));
            // a fallback queue longer than the passes enqueued behind the search: scan the rest, copy again
            int queued = 0;
            const bool had_pending = db->pending.valid;
            if ((rc = api_finish_pending(db, db->stream, &queued)) != FCS_OK) return rc;
            if (!had_pending || queued <= FB_ASYNC_PASSES * GEMV_MAX_NQ) break;
        }
        memcpy(out_ids, db->h_ids, entries * sizeof(int64_t));
        return FCS_OK;
    }
    memcpy(out_scores, db->h_scores, entries * sizeof(float));
    memcpy(out_ids, db->h_ids, entries * sizeof(int64_t));
    return FCS_OK;
}

extern "C" int fcs_debug_tc_approx(fcs_db* db, const float* q, int nq, int qnorm, float* out_scores) {
    if (!db || !q || !out_scores || nq < 1) return FCS_FAIL(FCS_ERR_INVALID, "fcs_debug_tc_approx: bad argument");
    if (!db->finalized || !db->tc) return FCS_FAIL(FCS_ERR_STATE, "fcs_debug_tc_approx: needs a finalized database with FCS_DB_KEEP_BF16");
    DeviceGuard guard(db->device);
    int rc = api_ensure_query_bufs(db, size_t(nq), true);
    if (rc != FCS_OK) return rc;
    memcpy(db->h_q, q, size_t(nq) * DIM * sizeof(float));
    FCS_CUDA(cudaMemcpyAsync(db->d_q, db->h_q, size_t(nq) * DIM * sizeof(float), cudaMemcpyHostToDevice, db->stream));
    rc = tc_debug_approx(db->tc, db->d_q, nq, qnorm, out_scores, db->stream);
    if (rc != FCS_OK) return FCS_FAIL(rc, "fcs_debug_tc_approx: %s", tc_last_error());
    return FCS_OK;
}

extern "C" int fcs_debug_tc_plan(int64_t n_rows, int kprime, int nq, int64_t* out_rounds, int max_rounds) {
    if (n_rows < 1 || nq < 1 || !out_rounds || max_rounds < 1) return FCS_FAIL(FCS_ERR_INVALID, "fcs_debug_tc_plan: bad argument");
    return tc_debug_plan(n_rows, kprime > 0 ? kprime : tc_default_kprime(10), (nq + 511) / 512, out_rounds, max_rounds);
}

extern "C" int64_t fcs_debug_tc_tile_of(int64_t j0, int64_t stride, int64_t comp_t, int64_t idx) {
    return tc_debug_tile_of(j0, stride, comp_t, idx);
}

extern "C" int fcs_merge_topk(int device, const uint64_t* keys_dev, int n_lists, int nq, int k, float* out_scores_dev,
                              int64_t* out_ids_dev, void* stream) {
    if (!keys_dev) return FCS_FAIL(FCS_ERR_INVALID, "fcs_merge_topk: keys_dev is NULL");
    if (n_lists < 1 || nq < 1 || k < 1) return FCS_FAIL(FCS_ERR_INVALID, "fcs_merge_topk: n_lists, nq and k must be >= 1");
    DeviceGuard guard(device);
    if (!guard.ok) return FCS_FAIL(FCS_ERR_CUDA, "fcs_merge_topk: cudaSetDevice(%d) failed", device);
    FCS_CUDA(merge_topk_launch(keys_dev, n_lists, nq, k, out_scores_dev, out_ids_dev, nullptr, static_cast<cudaStream_t>(stream)));
    return FCS_OK;
}

extern "C" int fcs_set_profiling(fcs_db* db, int enable) {
    if (!db) return FCS_FAIL(FCS_ERR_INVALID, "fcs_set_profiling: db is NULL");
    db->profiling = enable != 0;
    if (!db->profiling) db->ev_valid = false;
    if (db->tc) tc_set_timing(db->tc, db->profiling);
    return FCS_OK;
}

extern "C" int fcs_get_timing(const fcs_db* db_c, fcs_timing* out) {
    if (!db_c || !out) return FCS_FAIL(FCS_ERR_INVALID, "fcs_get_timing: NULL argument");
    fcs_db* db = const_cast<fcs_db*>(db_c);
    DeviceGuard guard(db->device);
    if (db->ev_valid) {
        FCS_CUDA(cudaEventSynchronize(db->ev1));
        float ms = 0.f;
        FCS_CUDA(cudaEventElapsedTime(&ms, db->ev0, db->ev1));
        db->timing.last_search_ms = ms;
        db->timing.last_kernel_ms = ms;
    } else {
        db->timing.last_search_ms = 0.f;
        db->timing.last_kernel_ms = 0.f;
    }
    db->timing.last_rounds = 1;
    if (db->timing.last_mode == FCS_MODE_TC && db->tc) {  // the TC path times its GEMM launches itself
        db->timing.last_kernel_ms = tc_last_kernel_ms(db->tc);  // waits for the search's device work
        db->timing.last_rounds = tc_last_rounds(db->tc);
        db->timing.last_tc_fallbacks = tc_last_flagged(db->tc);
    }
    *out = db->timing;
    return FCS_OK;
}
