/* fcsearch.h -- C ABI of the B200-native Foldclass database search.
 *
 * Drop-in boundary for the search hot path of psipred/merizo_search.  The
 * reference has no FFI of its own (it is pure Python); the three Python callables
 * this library sits behind are (paths relative to merizo_search/programs/Foldclass/):
 *
 *   read_database()            dbsearch.py:48-72   -> fcs_db_create / fcs_db_upload* / fcs_db_finalize
 *   search_query_against_db()  dbsearch.py:75-81   -> fcs_search (qnorm = FCS_QNORM_COSINE, lengths + mincov)
 *   knn_exact_faiss()          dbsearch.py:213-248 -> fcs_search (qnorm = FCS_QNORM_NONE / FCS_QNORM_L2, no mask)
 *   F.normalize(queries)       dbsearch.py:303-304 -> qnorm = FCS_QNORM_L2 (fused into the search kernels)
 *   db_memmap()/db_iterator()  dbutil.py:28-35     -> the host feeds fcs_db_upload block by block, ONCE
 *
 * Plain pointers and sizes only; no torch types.  Every function returns FCS_OK (0)
 * or a negative error code and never throws or exits; the message of the last error
 * on the calling thread is available from fcs_last_error().
 *
 * A handle owns ONE row shard on ONE device (global ids = id_offset + local row).
 * Multi-GPU = one handle per GPU (one per rank under torchrun, or several handles in
 * one process) plus fcs_merge_topk over the gathered per-shard key lists.
 */
#ifndef FCSEARCH_H
#define FCSEARCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCS_DIM 128 /* embedding width: FoldClassNet(128), dbsearch.py:39 */

/* error codes */
#define FCS_OK 0
#define FCS_ERR_INVALID (-1)     /* bad argument */
#define FCS_ERR_CUDA (-2)        /* CUDA runtime/driver error (message has the detail) */
#define FCS_ERR_STATE (-3)       /* call sequence error (e.g. search before finalize) */
#define FCS_ERR_UNSUPPORTED (-4) /* k / mode outside what the kernels implement */
#define FCS_ERR_NOMEM (-5)       /* device or pinned host allocation failed */

/* fcs_db_create flags */
#define FCS_DB_NORMALISE_ROWS 1u /* .pt flavour: rows are raw; divide each by max(|row|,1e-8)
                                    (F.cosine_similarity semantics, dbsearch.py:78)            */
#define FCS_DB_KEEP_BF16 2u      /* keep a bf16 copy of the (normalised) rows for the tcgen05 path */
#define FCS_DB_HAS_LENGTHS 4u    /* per-row domain lengths are uploaded (coverage mask, dbsearch.py:76) */

/* query normalisation, fused into the search kernels */
#define FCS_QNORM_NONE 0   /* use queries as given (already normalised) */
#define FCS_QNORM_COSINE 1 /* q / max(|q|, 1e-8)   -- F.cosine_similarity, dbsearch.py:78 */
#define FCS_QNORM_L2 2     /* q / max(|q|, 1e-12)  -- F.normalize,          dbsearch.py:304 */

/* search mode */
#define FCS_MODE_AUTO 0 /* GEMV for small batches, tensor-core path for large ones when a bf16 copy exists */
#define FCS_MODE_GEMV 1 /* exact fp32 streaming kernel (HBM-bound) */
#define FCS_MODE_TC 2   /* bf16 tcgen05 GEMM + fused top-k' + exact fp32 rescore (needs FCS_DB_KEEP_BF16) */

#define FCS_MAX_K 2048 /* like faiss-gpu's k-selection limit */

typedef struct fcs_db fcs_db;

typedef struct fcs_timing {
    float last_search_ms;     /* device time of the last fcs_search* call (CUDA events on the handle's stream) */
    float last_kernel_ms;     /* device time of the dominant kernel(s) of that call */
    int32_t last_mode;        /* FCS_MODE_GEMV or FCS_MODE_TC actually used */
    int32_t last_launches;    /* kernels launched by that call */
    int32_t last_tc_fallbacks; /* TC path: queries whose exactness certificate failed and were re-run on the GEMV path */
    int32_t last_rounds;      /* TC path: GEMM+filter launches (rounds) of that call; 1 otherwise */
} fcs_timing;

typedef struct fcs_info {
    int64_t n_rows;
    int64_t id_offset;
    int32_t device;
    uint32_t flags;
    int32_t finalized;
    int32_t sm_count;
    uint64_t bytes_fp32;
    uint64_t bytes_bf16;
} fcs_info;

/* library ------------------------------------------------------------------------- */
int fcs_version(void);                 /* MAJOR*10000 + MINOR*100 + PATCH */
const char* fcs_last_error(void);      /* thread-local; "" if none */
int fcs_device_count(int* out_count);  /* FCS_ERR_CUDA if there is no usable GPU */

/* database handle (replaces read_database, dbsearch.py:48-72) -------------------- */
int fcs_db_create(int device, int64_t n_rows, int dim /* must be 128 */, int64_t id_offset,
                  uint32_t flags, fcs_db** out);
/* Copy rows [row0, row0+n) from HOST memory (pageable is fine: staged through pinned
 * buffers).  lengths (int32, one per row) is required iff FCS_DB_HAS_LENGTHS. */
int fcs_db_upload(fcs_db* db, int64_t row0, int64_t n, const float* host_rows, const int32_t* host_lengths);
/* Same, straight from a FILE of headerless fp32 rows (the faiss flavour's `*_raw_128d_norm.db`,
 * dbutil.py:28-30): rows [row0, row0+n) are read from byte `file_offset` on with positional reads
 * into the pinned staging buffers -- no intermediate copy, no page faults on a mapping.  Not for
 * databases with domain lengths. */
int fcs_db_upload_file(fcs_db* db, int64_t row0, int64_t n, const char* path, int64_t file_offset);
/* Same, from DEVICE memory on the handle's device (synthetic DBs, torch CUDA tensors). */
int fcs_db_upload_device(fcs_db* db, int64_t row0, int64_t n, const float* dev_rows, const int32_t* dev_lengths);
/* Normalise rows (if asked), build the bf16 copy (if asked), allocate search scratch. */
int fcs_db_finalize(fcs_db* db);
int fcs_db_get_info(const fcs_db* db, fcs_info* out);
int fcs_db_destroy(fcs_db* db);

/* search (replaces search_query_against_db dbsearch.py:75-81 and knn_exact_faiss
 * dbsearch.py:213-248) -------------------------------------------------------------
 *   q          [nq,128] fp32 row-major, HOST memory
 *   qlen       per-query residue count (len(query_dict['seq'])) or NULL = no coverage mask
 *   mincov     coverage threshold (only used when qlen != NULL and the db has lengths)
 *   k          1..FCS_MAX_K.  Slots beyond the number of rows hold (-inf, -1) (faiss convention)
 *   kprime     TC path: candidates kept per query before the fp32 rescore; 0 = default
 *   out_scores [nq,k] fp32, out_ids [nq,k] int64 (global ids), HOST memory, sorted by
 *              descending score (ties: ascending id).  Fully written on return. */
int fcs_search(fcs_db* db, const float* q, int nq, const int32_t* qlen, float mincov, int k, int qnorm,
               int mode, int kprime, float* out_scores, int64_t* out_ids);
/* Same with DEVICE buffers, asynchronous on `stream` (a cudaStream_t; NULL = the
 * handle's own stream): the call only enqueues work and never synchronises, so it can sit
 * between an upstream kernel and a collective.  qlen stays a HOST array.  out_keys
 * ([nq,k] uint64, optional) receives the packed sort keys that fcs_merge_topk consumes.
 * Tensor-core searches queue the queries whose exactness certificate failed on the device and
 * re-run up to FCS_ASYNC_FALLBACK_QUERIES of them on the exact scan inside the same
 * enqueued work; a longer queue (pathological data: thousands of near-identical rows) is
 * completed by fcs_search_finish. */
#define FCS_ASYNC_FALLBACK_QUERIES 32
int fcs_search_device(fcs_db* db, const float* q_dev, int nq, const int32_t* qlen, float mincov, int k,
                      int qnorm, int mode, int kprime, float* out_scores_dev, int64_t* out_ids_dev,
                      uint64_t* out_keys_dev, void* stream);
/* Synchronises `stream` (NULL = the handle's own) and, if the last fcs_search_device queued more
 * than FCS_ASYNC_FALLBACK_QUERIES queries for the exact scan, scans the rest into the same output
 * buffers (which must still be valid) and waits.  *out_queued (optional) = length of that queue.
 * `stream` must be the stream the search was enqueued on.  Cheap when there is nothing to do; fcs_search
 * does this itself. */
int fcs_search_finish(fcs_db* db, void* stream, int* out_queued);
/* Enqueues on `stream` a copy of the last fcs_search_device's fallback-queue length (uint32) into DEVICE memory at
 * dst_dev_u32, so that a multi-rank host can ship it with the key lists (one collective) and learn after its one
 * synchronisation whether any rank has to call fcs_search_finish.  0 for searches without a queue. */
int fcs_search_queue_len_to(fcs_db* db, void* dst_dev_u32, void* stream);

/* cross-shard merge (replaces faiss.ResultHeap.add_result/finalize, dbsearch.py:224-245):
 *   keys_dev [n_lists][nq][k] packed keys from fcs_search_device of each shard (after the
 *   all-gather), -> the k best per query, decoded.  Asynchronous on `stream`. */
int fcs_merge_topk(int device, const uint64_t* keys_dev, int n_lists, int nq, int k, float* out_scores_dev,
                   int64_t* out_ids_dev, void* stream);

/* shard group: several row shards behind one handle, one host thread (replaces
 * faiss.index_cpu_to_all_gpus + ResultHeap, dbsearch.py:228-245) -----------------------------------
 * `devices[s]` is the GPU of shard s (an ordinal may repeat: several shards on one GPU).  Shard s
 * holds rows [s*ceil(N/G), min(N,(s+1)*ceil(N/G))); ids are global.  upload / upload_file /
 * finalize feed every shard from its own host thread; search replicates the queries, searches
 * all shards concurrently, moves the per-shard key lists device-to-device (NVLink peer copies)
 * to devices[0] and merges them there.  Same argument meaning and result contract as fcs_search. */
typedef struct fcs_group fcs_group;
int fcs_group_create(const int* devices, int n_shards, int64_t n_rows, int dim, uint32_t flags, fcs_group** out);
int fcs_group_upload(fcs_group* g, int64_t row0, int64_t n, const float* host_rows, const int32_t* host_lengths);
int fcs_group_upload_file(fcs_group* g, const char* path, int64_t file_offset, int64_t row0, int64_t n);
int fcs_group_finalize(fcs_group* g);
int fcs_group_search(fcs_group* g, const float* q, int nq, const int32_t* qlen, float mincov, int k, int qnorm,
                     int mode, int kprime, float* out_scores, int64_t* out_ids);
/* out_bounds: n_shards+1 row offsets; out_devices: n_shards ordinals (either may be NULL) */
int fcs_group_get_info(const fcs_group* g, int* out_n_shards, int64_t* out_bounds, int* out_devices, int max_shards);
int fcs_group_shard(fcs_group* g, int index, fcs_db** out); /* borrowed handle (timing, info) */
int fcs_group_last_fallbacks(const fcs_group* g);           /* exact-scan fallbacks of the last search, all shards */
int fcs_group_destroy(fcs_group* g);

/* last_search_ms and the tensor-core path's last_kernel_ms (event pairs around its GEMM+filter launches) need
 * fcs_set_profiling(db, 1); off by default because an event between two kernels prevents their
 * programmatic-dependent-launch overlap (scan after scan; GEMM round after selection). */
int fcs_set_profiling(fcs_db* db, int enable);
int fcs_get_timing(const fcs_db* db, fcs_timing* out);

/* Test hook (not part of the drop-in surface): the approximate bf16 tensor-core score of every
 * (query,row) pair, out_scores [nq, n_rows] HOST fp32; only for shards of <= 4096 rows. */
int fcs_debug_tc_approx(fcs_db* db, const float* q, int nq, int qnorm, float* out_scores);
/* Test hooks: the round plan of the tensor-core path for a shard of n_rows rows and a batch of nq queries (7 int64 per round: tiles, first sample
 * index, sample stride, complement size, round-0 flag, selection rank, partition flag; returns the number of rounds) and
 * the database tile a round visits at position idx. */
int fcs_debug_tc_plan(int64_t n_rows, int kprime, int nq, int64_t* out_rounds, int max_rounds);
int64_t fcs_debug_tc_tile_of(int64_t j0, int64_t stride, int64_t comp_t, int64_t idx);

#ifdef __cplusplus
}
#endif
#endif /* FCSEARCH_H */
