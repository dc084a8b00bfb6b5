// fcs_tc.h -- host-side interface of the tensor-core (tcgen05) batched path (fcs_tc.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fcs {

struct TcState;

// Queries of the last tc_search whose exactness certificate failed, queued on the device for the exact scan.
struct TcFallbackQueue {
    const unsigned* count_dev = nullptr;  // queue length (device)
    const int* list_dev = nullptr;        // [count] query indices (device)
    const float* q_dev = nullptr;         // [count][128] their raw queries (device)
    const unsigned* count_host = nullptr; // pinned copy of the length, valid once the search's stream work is complete
};

int tc_create(TcState** out, int device, int sm_count, const float* rows, int64_t n_rows, uint32_t id_base,
              cudaStream_t stream);
uint64_t tc_image_bytes(const TcState* s);
// The next three wait for the last search's device work (an event), then read what it recorded.
void tc_set_timing(TcState* s, bool on);  // event pairs around the GEMM+filter launches of every search (off by default)
float tc_last_kernel_ms(TcState* s);  // sum over the GEMM+filter launches of the last search (0 unless timing was on)
int tc_last_flagged(TcState* s);      // length of the last search's fallback queue
int tc_last_rounds(const TcState* s);
void tc_destroy(TcState* s);
// Enqueues the whole batched search on `stream`; outputs are device pointers.  Returns an FCS_* code without
// synchronising: the caller enqueues the exact scan of the fallback queue behind it (fcs_api.cu).
int tc_search(TcState* s, const float* q_dev, int nq, int k, int kprime, int qnorm, float* out_scores, int64_t* out_ids,
              uint64_t* out_keys, cudaStream_t stream, int* launches, TcFallbackQueue* fbq);
int tc_default_kprime(int k);
int tc_debug_approx(TcState* s, const float* q_dev, int nq, int qnorm, float* out_host, cudaStream_t stream);
int tc_debug_plan(int64_t n_rows, int kprime, int n_qgroups, int64_t* out, int max_rounds);
int64_t tc_debug_tile_of(int64_t j0, int64_t stride, int64_t comp_T, int64_t idx);
void tc_phase_mark(TcState* s, const char* label, cudaStream_t stream);  // diagnostics (FCS_TC_PHASES=1)
const char* tc_last_error();
int tc_min_batch();  // AUTO mode switches to the TC path at this many queries
int tc_max_k();

}  // namespace fcs
