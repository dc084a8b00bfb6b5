// fcs_tc.h -- host-side interface of the tensor-core (tcgen05) batched path (fcs_tc.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fcs {

struct TcState;

int tc_create(TcState** out, int device, int sm_count, const float* rows, int64_t n_rows, uint32_t id_base,
              cudaStream_t stream);
uint64_t tc_image_bytes(const TcState* s);
float tc_last_kernel_ms(const TcState* s);  // sum over the GEMM+filter launches of the last search
int tc_last_rounds(const TcState* s);
void tc_destroy(TcState* s);
// Enqueues the whole batched search on `stream`; outputs are device pointers.  Returns an FCS_* code.
// Synchronises `stream` at the end (the fallback decision is taken on the host).  If *n_flagged > 0,
// (*flagged_host)[q] & 3 != 0 marks the queries the caller must re-run on the exact scan.
int tc_search(TcState* s, const float* q_dev, int nq, int k, int kprime, int qnorm, float* out_scores, int64_t* out_ids,
              uint64_t* out_keys, cudaStream_t stream, int* launches, const unsigned** flagged_host, int* n_flagged);
int tc_default_kprime(int k);
int tc_debug_approx(TcState* s, const float* q_dev, int nq, int qnorm, float* out_host, cudaStream_t stream);
const char* tc_last_error();
int tc_min_batch();  // AUTO mode switches to the TC path at this many queries
int tc_max_k();

}  // namespace fcs
