// fcs_tc.h -- host-side interface of the tensor-core (tcgen05) batched path (fcs_tc.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fcs {

struct TcState;

int tc_create(TcState** out, int device, int sm_count, const float* rows, const void* rows_bf16, int64_t n_rows,
              uint32_t id_base);
void tc_destroy(TcState* s);
// Enqueues the whole batched search on `stream`; outputs are device pointers.  Returns an FCS_* code.
int tc_search(TcState* s, const float* q_dev, int nq, int k, int kprime, int qnorm, float* out_scores, int64_t* out_ids,
              uint64_t* out_keys, cudaStream_t stream, int* launches, int* fallbacks);
const char* tc_last_error();
int tc_min_batch();  // AUTO mode switches to the TC path at this many queries
int tc_max_k();

}  // namespace fcs
