// fcs_internal.h -- host-side declarations shared by the .cu files of libfcsearch.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fcsearch.h"

namespace fcs {

// thread-local message behind fcs_last_error() (fcs_api.cu); used by the other translation units of the C ABI
void set_last_error(const char* msg);

// ------------------------------------------------------------------ GEMV path (fcs_gemv.cu)
constexpr int GEMV_MAX_NQ = 8;     // queries scored per DB pass by one launch
constexpr int GEMV_MAX_K = 128;    // k per launch (register-resident lists); larger k = several passes
constexpr int GEMV_WARPS = 11;     // consumer warps == ring stages
constexpr int GEMV_STAGE_ROWS = 32;

struct GemvParams {
    const float* rows;      // [n_rows,128] fp32, row-swizzled (fcs_common.cuh), normalised when the flavour asks for it
    const uint16_t* lens;   // [n_rows] domain lengths or nullptr
    int64_t n_rows;
    uint32_t id_base;       // global id of row 0
    const float* q;         // [nq,128] raw queries (device)
    int nq;                 // 1..GEMV_MAX_NQ
    int qnorm;              // FCS_QNORM_*
    int use_mask;
    float mincov;
    float qlen[GEMV_MAX_NQ];
    int k;                  // 1..GEMV_MAX_K for this pass
    int out_stride;         // row pitch of the output arrays (total k of the call)
    int out_off;            // first output rank written by this pass
    int bounded;            // 1: only keys < out_keys[q*out_stride + out_off - 1] may enter (pass >= 2)
    uint64_t* scratch;      // [nq][grid][k]
    unsigned* ticket;       // last-block-done counter (self-resetting)
    uint64_t* out_keys;     // [nq][out_stride]
    float* out_scores;      // [nq][out_stride] or nullptr
    int64_t* out_ids;       // [nq][out_stride] or nullptr
    // Indirect launch (the tensor-core path's device-side fallback queue): the number of queries is read on the
    // device, nq = clamp(*nq_dev - nq_off, 0, GEMV_MAX_NQ) -- a launch with nothing to do exits at once -- and
    // query i writes output row out_index[i] instead of row i.
    const unsigned* nq_dev; // nullptr = direct launch
    int nq_off;
    const int* out_index;   // [nq] output rows or nullptr
};

size_t gemv_scratch_bytes(int max_grid);
// Launches the streaming kernel; grid <= sm_count.  Returns a cudaError_t.
cudaError_t gemv_launch(const GemvParams& p, int sm_count, cudaStream_t stream);
cudaError_t gemv_configure();  // one-time cudaFuncSetAttribute for every instantiation

// ------------------------------------------------------------------ loader kernels (fcs_loader.cu)
// in place: (optionally) divide each row by max(|row|, eps), then store it chunk-swizzled (fcs_common.cuh)
cudaError_t finalize_rows_launch(float* rows, int64_t n_rows, int normalise, float eps, cudaStream_t stream);
// int32 -> u16 with range check; *bad_flag (device int) is set to 1 if any length is outside [0, 65535]
cudaError_t lengths_to_u16_launch(const int32_t* lens, uint16_t* out, int64_t n, int* bad_flag, cudaStream_t stream);

// ------------------------------------------------------------------ merge kernel (fcs_merge.cu)
cudaError_t merge_topk_launch(const uint64_t* keys, int n_lists, int nq, int k, float* out_scores, int64_t* out_ids,
                              uint64_t* out_keys, cudaStream_t stream);

}  // namespace fcs
