// fcs_tc.cu -- K3/K4 placeholder: the tcgen05 batched path is not wired in yet.
#include <string>

#include "fcs_internal.h"
#include "fcs_tc.h"

namespace fcs {

struct TcState {
    int device;
};
static thread_local std::string g_tc_error;

int tc_create(TcState** out, int device, int, const float*, const void*, int64_t, uint32_t) {
    *out = new TcState{device};
    return FCS_OK;
}
void tc_destroy(TcState* s) { delete s; }
int tc_search(TcState*, const float*, int, int, int, int, float*, int64_t*, uint64_t*, cudaStream_t, int*, int*) {
    g_tc_error = "tensor-core path not built into this library yet";
    return FCS_ERR_UNSUPPORTED;
}
const char* tc_last_error() { return g_tc_error.c_str(); }
int tc_min_batch() { return 1 << 30; }
int tc_max_k() { return 0; }

}  // namespace fcs
