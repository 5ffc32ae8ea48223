// Small non-kernel entry points of the C ABI.
#include "b2f_common.cuh"

namespace b2f {
char* last_error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}
int& last_flow_kernel() {
    static thread_local int k = B2F_KERNEL_NONE;
    return k;
}
}  // namespace b2f

extern "C" int32_t b2f_last_flow_kernel(void) { return b2f::last_flow_kernel(); }
extern "C" const char* b2f_last_error(void) { return b2f::last_error_buffer(); }
extern "C" int32_t b2f_abi_version(void) { return 1; }
extern "C" int32_t b2f_params_per_element(int32_t tkind, int32_t n_bins) { return b2f::params_per_element(tkind, n_bins); }
extern "C" int32_t b2f_padded_params(int32_t P) { return b2f::padded_params(P); }

// b2f_flow_backward lives in b2f_flow_bwd.cu
