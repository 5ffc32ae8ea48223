// Small non-kernel entry points of the C ABI.
#include <map>
#include <mutex>
#include <utility>

#include "b2f_common.cuh"

namespace b2f {
char* last_error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}
int raise_smem_limit(const void* kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> limit;
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return (int)ce;
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = limit[std::make_pair(dev, kernel)];
    if (bytes <= cur) return (int)cudaSuccess;
    ce = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (ce == cudaSuccess) cur = bytes;
    return (int)ce;
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
