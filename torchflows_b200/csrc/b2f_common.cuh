// Shared host-side helpers of libb2f.so: error reporting, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "b2f.h"

namespace b2f {

char* last_error_buffer();   // thread-local, defined in b2f_api.cu
int& last_flow_kernel();     // thread-local, which kernel b2f_flow_apply launched last (B2F_KERNEL_*)

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B2F_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return B2F_OK;
}

// Raise (never lower) a kernel's dynamic shared-memory limit.  The limit is state of the FUNCTION, not of a launch: a kernel
// node replayed from a CUDA graph keeps the size it was captured with, so a later, smaller eager launch must not shrink
// the limit under it.  Per device; the attribute call is skipped when the limit is already high enough.
int raise_smem_limit(const void* kernel, size_t bytes);   // b2f_api.cu; returns a cudaError_t value

// in-kernel base draws of b2f_flow_sample (csrc/b2f_philox.cuh)
struct TcqNoise {
    unsigned long long seed, offset;
    const float* base_loc;
    const float* base_log_scale;
};

inline int params_per_element(int tkind, int n_bins) {
    switch (tkind) {
        case B2F_T_SHIFT_ADD: case B2F_T_SHIFT_SUB: return 1;
        case B2F_T_AFFINE_FWD: case B2F_T_AFFINE_INV: return 2;
        case B2F_T_RQ_FWD: case B2F_T_RQ_INV: return 3 * n_bins - 1;
        case B2F_T_LRS_FWD: case B2F_T_LRS_INV: return 4 * n_bins;
        case B2F_T_SCALE_FWD: case B2F_T_SCALE_INV: return 1;
        default: return -1;
    }
}
inline int padded_params(int P) { return P <= 2 ? P : (P + 3) / 4 * 4; }

}  // namespace b2f
