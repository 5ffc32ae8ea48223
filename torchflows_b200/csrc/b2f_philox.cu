// Materialised draws of the base distribution from the counter-based stream of csrc/b2f_philox.cuh:
// out[r, c] = loc[c] + exp(log_scale[c]) * n(r * D + c)   (DiagonalGaussian.sample, base_distributions/gaussian.py:41-44).
#include <algorithm>

#include "b2f_common.cuh"
#include "b2f_philox.cuh"

namespace b2f {

__global__ void __launch_bounds__(256) philox_normal_kernel(float* __restrict__ out, long long n, int D, const float* __restrict__ loc,
                                                            const float* __restrict__ log_scale, uint64_t seed, uint64_t offset) {
    const long long n4 = (n + 3) / 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += stride) {
        const float4 v = philox::normal4((uint64_t)g, seed, offset);
        const float vv[4] = {v.x, v.y, v.z, v.w};
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long e = 4 * g + u;
            const int c = (int)(e % D);
            const float sc = log_scale ? expf(__ldg(log_scale + c)) : 1.0f;
            o[u] = fmaf(sc, vv[u], loc ? __ldg(loc + c) : 0.0f);
        }
        if (4 * g + 3 < n && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
            *reinterpret_cast<float4*>(out + 4 * g) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (4 * g + u < n) out[4 * g + u] = o[u];
        }
    }
}

}  // namespace b2f

extern "C" int b2f_philox_normal(float* out, int64_t n_rows, int32_t D, const float* loc, const float* log_scale, uint64_t seed,
                                 uint64_t offset, void* stream) {
    using namespace b2f;
    if (n_rows < 0 || D < 1) return fail(B2F_ERR_INVALID, "b2f_philox_normal: shape");
    const long long n = (long long)n_rows * D;
    if (n == 0) return B2F_OK;
    if (!out) return fail(B2F_ERR_INVALID, "b2f_philox_normal: null output");
    const long long n4 = (n + 3) / 4;
    const unsigned grid = (unsigned)std::min<long long>((n4 + 255) / 256, 148 * 16);
    philox_normal_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, n, D, loc, log_scale, seed, offset);
    return check_launch("b2f_philox_normal");
}
