// Diagnostic entry point: one tcgen05 GEMM tile C[128 x N] = A[128 x K] * B[N x K]^T (tf32 inputs, fp32
// accumulate in TMEM) built from exactly the helpers and operand layout the fused coupling kernel uses
// (b2f_umma.cuh).  tests/test_gpu_tc.py checks it against torch.matmul so that a descriptor or layout mistake
// shows up here, in isolation, rather than inside the fused kernel.
#include "b2f_common.cuh"
#include "b2f_umma.cuh"

namespace b2f {

__global__ void __launch_bounds__(128) umma_gemm_debug_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                              float* __restrict__ C, int N, int K, int tmem_cols) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sA = smem_raw;                       // 128 x K floats, canonical layout
    uint8_t* sB = sA + 128 * K * 4;               // N x K floats
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int idx = tid; idx < 128 * (K / 4); idx += blockDim.x) {
        const int m = idx / (K / 4), kc = idx % (K / 4);
        *reinterpret_cast<float4*>(sA + umma::canon_off(m, 4 * kc, K)) = *reinterpret_cast<const float4*>(A + (size_t)m * K + 4 * kc);
    }
    for (int idx = tid; idx < N * (K / 4); idx += blockDim.x) {
        const int n = idx / (K / 4), kc = idx % (K / 4);
        *reinterpret_cast<float4*>(sB + umma::canon_off(n, 4 * kc, K)) = *reinterpret_cast<const float4*>(B + (size_t)n * K + 4 * kc);
    }
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_barrier_init(); }
    umma::fence_proxy_async_smem();               // generic-proxy smem writes -> visible to the tensor core (async proxy)
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, tmem_cols);
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tc_fence_after_sync();
    const uint32_t tbase = tmem_base_s;

    if (tid == 0) {
        const uint32_t idesc = umma::make_idesc_tf32(128, N);
        const uint32_t sbo = K * 32;
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t ad = umma::make_smem_desc(umma::smem_u32(sA) + ks * 256, 128, sbo);
            const uint64_t bd = umma::make_smem_desc(umma::smem_u32(sB) + ks * 256, 128, sbo);
            umma::mma_tf32_ss(tbase, ad, bd, idesc, ks > 0);
        }
        umma::mma_commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::tc_fence_after_sync();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 8) {
        float v[8];
        umma::tmem_ld8(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) C[(size_t)row * N + c0 + i] = v[i];
    }
    umma::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, tmem_cols);
}

}  // namespace b2f

using namespace b2f;

extern "C" int b2f_debug_umma_gemm(const float* A, const float* B, float* C, int32_t N, int32_t K, void* stream) {
    if (!A || !B || !C || N < 16 || N > 256 || N % 16 || K < 8 || K % 8)
        return fail(B2F_ERR_INVALID, "b2f_debug_umma_gemm: need 16 <= N <= 256, N %% 16 == 0, K %% 8 == 0");
    const size_t smem = (size_t)(128 + N) * K * 4;
    if (smem > 200 * 1024) return fail(B2F_ERR_UNSUPPORTED, "b2f_debug_umma_gemm: tile too large");
    int cols = 32;
    while (cols < N) cols <<= 1;
    cudaError_t ce = (cudaError_t)raise_smem_limit((const void*)umma_gemm_debug_kernel, smem);
    if (ce != cudaSuccess) return fail(B2F_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
    umma_gemm_debug_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, C, N, K, cols);
    return check_launch("b2f_debug_umma_gemm");
}
