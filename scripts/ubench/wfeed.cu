// Micro-benchmark: how to feed warp-uniform weights to a row-per-thread GEMV (the output layer of the rows kernel:
// acc[p] = sum_j W2[e][p][j] * hid[j], 128 elements x 2 parameters x 8 hidden slots per row and layer).
//   A: weights in shared memory, one broadcast LDS.128 per 4 weights (what b2f_flow_rows.cuh does)
//   B: weights in __constant__ memory indexed by the (uniform) element counter
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/ubench/wfeed.cu -o scripts/ubench/wfeed && ./wfeed
#include <cstdio>
#include <cuda_runtime.h>

constexpr int E = 128, P = 2, HP = 8, NW = E * P * HP;      // 2048 floats = 8 KB
__constant__ float4 cw[NW / 4];

template <int MODE>
__global__ void __launch_bounds__(128) k(const float* __restrict__ wg, const float* __restrict__ x, float* __restrict__ out, int reps) {
    __shared__ float4 sw[NW / 4];
    for (int i = threadIdx.x; i < NW / 4; i += blockDim.x) sw[i] = reinterpret_cast<const float4*>(wg)[i];
    __syncthreads();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    float hid[HP];
#pragma unroll
    for (int j = 0; j < HP; ++j) hid[j] = x[row * HP + j];
    float total = 0.0f;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 4
        for (int e = 0; e < E; ++e) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float a = 0.0f;
#pragma unroll
                for (int j4 = 0; j4 < HP / 4; ++j4) {
                    const float4 w = MODE == 0 ? sw[(e * P + p) * (HP / 4) + j4] : cw[(e * P + p) * (HP / 4) + j4];
                    a = fmaf(w.x, hid[4 * j4 + 0], a);
                    a = fmaf(w.y, hid[4 * j4 + 1], a);
                    a = fmaf(w.z, hid[4 * j4 + 2], a);
                    a = fmaf(w.w, hid[4 * j4 + 3], a);
                }
                total += a;
            }
        }
        hid[r & 7] += total * 1e-9f;
    }
    out[row] = total;
}

int main() {
    const int rows = 1 << 18, reps = 8;
    float *wg, *x, *out;
    cudaMalloc(&wg, NW * 4); cudaMalloc(&x, (size_t)rows * HP * 4); cudaMalloc(&out, rows * 4);
    cudaMemset(wg, 0, NW * 4); cudaMemset(x, 0, (size_t)rows * HP * 4);
    cudaMemcpyToSymbol(cw, wg, NW * 4, 0, cudaMemcpyDeviceToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<rows / 128, 128>>>(wg, x, out, reps); else k<1><<<rows / 128, 128>>>(wg, x, out, reps);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it == 2) printf("%s: %.3f ms for %d rows x %d reps x %d FFMA  (%.1f GFFMA/s)  err=%s\n", mode == 0 ? "A shared LDS.128 " : "B __constant__    ",
                                ms, rows, reps, NW, (double)rows * reps * NW / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
