// Runs of per-column layers (ElementwiseAffine / ActNorm with global parameters, ReversePermutation) at ANY event size,
// as one pass over the batch: the run composes into y[r, c] = A[c] * x[r, s(c)] + C[c] with s = identity or reversal.
// Used where such layers sit between layers that are not part of a whole-flow program (the wide-conditioner coupling
// layers of csrc/b2f_wide.cu at n_dim = 1024, where the whole-flow kernels' shared-memory tile does not fit).
//
// Replaces (file:line relative to /root/reference/torchflows/bijections/finite): autoregressive/layers_base.py:300-318
// (ElementwiseBijection.forward / inverse with value repeated over the batch), autoregressive/transformers/linear/
// affine.py:33-59 (Affine / InverseAffine), autoregressive/layers.py:39-69 (ActNorm after initialisation),
// matrix/permutation.py:19-37 (ReversePermutationMatrix), and their autograd backward.
#include <string.h>

#include <algorithm>

#include "b2f_common.cuh"
#include "b2f_math.cuh"

namespace b2f {
namespace colrun {

constexpr int kMaxOps = 8;
constexpr int kRowsPerBlock = 64;

struct Ops {
    int n;
    int kind[kMaxOps];            // B2F_COL_AFFINE_FWD / _INV / _FLIP
    const float* value[kMaxOps];  // (D, 2): unconstrained scale, shift
    float* gvalue[kMaxOps];       // backward only (nullable)
};

// scale and shift of one affine op at one column: z = a x + b   (affine.py:33-59; InverseAffine swaps the directions)
__device__ __forceinline__ void coeff(int kind, const float* __restrict__ value, int col, float& a, float& b, float& alpha) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(value) + col);
    alpha = expf(kAffineC0 + 0.5f * v.x) + kAffineM;
    if (kind == B2F_COL_AFFINE_FWD) { a = alpha; b = v.y; }
    else { a = 1.0f / alpha; b = -v.y * a; }
}

// composed map of output column c: input column and (A, C)
__device__ __forceinline__ void compose(const Ops& P, int D, int c, int& in_col, float& A, float& C) {
    int colk[kMaxOps];
    int cur = c;
#pragma unroll
    for (int k = kMaxOps - 1; k >= 0; --k) {
        if (k < P.n) {
            colk[k] = cur;
            if (P.kind[k] == B2F_COL_FLIP) cur = D - 1 - cur;
        }
    }
    in_col = cur;
    A = 1.0f; C = 0.0f;
#pragma unroll
    for (int k = 0; k < kMaxOps; ++k) {
        if (k < P.n && P.kind[k] != B2F_COL_FLIP) {
            float a, b, alpha;
            coeff(P.kind[k], P.value[k], colk[k], a, b, alpha);
            A *= a;
            C = fmaf(a, C, b);
        }
    }
}

__global__ void __launch_bounds__(256) apply_kernel(const Ops P, const float* __restrict__ x, float* __restrict__ y, long long B, int D) {
    const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (c >= D) return;
    float A[4], C[4];
    int in0 = 0, in3 = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        int ic;
        compose(P, D, c + u, ic, A[u], C[u]);
        if (u == 0) in0 = ic;
        if (u == 3) in3 = ic;
    }
    const bool flipped = in3 < in0;                 // the four inputs are contiguous either way
    const int base = flipped ? in3 : in0;
    const long long r0 = (long long)blockIdx.y * kRowsPerBlock, r1 = min(B, r0 + kRowsPerBlock);
    for (long long r = r0; r < r1; ++r) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * D + base));
        float4 o;
        o.x = fmaf(A[0], flipped ? v.w : v.x, C[0]);
        o.y = fmaf(A[1], flipped ? v.z : v.y, C[1]);
        o.z = fmaf(A[2], flipped ? v.y : v.z, C[2]);
        o.w = fmaf(A[3], flipped ? v.x : v.w, C[3]);
        *reinterpret_cast<float4*>(y + r * D + c) = o;
    }
}

// sum over columns of log A[c] (the run's log-determinant, the same for every row); one block, fixed summation order
__global__ void __launch_bounds__(256) logdet_kernel(const Ops P, int D, float* __restrict__ out) {
    __shared__ float s[256];
    float acc = 0.0f;
    for (int c = threadIdx.x; c < D; c += 256) {
        int ic;
        float A, C;
        compose(P, D, c, ic, A, C);
        acc += logf(A);
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s[0];
}

// gx[r, s(c)] = A[c] gy[r, c];  sums[0][c] += sum_r gy[r, c] x[r, s(c)],  sums[1][c] += sum_r gy[r, c]
__global__ void __launch_bounds__(256) backward_kernel(const Ops P, const float* __restrict__ x, const float* __restrict__ gy,
                                                       float* __restrict__ gx, float* __restrict__ sums, long long B, int D) {
    const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (c >= D) return;
    float A[4], C[4];
    int in0 = 0, in3 = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        int ic;
        compose(P, D, c + u, ic, A[u], C[u]);
        if (u == 0) in0 = ic;
        if (u == 3) in3 = ic;
    }
    const bool flipped = in3 < in0;
    const int base = flipped ? in3 : in0;
    float sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
    const long long r0 = (long long)blockIdx.y * kRowsPerBlock, r1 = min(B, r0 + kRowsPerBlock);
    for (long long r = r0; r < r1; ++r) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gy + r * D + c));
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * D + base));
        const float xi[4] = {flipped ? v.w : v.x, flipped ? v.z : v.y, flipped ? v.y : v.z, flipped ? v.x : v.w};
        const float gg[4] = {g.x, g.y, g.z, g.w};
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            o[u] = A[u] * gg[u];
            sa[u] = fmaf(gg[u], xi[u], sa[u]);
            sb[u] += gg[u];
        }
        *reinterpret_cast<float4*>(gx + r * D + base) = flipped ? make_float4(o[3], o[2], o[1], o[0]) : make_float4(o[0], o[1], o[2], o[3]);
    }
    if (sums) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            atomicAdd(sums + c + u, sa[u]);
            atomicAdd(sums + D + c + u, sb[u]);
        }
    }
}

// per output column: chain rule from (dL/dA, dL/dC) of the composed map back to every op's (unconstrained scale, shift)
__global__ void __launch_bounds__(256) finalize_kernel(const Ops P, const float* __restrict__ sums, const float* __restrict__ g_logdet,
                                                       int D) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= D) return;
    int colk[kMaxOps];
    float a_k[kMaxOps], b_k[kMaxOps], al_k[kMaxOps], Ap[kMaxOps], Cp[kMaxOps];      // op coefficients, composed map BEFORE op k
    int cur = c;
#pragma unroll
    for (int k = kMaxOps - 1; k >= 0; --k) {
        if (k < P.n) {
            colk[k] = cur;
            if (P.kind[k] == B2F_COL_FLIP) cur = D - 1 - cur;
        }
    }
    float A = 1.0f, C = 0.0f;
#pragma unroll
    for (int k = 0; k < kMaxOps; ++k) {
        if (k < P.n && P.kind[k] != B2F_COL_FLIP) {
            coeff(P.kind[k], P.value[k], colk[k], a_k[k], b_k[k], al_k[k]);
            Ap[k] = A; Cp[k] = C;
            A *= a_k[k];
            C = fmaf(a_k[k], C, b_k[k]);
        }
    }
    // log-det of the run = sum_c log A[c]; its upstream gradient is one scalar (sum over the rows)
    float GA = sums[c] + (g_logdet ? g_logdet[0] / A : 0.0f), GC = sums[D + c];
#pragma unroll
    for (int k = kMaxOps - 1; k >= 0; --k) {
        if (k < P.n && P.kind[k] != B2F_COL_FLIP) {
            const float ga = fmaf(GA, Ap[k], GC * Cp[k]), gb = GC;
            GA *= a_k[k];
            GC *= a_k[k];
            if (P.gvalue[k]) {
                const float alpha = al_k[k];
                float g0, g1;
                if (P.kind[k] == B2F_COL_AFFINE_FWD) {              // a = alpha, b = v1
                    g0 = ga * (alpha - kAffineM) * 0.5f;
                    g1 = gb;
                } else {                                            // a = 1 / alpha, b = -v1 / alpha
                    const float v1 = -b_k[k] * alpha;
                    const float ia = a_k[k];
                    const float galpha = (-ga + gb * v1) * ia * ia;
                    g0 = galpha * (alpha - kAffineM) * 0.5f;
                    g1 = -gb * ia;
                }
                reinterpret_cast<float2*>(P.gvalue[k])[colk[k]] = make_float2(g0, g1);
            }
        }
    }
}

// DiagonalGaussian.log_prob of rows (base_distributions/gaussian.py:46-54): lp[r] = sum_c -(t^2 / 2 + log(2 pi) / 2 + ls_c),
// t = (z - loc_c) exp(-ls_c); one warp per row.  Backward: gz[r, c] = -g[r] t exp(-ls_c).
__global__ void __launch_bounds__(256) gauss_logp_kernel(const float* __restrict__ z, const float* __restrict__ loc,
                                                         const float* __restrict__ ls, float* __restrict__ lp, long long B, int D) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < B; r += n_warps) {
        float acc = 0.0f;
        for (int c = lane; c < D; c += 32) acc += gauss_logp(__ldg(z + r * D + c), loc ? __ldg(loc + c) : 0.0f, ls ? __ldg(ls + c) : 0.0f);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) lp[r] = acc;
    }
}

__global__ void __launch_bounds__(256) gauss_logp_backward_kernel(const float* __restrict__ z, const float* __restrict__ loc,
                                                                  const float* __restrict__ ls, const float* __restrict__ g,
                                                                  float* __restrict__ gz, long long B, int D) {
    const long long n = B * D, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long r = i / D;
        const int c = (int)(i - r * D);
        const float is = ls ? expf(-__ldg(ls + c)) : 1.0f;
        const float t = (__ldg(z + i) - (loc ? __ldg(loc + c) : 0.0f)) * is;
        gz[i] = -__ldg(g + r) * t * is;
    }
}

static int make_ops(Ops& P, const b2f_colop_t* ops, int32_t n_ops, int32_t D, bool backward) {
    if (!ops || n_ops < 1 || n_ops > kMaxOps) return fail(B2F_ERR_INVALID, "column run: 1..%d ops", kMaxOps);
    if (D < 4 || D % 4 != 0) return fail(B2F_ERR_UNSUPPORTED, "column run: n_dim must be a multiple of 4");
    memset(&P, 0, sizeof(P));
    P.n = n_ops;
    for (int i = 0; i < n_ops; ++i) {
        P.kind[i] = ops[i].kind;
        if (ops[i].kind == B2F_COL_FLIP) continue;
        if (ops[i].kind != B2F_COL_AFFINE_FWD && ops[i].kind != B2F_COL_AFFINE_INV) return fail(B2F_ERR_INVALID, "column run: op %d kind", i);
        if (!ops[i].value) return fail(B2F_ERR_INVALID, "column run: op %d has no parameters", i);
        P.value[i] = ops[i].value;
        P.gvalue[i] = backward ? ops[i].gvalue : nullptr;
    }
    return B2F_OK;
}

}  // namespace colrun
}  // namespace b2f

using namespace b2f;
using namespace b2f::colrun;

extern "C" int b2f_column_run_apply(const b2f_colop_t* ops, int32_t n_ops, const float* x, float* y, float* log_det_sum, int64_t B,
                                    int32_t D, void* stream) {
    Ops P;
    int rc = make_ops(P, ops, n_ops, D, false);
    if (rc != B2F_OK) return rc;
    if (B < 0 || (B > 0 && (!x || !y))) return fail(B2F_ERR_INVALID, "column run: null buffer");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) return fail(B2F_ERR_INVALID, "column run: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (log_det_sum) logdet_kernel<<<1, 256, 0, st>>>(P, D, log_det_sum);
    if (B > 0) {
        dim3 grid((D / 4 + 255) / 256, (unsigned)((B + kRowsPerBlock - 1) / kRowsPerBlock));
        apply_kernel<<<grid, 256, 0, st>>>(P, x, y, B, D);
    }
    return check_launch("b2f_column_run_apply");
}

extern "C" int b2f_column_run_backward(const b2f_colop_t* ops, int32_t n_ops, const float* x, const float* gy, const float* g_log_det_sum,
                                       float* gx, float* scratch, int64_t B, int32_t D, void* stream) {
    Ops P;
    int rc = make_ops(P, ops, n_ops, D, true);
    if (rc != B2F_OK) return rc;
    if (!x || !gy || !gx || !scratch) return fail(B2F_ERR_INVALID, "column run backward: null buffer");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(gx)) & 15)
        return fail(B2F_ERR_INVALID, "column run: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(scratch, 0, (size_t)2 * D * sizeof(float), st);
    if (B > 0) {
        dim3 grid((D / 4 + 255) / 256, (unsigned)((B + kRowsPerBlock - 1) / kRowsPerBlock));
        backward_kernel<<<grid, 256, 0, st>>>(P, x, gy, gx, scratch, B, D);
    }
    finalize_kernel<<<(D + 255) / 256, 256, 0, st>>>(P, scratch, g_log_det_sum, D);
    return check_launch("b2f_column_run_backward");
}

extern "C" int b2f_gauss_log_prob(const float* z, const float* loc, const float* log_scale, float* log_prob, int64_t B, int32_t D,
                                  void* stream) {
    if (B < 0 || D < 1) return fail(B2F_ERR_INVALID, "b2f_gauss_log_prob: shape");
    if (B == 0) return B2F_OK;
    if (!z || !log_prob) return fail(B2F_ERR_INVALID, "b2f_gauss_log_prob: null buffer");
    const unsigned grid = (unsigned)std::min<long long>((B + 7) / 8, 148 * 8);
    gauss_logp_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, loc, log_scale, log_prob, B, D);
    return check_launch("b2f_gauss_log_prob");
}

extern "C" int b2f_gauss_log_prob_backward(const float* z, const float* loc, const float* log_scale, const float* g_log_prob, float* gz,
                                           int64_t B, int32_t D, void* stream) {
    if (B < 0 || D < 1) return fail(B2F_ERR_INVALID, "b2f_gauss_log_prob_backward: shape");
    if (B == 0) return B2F_OK;
    if (!z || !g_log_prob || !gz) return fail(B2F_ERR_INVALID, "b2f_gauss_log_prob_backward: null buffer");
    const unsigned grid = (unsigned)std::min<long long>((B * D + 255) / 256, 148 * 16);
    gauss_logp_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, loc, log_scale, g_log_prob, gz, B, D);
    return check_launch("b2f_gauss_log_prob_backward");
}
