// Elementwise transformers given materialised parameters h (kernel K1): Affine / InverseAffine / Shift /
// rational-quadratic spline, forward, inverse and backward, with the per-row log-det reduced by warp
// shuffles.  Replaces TensorTransformer.forward / inverse of the reference
// (/root/reference/torchflows/bijections/finite/autoregressive/transformers/linear/affine.py:39-59,149-159,
//  transformers/spline/base.py:53-72, transformers/spline/rational_quadratic.py:65-200), i.e. ~60 ATen
// launches, 6 host syncs and a boolean-mask copy of h per call, with one launch that reads x and h once.
//
// Mapping: GS = smallest power of two >= min(n_event, 32) lanes cooperate on one row (32/GS rows per
// warp); a lane walks elements sub, sub+GS, ... of its row; the row's log-det is an xor-butterfly
// reduction inside the GS-lane group.  h is read in place (element-major, parameter-minor, as produced by
// the conditioner, layers_base.py:143); a row stride of 0 broadcasts one parameter set to all rows.
#include <algorithm>

#include "b2f_common.cuh"
#include "b2f_math.cuh"
#include "b2f_lrs.cuh"

namespace b2f {

struct HGlobal {
    const float* p;
    __device__ __forceinline__ float operator()(int i) const { return __ldg(p + i); }
};
struct GGlobal {
    float* p;
    __device__ __forceinline__ void operator()(int i, float v) const { p[i] = v; }
};

template <int TK, int NB, int MODE>
__global__ void __launch_bounds__(256) transformer_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                          float* __restrict__ out, float* __restrict__ log_det,
                                                          int32_t* __restrict__ k_out, long long n_rows, int E,
                                                          long long h_row_stride, int nb, float boundary, int GS) {
    const int lane = threadIdx.x & 31;
    const long long wg = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int rpw = 32 / GS;
    const long long row = wg * rpw + lane / GS;
    const int sub = lane % GS;
    const bool valid = row < n_rows;
    const int P = (TK == B2F_T_RQ_FWD || TK == B2F_T_RQ_INV) ? 3 * nb - 1 : (TK == B2F_T_LRS_FWD || TK == B2F_T_LRS_INV) ? 4 * nb
                  : ((TK == B2F_T_AFFINE_FWD || TK == B2F_T_AFFINE_INV) ? 2 : 1);
    float ld = 0.0f;
    if (valid) {
        const float* hr = h + row * h_row_stride;
        for (int e = sub; e < E; e += GS) {
            const long long idx = row * E + e;
            const float v = __ldg(x + idx);
            const float* he = hr + (long long)e * P;
            float o, l;
            if constexpr (TK == B2F_T_SHIFT_ADD) { o = v + __ldg(he); l = 0.0f; }
            else if constexpr (TK == B2F_T_SHIFT_SUB) { o = v - __ldg(he); l = 0.0f; }
            else if constexpr (TK == B2F_T_AFFINE_FWD) affine_fwd<MODE>(v, __ldg(he), __ldg(he + 1), o, l);
            else if constexpr (TK == B2F_T_AFFINE_INV) affine_inv<MODE>(v, __ldg(he), __ldg(he + 1), o, l);
            else if constexpr (TK == B2F_T_SCALE_FWD) scale_fwd<MODE>(v, __ldg(he), o, l);
            else if constexpr (TK == B2F_T_SCALE_INV) scale_inv<MODE>(v, __ldg(he), o, l);
            else if constexpr (TK == B2F_T_LRS_FWD || TK == B2F_T_LRS_INV)
                lrs_apply<NB, TK == B2F_T_LRS_INV>(v, HGlobal{he}, nb, boundary, o, l);
            else {
                int k;
                rq_apply<NB, TK == B2F_T_RQ_INV, MODE>(v, HGlobal{he}, nb, boundary, o, l, k);
                if (k_out) k_out[idx] = k;
            }
            out[idx] = o;
            ld += l;
        }
    }
    for (int o = GS >> 1; o > 0; o >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, o);
    if (valid && sub == 0 && log_det) log_det[row] = ld;
}

template <int TK, int NB, int MODE>
__global__ void __launch_bounds__(256) transformer_backward_kernel(
    const float* __restrict__ x, const float* __restrict__ h, const float* __restrict__ gout,
    const float* __restrict__ glog_det, float* __restrict__ gx, float* __restrict__ gh, long long n_rows, int E,
    long long h_row_stride, int nb, float boundary) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * E) return;
    const long long row = idx / E;
    const int e = (int)(idx - row * E);
    const int P = (TK == B2F_T_RQ_FWD || TK == B2F_T_RQ_INV) ? 3 * nb - 1 : (TK == B2F_T_LRS_FWD || TK == B2F_T_LRS_INV) ? 4 * nb
                  : ((TK == B2F_T_AFFINE_FWD || TK == B2F_T_AFFINE_INV) ? 2 : 1);
    const float* he = h + row * h_row_stride + (long long)e * P;
    float* ge = gh + idx * P;
    const float v = __ldg(x + idx);
    const float GZ = gout ? __ldg(gout + idx) : 0.0f;
    const float GL = glog_det ? __ldg(glog_det + row) : 0.0f;
    float dv;
    if constexpr (TK == B2F_T_SHIFT_ADD) { dv = GZ; ge[0] = GZ; }
    else if constexpr (TK == B2F_T_SHIFT_SUB) { dv = GZ; ge[0] = -GZ; }
    else if constexpr (TK == B2F_T_AFFINE_FWD) affine_fwd_backward<MODE>(v, __ldg(he), GZ, GL, dv, ge[0], ge[1]);
    else if constexpr (TK == B2F_T_AFFINE_INV) affine_inv_backward<MODE>(v, __ldg(he), __ldg(he + 1), GZ, GL, dv, ge[0], ge[1]);
    else if constexpr (TK == B2F_T_SCALE_FWD) { float unused; affine_fwd_backward<MODE>(v, __ldg(he), GZ, GL, dv, ge[0], unused); }
    else if constexpr (TK == B2F_T_SCALE_INV) { float unused; affine_inv_backward<MODE>(v, __ldg(he), 0.0f, GZ, GL, dv, ge[0], unused); }
    else if constexpr (TK == B2F_T_LRS_FWD || TK == B2F_T_LRS_INV)
        lrs_backward<NB, TK == B2F_T_LRS_INV>(v, HGlobal{he}, nb, boundary, GZ, GL, dv, GGlobal{ge});
    else if constexpr (TK == B2F_T_RQ_FWD) rq_backward_fwd<NB, MODE>(v, HGlobal{he}, nb, boundary, GZ, GL, dv, GGlobal{ge});
    else rq_backward_inv<NB, MODE>(v, HGlobal{he}, nb, boundary, GZ, GL, dv, GGlobal{ge});
    gx[idx] = dv;
}

// Spline backward with dense parameters (h_row_stride == E * 23, n_bins == 8): a block's 256 elements own one contiguous
// 256 x 23-float slab of h and of gh, so both travel through shared memory with coalesced 16-byte accesses (the
// thread-per-element loads above touch 32 different sectors per instruction and write 4 bytes of every sector 8 times).
template <int TK, int MODE>
__global__ void __launch_bounds__(256) transformer_backward_staged_kernel(
    const float* __restrict__ x, const float* __restrict__ h, const float* __restrict__ gout,
    const float* __restrict__ glog_det, float* __restrict__ gx, float* __restrict__ gh, long long n_elem, int E,
    float boundary) {
    constexpr int P = 23;
    __shared__ __align__(16) float buf[256 * P];
    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.x * 256;
    const int n_here = (int)min(256LL, n_elem - base);
    const int n_f = n_here * P;
    {
        const float4* src = reinterpret_cast<const float4*>(h + base * P);      // base * 92 bytes: a multiple of 16
        for (int q = tid; q < (n_f >> 2); q += 256) reinterpret_cast<float4*>(buf)[q] = __ldg(src + q);
        for (int q = (n_f & ~3) + tid; q < n_f; q += 256) buf[q] = __ldg(h + base * P + q);
    }
    __syncthreads();
    if (tid < n_here) {
        const long long idx = base + tid;
        const long long row = idx / E;
        float* mine = buf + tid * P;                 // stride 23 floats: conflict-free
        float hv[P + 1], gv[P + 1];
#pragma unroll
        for (int i = 0; i < P; ++i) { hv[i] = mine[i]; gv[i] = 0.0f; }
        const float v = __ldg(x + idx);
        const float GZ = gout ? __ldg(gout + idx) : 0.0f;
        const float GL = glog_det ? __ldg(glog_det + row) : 0.0f;
        float dv;
        auto hf = [&](int i) { return hv[i]; };
        auto gf = [&](int i, float val) { gv[i] = val; };
        if constexpr (TK == B2F_T_RQ_FWD) rq_backward_fwd<8, MODE>(v, hf, 8, boundary, GZ, GL, dv, gf);
        else rq_backward_inv<8, MODE>(v, hf, 8, boundary, GZ, GL, dv, gf);
        gx[idx] = dv;
#pragma unroll
        for (int i = 0; i < P; ++i) mine[i] = gv[i];
    }
    __syncthreads();
    {
        float4* dst = reinterpret_cast<float4*>(gh + base * P);
        for (int q = tid; q < (n_f >> 2); q += 256) __stcs(dst + q, reinterpret_cast<const float4*>(buf)[q]);
        for (int q = (n_f & ~3) + tid; q < n_f; q += 256) gh[base * P + q] = buf[q];
    }
}

__global__ void __launch_bounds__(256) column_stats_kernel(const float* __restrict__ x, double* __restrict__ sum,
                                                           double* __restrict__ sumsq, long long B, int D,
                                                           long long rows_per_block) {
    // thread j walks column j over this block's row slab (coalesced across j), fp64 accumulation
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(B, r0 + rows_per_block);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < D; j += gridDim.x * blockDim.x) {
        double s = 0.0, q = 0.0;
        for (long long r = r0; r < r1; ++r) {
            const double v = (double)__ldg(x + r * D + j);
            s += v; q += v * v;
        }
        atomicAdd(sum + j, s);
        atomicAdd(sumsq + j, q);
    }
}

template <int TK, int MODE>
static int launch_fwd(const float* x, const float* h, float* out, float* log_det, int32_t* k_out, int64_t n_rows,
                      int32_t E, int64_t hs, int32_t nb, float boundary, cudaStream_t st) {
    int GS = 1;
    while (GS < E && GS < 32) GS <<= 1;
    const long long warps = (n_rows + (32 / GS) - 1) / (32 / GS);
    const int block = 256;
    const long long blocks = (warps * 32 + block - 1) / block;
    if (blocks > 0x7fffffffLL) return fail(B2F_ERR_UNSUPPORTED, "b2f_transformer_apply: too many rows");
    constexpr bool rq = TK == B2F_T_RQ_FWD || TK == B2F_T_RQ_INV || TK == B2F_T_LRS_FWD || TK == B2F_T_LRS_INV;
    if (rq && nb == 8)
        transformer_kernel<TK, 8, MODE><<<(unsigned)blocks, block, 0, st>>>(x, h, out, log_det, k_out, n_rows, E, hs, nb, boundary, GS);
    else
        transformer_kernel<TK, 0, MODE><<<(unsigned)blocks, block, 0, st>>>(x, h, out, log_det, k_out, n_rows, E, hs, nb, boundary, GS);
    return check_launch("b2f_transformer_apply");
}

template <int TK, int MODE>
static int launch_bwd(const float* x, const float* h, const float* gout, const float* gld, float* gx, float* gh,
                      int64_t n_rows, int32_t E, int64_t hs, int32_t nb, float boundary, cudaStream_t st) {
    const long long n = n_rows * E;
    const int block = 256;
    const long long blocks = (n + block - 1) / block;
    if (blocks > 0x7fffffffLL) return fail(B2F_ERR_UNSUPPORTED, "b2f_transformer_backward: too many elements");
    constexpr bool rq = TK == B2F_T_RQ_FWD || TK == B2F_T_RQ_INV;
    constexpr bool lrs = TK == B2F_T_LRS_FWD || TK == B2F_T_LRS_INV;
    if constexpr (rq) {
        if (nb == 8 && hs == (int64_t)E * 23 && !((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(gh)) & 15)) {
            transformer_backward_staged_kernel<TK, MODE><<<(unsigned)blocks, block, 0, st>>>(x, h, gout, gld, gx, gh, n, E, boundary);
            return check_launch("b2f_transformer_backward");
        }
    }
    if ((rq || lrs) && nb == 8)
        transformer_backward_kernel<TK, 8, MODE><<<(unsigned)blocks, block, 0, st>>>(x, h, gout, gld, gx, gh, n_rows, E, hs, nb, boundary);
    else
        transformer_backward_kernel<TK, 0, MODE><<<(unsigned)blocks, block, 0, st>>>(x, h, gout, gld, gx, gh, n_rows, E, hs, nb, boundary);
    return check_launch("b2f_transformer_backward");
}

}  // namespace b2f

using namespace b2f;

extern "C" int b2f_transformer_apply(int32_t tkind, const float* x, const float* h, float* out, float* log_det,
                                     int32_t* k_out, int64_t n_rows, int32_t n_event, int64_t h_row_stride,
                                     int32_t n_bins, float boundary, int32_t flags, void* stream) {
    if (n_rows == 0 && n_event > 0) return B2F_OK;
    if (!x || !h || !out || n_rows < 0 || n_event <= 0 || h_row_stride < 0)
        return fail(B2F_ERR_INVALID, "b2f_transformer_apply: bad arguments");
    const bool rq = tkind == B2F_T_RQ_FWD || tkind == B2F_T_RQ_INV || tkind == B2F_T_LRS_FWD || tkind == B2F_T_LRS_INV;
    if (rq && (n_bins < 1 || n_bins > kRqMaxBins || !(boundary > 0.0f)))
        return fail(B2F_ERR_UNSUPPORTED, "b2f_transformer_apply: n_bins=%d (1..%d) boundary=%g", n_bins, kRqMaxBins, boundary);
    if (n_rows == 0) return B2F_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool precise = flags & B2F_FLOW_MODE_PRECISE;
#define B2F_DISPATCH(TKV)                                                                                          \
    case TKV:                                                                                                      \
        return precise ? launch_fwd<TKV, 0>(x, h, out, log_det, k_out, n_rows, n_event, h_row_stride, n_bins, boundary, st) \
                       : launch_fwd<TKV, 1>(x, h, out, log_det, k_out, n_rows, n_event, h_row_stride, n_bins, boundary, st);
    switch (tkind) {
        B2F_DISPATCH(B2F_T_SHIFT_ADD)
        B2F_DISPATCH(B2F_T_SHIFT_SUB)
        B2F_DISPATCH(B2F_T_AFFINE_FWD)
        B2F_DISPATCH(B2F_T_AFFINE_INV)
        B2F_DISPATCH(B2F_T_RQ_FWD)
        B2F_DISPATCH(B2F_T_RQ_INV)
        B2F_DISPATCH(B2F_T_LRS_FWD)
        B2F_DISPATCH(B2F_T_LRS_INV)
        B2F_DISPATCH(B2F_T_SCALE_FWD)
        B2F_DISPATCH(B2F_T_SCALE_INV)
    }
#undef B2F_DISPATCH
    return fail(B2F_ERR_INVALID, "b2f_transformer_apply: unknown transformer kind %d", tkind);
}

extern "C" int b2f_transformer_backward(int32_t tkind, const float* x, const float* h, const float* gout,
                                        const float* glog_det, float* gx, float* gh, int64_t n_rows, int32_t n_event,
                                        int64_t h_row_stride, int32_t n_bins, float boundary, int32_t flags,
                                        void* stream) {
    if (n_rows == 0 && n_event > 0) return B2F_OK;
    if (!x || !h || !gx || !gh || n_rows < 0 || n_event <= 0 || h_row_stride < 0)
        return fail(B2F_ERR_INVALID, "b2f_transformer_backward: bad arguments");
    if ((tkind == B2F_T_RQ_FWD || tkind == B2F_T_RQ_INV || tkind == B2F_T_LRS_FWD || tkind == B2F_T_LRS_INV) && (n_bins < 1 || n_bins > kRqMaxBins || !(boundary > 0.0f)))
        return fail(B2F_ERR_UNSUPPORTED, "b2f_transformer_backward: n_bins=%d", n_bins);
    if (n_rows == 0) return B2F_OK;
    cudaStream_t st = (cudaStream_t)stream;
    (void)flags;
#define B2F_DISPATCH(TKV) \
    case TKV: return launch_bwd<TKV, 0>(x, h, gout, glog_det, gx, gh, n_rows, n_event, h_row_stride, n_bins, boundary, st);
    switch (tkind) {
        B2F_DISPATCH(B2F_T_SHIFT_ADD)
        B2F_DISPATCH(B2F_T_SHIFT_SUB)
        B2F_DISPATCH(B2F_T_AFFINE_FWD)
        B2F_DISPATCH(B2F_T_AFFINE_INV)
        B2F_DISPATCH(B2F_T_RQ_FWD)
        B2F_DISPATCH(B2F_T_RQ_INV)
        B2F_DISPATCH(B2F_T_LRS_FWD)
        B2F_DISPATCH(B2F_T_LRS_INV)
        B2F_DISPATCH(B2F_T_SCALE_FWD)
        B2F_DISPATCH(B2F_T_SCALE_INV)
    }
#undef B2F_DISPATCH
    return fail(B2F_ERR_INVALID, "b2f_transformer_backward: unknown transformer kind %d", tkind);
}

extern "C" int b2f_column_stats(const float* x, double* sum, double* sumsq, int64_t B, int32_t D, void* stream) {
    if (B == 0 && D > 0) return B2F_OK;
    if (!x || !sum || !sumsq || B < 0 || D <= 0) return fail(B2F_ERR_INVALID, "b2f_column_stats: bad arguments");
    if (B == 0) return B2F_OK;
    const int block = 128;
    const int gx = (D + block - 1) / block;
    long long slabs = std::min<long long>((B + 255) / 256, std::max(1, 148 * 8 / gx));
    const long long rpb = (B + slabs - 1) / slabs;
    slabs = (B + rpb - 1) / rpb;
    dim3 grid(gx, (unsigned)slabs);
    column_stats_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, sum, sumsq, B, D, rpb);
    return check_launch("b2f_column_stats");
}
