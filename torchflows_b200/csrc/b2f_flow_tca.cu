// Affine / shift coupling whole-flow kernel ("tca"): RealNVP / NICE-style programs (ElementwiseAffine / ActNorm,
// ReversePermutation, affine or shift coupling layers) in ONE persistent launch, the conditioner on tcgen05 with
// SEVERAL 128-row tiles in flight per SM.
//
// Replaces (file:line relative to /root/reference/torchflows): bijections/base.py:203-232 (composition),
// bijections/finite/autoregressive/layers_base.py:119-163 (CouplingBijection), conditioning/transforms.py:293-307
// (FeedForward: Linear -> Tanh -> Linear), transformers/linear/affine.py:33-59,149-159 (Affine, Shift),
// flows.py:628-648 + base_distributions/gaussian.py:46-54 (log_prob).
//
// Why another kernel: for these presets the conditioner is tiny (RealNVP-64: 32 -> 9 -> 64), so one tile is a chain of
// DEPENDENT short steps (GEMM1 -> tanh -> GEMM2 -> affine, per layer), each a tensor-core round trip of several hundred
// cycles.  The first-generation tensor-core kernel ran one tile per SM and was latency-bound (1.2 ms per 2^20 rows);
// the row-per-thread FFMA kernel (b2f_flow_rows) is bound by shared-memory wavefronts (0.33 ms).  Here a CTA runs up to
// four independent tile pipelines ("groups": 4 epilogue warps = 128 rows = one tile each); one MMA-issuing warp and one
// tile-IO warp serve whichever group is ready, so a group's round trips hide behind the other groups' arithmetic.
//
// Precision: affine flows need fp32-faithful conditioners, so every product runs as the 3xTF32
// split hi*hi + lo*hi + hi*lo (fp32 accumulation in tensor memory): activations are split by the epilogue warps
// (x_lo = x - trunc_tf32(x); the tensor core itself truncates the fp32 tile to tf32, which is the hi part), weights by
// the host (torchflows_b200/_tca.py).
//
// Shared memory: all layers' operands (loaded once per CTA), then per group the tile (two canonical [128 x D/2]
// halves: resident activations AND the hi operand of GEMM1) and one scratch operand (x_lo for GEMM1, then
// tanh(hidden) hi / lo for GEMM2).  Tensor memory per group: D1 [128 x N1], D2 [128 x N2].
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "b2f_flow_device.cuh"
#include "b2f_philox.cuh"
#include "b2f_umma.cuh"

namespace b2f {

constexpr int kAMaxGroups = 4;
constexpr int kAThreads = 4 * kAMaxGroups * 32;
constexpr int kAHdr = 8;
constexpr int kAMaxLayers = 8;
constexpr float kALog2e = 1.4426950408889634f;
constexpr float kALn2 = 0.6931471805599453f;

struct ALayer {
    const float* blob;
    int H, N1, K2, N2, P, inverse, src_half;
    int w_off;              // float offset of this layer's operands in the shared-memory weight region
};

struct AArgs {
    ALayer layers[kAMaxLayers];
    int n_layers, D, flags, n_tiles, use_tma, n_groups, w_floats, scratch_floats, tmem_cols, tmem_stride, d2_col;
    long long B;
    const float* x;
    float* y;
    float* log_det;
    float* log_prob;
    const float* prog;
    int philox;
    unsigned long long seed, offset;
    const float* base_loc;
    const float* base_log_scale;
};

enum { AB_X_FULL = 0, AB_A1_READY, AB_D1_FULL, AB_A2_READY, AB_D2_FULL, AB_TILE_DONE, AB_PER_GROUP };

__device__ __forceinline__ bool a_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(umma::smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void a_group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ float a_tf32_rn(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ float a_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float a_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float a_rcp(float x) {          // reciprocal + one Newton step (~1 ulp)
    float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return fmaf(y, fmaf(-x, y, 1.0f), y);
}
__device__ __forceinline__ int a_hdr(const float* blob, int i) { return __ldg(reinterpret_cast<const int*>(blob) + i); }
__device__ __forceinline__ float4 a_lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void a_sts128(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void a_tma_load4(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void a_tma_store4(const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(src) : "memory");
}

// tile index of the it-th tile of group g of this CTA
__device__ __forceinline__ long long a_tile(int it, int g, int n_groups) {
    return (long long)blockIdx.x * n_groups + g + (long long)it * gridDim.x * n_groups;
}

__global__ void __launch_bounds__(kAThreads, 1)
flow_tca_kernel(const __grid_constant__ AArgs A, const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int D = A.D, Dh = D >> 1, NG = A.n_groups;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    float* wreg = reinterpret_cast<float*>(smem_raw);
    const uint32_t half_bytes = 128u * Dh * 4;
    const uint32_t group_bytes = 2 * half_bytes + (uint32_t)A.scratch_floats * 4;
    const uint32_t g_base0 = umma::smem_u32(smem_raw) + (uint32_t)A.w_floats * 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)A.w_floats * 4 + (size_t)NG * group_bytes);
    uint64_t* w_full = bars + kAMaxGroups * AB_PER_GROUP;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_full + 1);

    if (tid == 0) {
        for (int g = 0; g < kAMaxGroups; ++g) {
            umma::mbar_init(&bars[g * AB_PER_GROUP + AB_X_FULL], 1);
            umma::mbar_init(&bars[g * AB_PER_GROUP + AB_A1_READY], 4);
            umma::mbar_init(&bars[g * AB_PER_GROUP + AB_D1_FULL], 1);
            umma::mbar_init(&bars[g * AB_PER_GROUP + AB_A2_READY], 4);
            umma::mbar_init(&bars[g * AB_PER_GROUP + AB_D2_FULL], 1);
            umma::mbar_init(&bars[g * AB_PER_GROUP + AB_TILE_DONE], 4);
        }
        umma::mbar_init(w_full, 1);
        umma::fence_barrier_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_ptr, A.tmem_cols);
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tc_fence_after_sync();
    const uint32_t tbase = *tmem_ptr;
    const int s_first = A.layers[0].src_half;
    const int L = A.n_layers;

    // tiles per group
    auto tiles_of = [&](int g) {
        const long long first = (long long)blockIdx.x * NG + g, step = (long long)gridDim.x * NG;
        return first >= A.n_tiles ? 0 : (int)((A.n_tiles - 1 - first) / step + 1);
    };

    // every layer's operands, once per CTA
    if (tid == 0) {
        uint32_t wbytes = 0;
        for (int li = 0; li < L; ++li) wbytes += (uint32_t)(2 * A.layers[li].N1 * Dh + 2 * A.layers[li].N2 * A.layers[li].K2) * 4;
        umma::mbar_arrive_expect_tx(w_full, wbytes);
        for (int li = 0; li < L; ++li) {
            const ALayer& Ly = A.layers[li];
            const uint32_t w1b = (uint32_t)(2 * Ly.N1 * Dh) * 4, w2b = (uint32_t)(2 * Ly.N2 * Ly.K2) * 4;
            umma::bulk_g2s(wreg + Ly.w_off, Ly.blob + kAHdr, w1b, w_full);
            umma::bulk_g2s(wreg + Ly.w_off + 2 * Ly.N1 * Dh, Ly.blob + kAHdr + 2 * Ly.N1 * Dh + 32, w2b, w_full);
        }
    }
    if ((warp >> 2) < NG) {
        // ===================== epilogue warps: group g = one tile pipeline, thread = one row =====================
        const int g = warp >> 2, q = warp & 3;
        const int m = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint64_t* gb_bars = &bars[g * AB_PER_GROUP];
        const uint32_t gb = g_base0 + g * group_bytes, scratch = gb + 2 * half_bytes;
        const uint32_t dcol = tbase + lane_addr + (uint32_t)g * (uint32_t)A.tmem_stride;
        const uint32_t row_off = (m >> 3) * (Dh * 32) + (m & 7) * 16;         // canonical offset of (m, 0) in a half
        const float* prog = A.prog;
        const int fin_pass0 = a_hdr(prog, 1), fin_pass1 = a_hdr(prog, 2);
        const float const_ld = __ldg(prog + 4), const_lp = __ldg(prog + 5);
        const float* fin_params = prog + 8;
        const bool want_lp = A.log_prob != nullptr;
        const bool lp_in = want_lp && (A.flags & B2F_FLOW_LOGP_OF_INPUT);
        const int n_my = tiles_of(g);
        const bool driver = q == 0 && lane == 0;          // the group's thread that talks to the TMA engine and the tensor core
        const uint64_t d1c = umma::make_smem_desc(0, 128, Dh * 32);
        const uint32_t d1_lo = (uint32_t)d1c, d1_hi = (uint32_t)(d1c >> 32);
        const uint32_t w0 = umma::smem_u32(wreg);
        auto is_full = [&](long long t) { return A.use_tma && (t * 128 + 128 <= A.B); };
        auto load_tile = [&](long long t) {               // driver only
            if (is_full(t) && !A.philox) {
                umma::mbar_arrive_expect_tx(&gb_bars[AB_X_FULL], 2 * half_bytes);
                a_tma_load4(gb, &map_x, 0, 0, 0, (int)(t * 16), &gb_bars[AB_X_FULL]);
                a_tma_load4(gb + half_bytes, &map_x, 0, 0, Dh / 4, (int)(t * 16), &gb_bars[AB_X_FULL]);
            } else {
                umma::mbar_arrive(&gb_bars[AB_X_FULL]);   // ragged tile / in-kernel noise: the group fills the tile itself
            }
        };
        if (driver && n_my > 0) load_tile(a_tile(0, g, NG));
        umma::mbar_wait(w_full, 0);
        uint32_t k = 0;                 // (tile, layer) counter of this group: barrier phases
        for (int it = 0; it < n_my; ++it) {
            const long long tile = a_tile(it, g, NG);
            const long long row0 = tile * 128;
            const int rows = (int)min(128LL, A.B - row0);
            const bool full = A.use_tma && rows == 128;
            const bool tma_in = full && !A.philox;
            const bool live = m < rows;
            umma::mbar_wait(&gb_bars[AB_X_FULL], it & 1);
            if (!tma_in) {
                // this thread's row, straight into the canonical tile
                if (A.philox) {
                    const unsigned long long g0 = (unsigned long long)(row0 + m) * (unsigned long long)(D / 4);
#pragma unroll 4
                    for (int kc = 0; kc < D / 4; ++kc) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (live) {
                            v = philox::normal4(g0 + kc, A.seed, A.offset);
                            if (A.base_log_scale) {
                                const float4 ls = __ldg(reinterpret_cast<const float4*>(A.base_log_scale) + kc);
                                v.x *= __expf(ls.x); v.y *= __expf(ls.y); v.z *= __expf(ls.z); v.w *= __expf(ls.w);
                            }
                            if (A.base_loc) {
                                const float4 lc = __ldg(reinterpret_cast<const float4*>(A.base_loc) + kc);
                                v.x += lc.x; v.y += lc.y; v.z += lc.z; v.w += lc.w;
                            }
                        }
                        const uint32_t hf = (4 * kc) >= Dh, k4 = kc - hf * (Dh / 4);
                        a_sts128(gb + hf * half_bytes + row_off + k4 * 128, v);
                    }
                } else {
                    const float4* src = reinterpret_cast<const float4*>(A.x + (row0 + m) * D);
                    for (int kc = 0; kc < D / 4; ++kc) {
                        const float4 v = live ? __ldg(src + kc) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const uint32_t hf = (4 * kc) >= Dh, k4 = kc - hf * (Dh / 4);
                        a_sts128(gb + hf * half_bytes + row_off + k4 * 128, v);
                    }
                }
            }
            float sq_in = 0.0f;
            if (lp_in) {
                // Flow.sample(return_log_prob=True): base density of the INPUT row (flows.py:710-712)
                const float* ip = prog + 8 + 2 * D;
                for (int kc = 0; kc < D / 4; ++kc) {
                    const uint32_t hf = (4 * kc) >= Dh, k4 = kc - hf * (Dh / 4);
                    const float4 v = a_lds128(gb + hf * half_bytes + row_off + k4 * 128);
                    const float4 p0 = __ldg(reinterpret_cast<const float4*>(ip + 8 * kc));
                    const float4 p1 = __ldg(reinterpret_cast<const float4*>(ip + 8 * kc + 4));
                    const float t0 = fmaf(v.x, p0.x, p0.y), t1 = fmaf(v.y, p0.z, p0.w), t2 = fmaf(v.z, p1.x, p1.y), t3 = fmaf(v.w, p1.z, p1.w);
                    sq_in = fmaf(t0, t0, sq_in); sq_in = fmaf(t1, t1, sq_in); sq_in = fmaf(t2, t2, sq_in); sq_in = fmaf(t3, t3, sq_in);
                }
            }
            float ld = 0.0f, sq = 0.0f;
            for (int li = 0; li < L; ++li, ++k) {
                const ALayer& Ly = A.layers[li];
                const float* blob = Ly.blob;
                const uint32_t src_addr = gb + (uint32_t)Ly.src_half * half_bytes + row_off;
                const uint32_t tgt_addr = gb + (uint32_t)(Ly.src_half ^ 1) * half_bytes + row_off;
                const float* b1 = blob + kAHdr + 2 * Ly.N1 * Dh;
                const float* tp = b1 + 32 + 2 * Ly.N2 * Ly.K2;
                const float* sp = tp + Dh * 8;
                const int src_pass = a_hdr(blob, 2);
                // ---- source half: pending elementwise layers, then its tf32 remainder for the 3xTF32 product ----
                const uint32_t lo_addr = scratch + row_off;
                for (int kc = 0; kc < Dh / 4; ++kc) {
                    float4 v = a_lds128(src_addr + kc * 128);
                    if (src_pass) {
                        const float4 p0 = __ldg(reinterpret_cast<const float4*>(sp + 8 * kc));
                        const float4 p1 = __ldg(reinterpret_cast<const float4*>(sp + 8 * kc + 4));
                        v.x = fmaf(v.x, p0.x, p0.y); v.y = fmaf(v.y, p0.z, p0.w); v.z = fmaf(v.z, p1.x, p1.y); v.w = fmaf(v.w, p1.z, p1.w);
                        a_sts128(src_addr + kc * 128, v);
                    }
                    float4 r;
                    r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                    r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                    r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                    r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                    a_sts128(lo_addr + kc * 128, r);
                }
                umma::fence_proxy_async_smem();
                umma::tc_fence_before_sync();
                a_group_sync(g);                         // all 128 rows of both operands are in shared memory
                if (driver) {
                    // D1 = x_hi W1hi^T + x_lo W1hi^T + x_hi W1lo^T   (x_hi = the tile itself, truncated by the tensor core)
                    umma::tc_fence_after_sync();
                    const uint32_t idesc = umma::make_idesc_tf32(128, Ly.N1);
                    const uint32_t xh = d1_lo + ((gb + (uint32_t)Ly.src_half * half_bytes) >> 4), xl = d1_lo + (scratch >> 4);
                    const uint32_t wh = d1_lo + ((w0 + Ly.w_off * 4) >> 4), wl = wh + ((Ly.N1 * Dh * 4) >> 4);
                    const uint32_t dc = dcol - lane_addr;
                    for (int ks = 0; ks < Dh / 8; ++ks) {
                        umma::mma_tf32_ss_parts(dc, xh + ks * 16, d1_hi, wh + ks * 16, d1_hi, idesc, ks > 0);
                        umma::mma_tf32_ss_parts(dc, xl + ks * 16, d1_hi, wh + ks * 16, d1_hi, idesc, 1);
                        umma::mma_tf32_ss_parts(dc, xh + ks * 16, d1_hi, wl + ks * 16, d1_hi, idesc, 1);
                    }
                    umma::mma_commit(&gb_bars[AB_D1_FULL]);
                }
                // ---- hidden layer: D1 + b1 -> tanh -> hi / lo operands of GEMM2 (column H is the 1 that multiplies the bias) ----
                umma::mbar_wait(&gb_bars[AB_D1_FULL], k & 1);
                umma::tc_fence_after_sync();
                {
                    const int K2 = Ly.K2, H = Ly.H;
                    const uint32_t a2h = scratch + (m >> 3) * (K2 * 32) + (m & 7) * 16, a2l = a2h + 128 * K2 * 4;
                    for (int c8 = 0; c8 < K2; c8 += 8) {
                        float v[8];
                        if (c8 < Ly.N1) {
                            umma::tmem_ld8(dcol + c8, v);
                            umma::tmem_ld_wait();
                        }
                        float hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int j = c8 + i;
                            float t = 0.0f;
                            if (j < H) {
                                const float e = a_ex2((v[i] + __ldg(b1 + j)) * (2.0f * kALog2e));
                                t = 1.0f - 2.0f * a_rcp(1.0f + e);               // tanh
                            } else if (j == H) {
                                t = 1.0f;
                            }
                            hi[i] = a_tf32_rn(t);
                            lo[i] = a_tf32_rn(t - hi[i]);
                        }
                        a_sts128(a2h + (c8 >> 2) * 128, make_float4(hi[0], hi[1], hi[2], hi[3]));
                        a_sts128(a2h + (c8 >> 2) * 128 + 128, make_float4(hi[4], hi[5], hi[6], hi[7]));
                        a_sts128(a2l + (c8 >> 2) * 128, make_float4(lo[0], lo[1], lo[2], lo[3]));
                        a_sts128(a2l + (c8 >> 2) * 128 + 128, make_float4(lo[4], lo[5], lo[6], lo[7]));
                    }
                }
                umma::tc_fence_before_sync();
                umma::fence_proxy_async_smem();
                a_group_sync(g);
                if (driver) {
                    umma::tc_fence_after_sync();
                    const uint64_t d2c = umma::make_smem_desc(0, 128, Ly.K2 * 32);
                    const uint32_t d2_lo = (uint32_t)d2c, d2_hi = (uint32_t)(d2c >> 32);
                    const uint32_t idesc = umma::make_idesc_tf32(128, Ly.N2);
                    const uint32_t ah = d2_lo + (scratch >> 4), al = ah + ((128 * Ly.K2 * 4) >> 4);
                    const uint32_t wh = d2_lo + ((w0 + (Ly.w_off + 2 * Ly.N1 * Dh) * 4) >> 4), wl = wh + ((Ly.N2 * Ly.K2 * 4) >> 4);
                    const uint32_t dc = dcol - lane_addr + A.d2_col;
                    for (int ks = 0; ks < Ly.K2 / 8; ++ks) {
                        umma::mma_tf32_ss_parts(dc, ah + ks * 16, d2_hi, wh + ks * 16, d2_hi, idesc, ks > 0);
                        umma::mma_tf32_ss_parts(dc, al + ks * 16, d2_hi, wh + ks * 16, d2_hi, idesc, 1);
                        umma::mma_tf32_ss_parts(dc, ah + ks * 16, d2_hi, wl + ks * 16, d2_hi, idesc, 1);
                    }
                    umma::mma_commit(&gb_bars[AB_D2_FULL]);
                }
                // ---- transformer: this row's Dh elements ----
                umma::mbar_wait(&gb_bars[AB_D2_FULL], k & 1);
                umma::tc_fence_after_sync();
                const uint32_t d2 = dcol + A.d2_col;
                if (Ly.P == 2) {
                    // affine: the elementwise layers that follow are folded into the output layer by the host (_tca.py), the
                    // log-det is +-(c0 + u0 / 2) summed as plain adds (log(e^a + 1e-10) = a to fp32 unless a << 0)
                    const int bits = a_hdr(blob, 7);
                    const bool has_pre = (bits >> 9) & 1, has_fin = ((bits >> 10) & 3) != 0;
                    float usum = 0.0f, fix = 0.0f;
                    for (int e4 = 0; e4 < Dh / 4; ++e4) {                 // 4 elements = 8 parameter columns = one 16-byte tile access
                        float u[8];
                        umma::tmem_ld8(d2 + 8 * e4, u);
                        umma::tmem_ld_wait();
                        const float4 xv = a_lds128(tgt_addr + e4 * 128);
                        float xin[4] = {xv.x, xv.y, xv.z, xv.w};
                        float out[4];
                        if (has_pre) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 pa = __ldg(reinterpret_cast<const float2*>(tp + (4 * e4 + i) * 8));
                                xin[i] = fmaf(xin[i], pa.x, pa.y);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float u0 = u[2 * i], u1 = u[2 * i + 1];
                            const float a2 = fmaf(u0, 0.5f * kALog2e, kAffineC0 * kALog2e);             // log2 of exp(c0 + u0 / 2)
                            const float alpha = a_ex2(a2) + kAffineM;
                            usum += u0;
                            if (a2 < -13.0f) fix += a_lg2(alpha) - a2;                                   // e^a no longer dwarfs 1e-10
                            if (Ly.inverse) {
                                const float pb = __ldg(tp + (4 * e4 + i) * 8 + 3);
                                out[i] = fmaf(xin[i] - u1, a_rcp(alpha), pb);
                            } else {
                                out[i] = fmaf(alpha, xin[i], u1);
                            }
                        }
                        if (has_fin) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 fa = __ldg(reinterpret_cast<const float2*>(tp + (4 * e4 + i) * 8 + 4));
                                const float t = fmaf(out[i], fa.x, fa.y);
                                sq = fmaf(t, t, sq);
                            }
                        }
                        a_sts128(tgt_addr + e4 * 128, make_float4(out[0], out[1], out[2], out[3]));
                    }
                    const float lsum = fmaf(usum, 0.5f, (float)Dh * kAffineC0) + fix * kALn2;
                    ld += Ly.inverse ? -lsum : lsum;
                } else {
                    for (int e8 = 0; e8 < Dh / 8; ++e8) {                 // shift: one parameter per element
                        float u[8];
                        umma::tmem_ld8(d2 + 8 * e8, u);
                        umma::tmem_ld_wait();
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const float4 xv = a_lds128(tgt_addr + (2 * e8 + hh) * 128);
                            const float xin[4] = {xv.x, xv.y, xv.z, xv.w};
                            float out[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int e = 8 * e8 + 4 * hh + i;
                                const float4 pa = __ldg(reinterpret_cast<const float4*>(tp + e * 8));
                                const float2 fa = __ldg(reinterpret_cast<const float2*>(tp + e * 8 + 4));
                                const float v = fmaf(xin[i], pa.x, pa.y);
                                const float o = Ly.inverse ? v - u[4 * hh + i] : v + u[4 * hh + i];
                                const float s = fmaf(o, pa.z, pa.w);
                                out[i] = s;
                                const float t = fmaf(s, fa.x, fa.y);
                                sq = fmaf(t, t, sq);
                            }
                            a_sts128(tgt_addr + (2 * e8 + hh) * 128, make_float4(out[0], out[1], out[2], out[3]));
                        }
                    }
                }
                umma::tc_fence_before_sync();
            }
            // ---- outputs of this row ----
            if (live) {
                const float ldt = ld + const_ld;
                if (A.log_det) A.log_det[row0 + m] = ldt;
                if (want_lp) A.log_prob[row0 + m] = fmaf(-0.5f, lp_in ? sq_in : sq, const_lp) + ldt;
            }
            if (A.y) {
                // elementwise layers still pending at the end of the program
                for (int hf = 0; hf < 2; ++hf) {
                    if (!(hf ? fin_pass1 : fin_pass0)) continue;
                    const float* fp = fin_params + hf * Dh * 2;
                    for (int kc = 0; kc < Dh / 4; ++kc) {
                        float4 v = a_lds128(gb + hf * half_bytes + row_off + kc * 128);
                        const float4 p0 = __ldg(reinterpret_cast<const float4*>(fp + 8 * kc));
                        const float4 p1 = __ldg(reinterpret_cast<const float4*>(fp + 8 * kc + 4));
                        v.x = fmaf(v.x, p0.x, p0.y); v.y = fmaf(v.y, p0.z, p0.w); v.z = fmaf(v.z, p1.x, p1.y); v.w = fmaf(v.w, p1.z, p1.w);
                        a_sts128(gb + hf * half_bytes + row_off + kc * 128, v);
                    }
                }
                if (!full && live) {
                    float4* dst = reinterpret_cast<float4*>(A.y + (row0 + m) * D);
                    for (int kc = 0; kc < D / 4; ++kc) {
                        const uint32_t hf = (4 * kc) >= Dh, k4 = kc - hf * (Dh / 4);
                        dst[kc] = a_lds128(gb + hf * half_bytes + row_off + k4 * 128);
                    }
                }
            }
            umma::fence_proxy_async_smem();
            a_group_sync(g);                             // the whole tile is final
            if (driver) {
                if (A.y && full) {
                    a_tma_store4(&map_y, 0, 0, 0, (int)(tile * 16), gb);
                    a_tma_store4(&map_y, 0, 0, Dh / 4, (int)(tile * 16), gb + half_bytes);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                if (it + 1 < n_my) load_tile(a_tile(it + 1, g, NG));
            }
        }
        if (driver) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    umma::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, A.tmem_cols);
}

// ---- host side --------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFnA encode_tiled_fn_a() {
    static EncodeTiledFnA fn = []() -> EncodeTiledFnA {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFnA>(p);
    }();
    return fn;
}

// (B, D) fp32 row-major seen as {k % 4, row % 8, k / 4, row / 8}; one box = one half of a 128-row tile in canonical order
static bool make_tile_map_a(CUtensorMap* map, const float* base, long long B, int D) {
    EncodeTiledFnA fn = encode_tiled_fn_a();
    if (!fn || B < 128) return false;
    const cuuint64_t gdim[4] = {4, 8, (cuuint64_t)(D / 4), (cuuint64_t)(B / 8)};
    const cuuint64_t gstride[3] = {(cuuint64_t)D * 4, 16, (cuuint64_t)D * 32};
    const cuuint32_t box[4] = {4, 8, (cuuint32_t)(D / 8), 16};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstride, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Returns 1 if the kernel was launched, 0 if the program is not for this kernel (caller falls through), < 0 on error.
int try_launch_flow_tca(const b2f_op_t* ops, int32_t n_ops, const float* x, float* y, float* log_det, float* log_prob,
                        int64_t B, int32_t D, int32_t flags, void* stream, const TcqNoise* noise) {
    if (getenv("B2F_DISABLE_TCA") || getenv("B2F_DISABLE_TC") || (flags & B2F_FLOW_MODE_PRECISE)) return 0;
    if (D % 32 != 0 || D < 32 || D > 128) return 0;
    if ((!noise && (reinterpret_cast<uintptr_t>(x) & 15)) || (y && (reinterpret_cast<uintptr_t>(y) & 15))) return 0;
    if (noise && ((reinterpret_cast<uintptr_t>(noise->base_loc) | reinterpret_cast<uintptr_t>(noise->base_log_scale)) & 15)) return 0;
    AArgs A;
    memset(&A, 0, sizeof(A));
    const int Dh = D / 2;
    int flip = 0, w_floats = 0, scratch = 128 * Dh, n1max = 16, n2max = 16;
    const float* prog = nullptr;
    for (int i = 0; i < n_ops; ++i) {
        const b2f_op_t& o = ops[i];
        if (o.kind == B2F_OP_FLIP) { flip ^= 1; continue; }
        if (o.kind == B2F_OP_ELEMENTWISE && (o.flags & B2F_FLAG_ROW_BIAS)) return 0;      // per-row parameters: generic kernel
        if (o.kind == B2F_OP_ELEMENTWISE) continue;                 // folded into the blobs by the caller
        if (o.kind != B2F_OP_COUPLING || !(o.flags & B2F_FLAG_TCA_OPERANDS)) return 0;
        if (o.flags & B2F_FLAG_ROW_BIAS) return 0;
        const bool affine = o.tkind == B2F_T_AFFINE_FWD || o.tkind == B2F_T_AFFINE_INV;
        const bool shift = o.tkind == B2F_T_SHIFT_ADD || o.tkind == B2F_T_SHIFT_SUB;
        if ((!affine && !shift) || o.n_hidden < 1 || o.n_hidden > 31 || !o.p[4] || A.n_layers >= kAMaxLayers) return 0;
        if (reinterpret_cast<uintptr_t>(o.p[4]) & 15) return fail(B2F_ERR_INVALID, "op %d: tca operand blob must be 16-byte aligned", i);
        if (!prog) prog = (const float*)o.p[5];
        ALayer& L = A.layers[A.n_layers++];
        L.blob = (const float*)o.p[4];
        L.H = o.n_hidden;
        L.P = affine ? 2 : 1;
        L.N1 = (L.H + 15) / 16 * 16;
        L.K2 = (L.H + 1 + 7) / 8 * 8;
        L.N2 = (Dh * L.P + 15) / 16 * 16;
        L.inverse = (o.tkind == B2F_T_AFFINE_INV || o.tkind == B2F_T_SHIFT_SUB);
        L.src_half = flip;
        L.w_off = w_floats;
        w_floats += 2 * L.N1 * Dh + 2 * L.N2 * L.K2;
        scratch = std::max(scratch, 2 * 128 * L.K2);
        n1max = std::max(n1max, L.N1);
        n2max = std::max(n2max, L.N2);
    }
    if (flip != 0 || A.n_layers == 0 || !prog) return 0;
    A.w_floats = (w_floats + 255) / 256 * 256;           // keeps the group regions 1024-byte aligned
    A.scratch_floats = (scratch + 255) / 256 * 256;
    // tensor memory: per group D1 at column 0, D2 at d2_col
    A.d2_col = n1max;
    A.tmem_stride = (n1max + n2max + 31) / 32 * 32;
    const size_t group_bytes = (size_t)2 * 128 * Dh * 4 + (size_t)A.scratch_floats * 4;
    const size_t fixed = (size_t)A.w_floats * 4 + (kAMaxGroups * AB_PER_GROUP + 1) * 8 + 16;
    int ng = kAMaxGroups;
    while (ng > 1 && (fixed + ng * group_bytes > 227 * 1024 || ng * A.tmem_stride > 512)) --ng;
    if (fixed + ng * group_bytes > 227 * 1024 || ng * A.tmem_stride > 512) return 0;
    int cols = 32;
    while (cols < ng * A.tmem_stride) cols <<= 1;
    A.tmem_cols = cols;
    // fewer than four tile pipelines per SM do not hide the tensor-core round trips: the row-per-thread kernel is faster
    // there (RealNVP-128, 2^19 rows: 0.41 ms against 0.50 ms); B2F_TCA_ANY=1 (tests) takes every shape that fits
    if (ng < kAMaxGroups && !getenv("B2F_TCA_ANY")) return 0;
    A.n_groups = ng;
    const size_t smem = fixed + ng * group_bytes;
    A.D = D; A.flags = flags; A.B = B;
    A.n_tiles = (int)((B + 127) / 128);
    A.x = x; A.y = y; A.log_det = log_det; A.log_prob = log_prob; A.prog = prog;
    CUtensorMap map_x, map_y;
    memset(&map_x, 0, sizeof(map_x));
    memset(&map_y, 0, sizeof(map_y));
    A.use_tma = getenv("B2F_TCA_NO_TMA") ? 0 : 1;
    if (noise) {
        A.philox = 1; A.seed = noise->seed; A.offset = noise->offset;
        A.base_loc = noise->base_loc; A.base_log_scale = noise->base_log_scale;
    }
    if (A.use_tma && !noise && !make_tile_map_a(&map_x, x, B, D)) A.use_tma = 0;
    if (A.use_tma && y && !make_tile_map_a(&map_y, y, B, D)) A.use_tma = 0;
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int grid = std::max(1, std::min((A.n_tiles + ng - 1) / ng, n_sm));
    cudaError_t ce = (cudaError_t)raise_smem_limit((const void*)flow_tca_kernel, smem);
    if (ce != cudaSuccess) return fail(B2F_ERR_CUDA, "cudaFuncSetAttribute(tca): %s", cudaGetErrorString(ce));
    flow_tca_kernel<<<grid, kAThreads, smem, (cudaStream_t)stream>>>(A, map_x, map_y);
    const int rc = check_launch("b2f_flow_apply (affine tensor-core kernel)");
    return rc == B2F_OK ? 1 : rc;
}

}  // namespace b2f
