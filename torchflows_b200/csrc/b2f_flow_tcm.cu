// Masked-autoregressive spline flows, one-pass direction ("tcm"): MaskedAutoregressiveRQNSF densities /
// InverseAutoregressiveRQNSF sampling passes (ElementwiseAffine / ActNorm, ReversePermutation, MADE-conditioned
// rational-quadratic layers) in ONE persistent launch, the MADE conditioner on tcgen05 with its masks folded into the
// weight tiles, the sample tile moved by the TMA engine.
//
// Replaces (file:line relative to /root/reference/torchflows): bijections/base.py:203-232, bijections/finite/autoregressive/
// layers_base.py:202-211 (MaskedAutoregressiveBijection, one pass), conditioning/transforms.py:184-266 (MADE: masked
// Linear -> Tanh -> masked Linear), transformers/spline/rational_quadratic.py:45-200, flows.py:628-648.
//
// Same machinery as the spline coupling kernel (b2f_flow_tcq.cu: folded output-layer columns of b2f_rqfast.cuh, four
// independent epilogue groups with double-buffered TMEM accumulators, W2 streamed through a ring, per-column affine maps
// instead of elementwise passes); what differs is the data flow of a MADE layer: ALL D columns feed the conditioner
// (GEMM1 has K = D) and ALL D columns are transformed (D / 2 GEMM2 chunks of two elements).  One pass is well defined in
// place: the hidden activations are complete (GEMM1 has read the whole tile) before the first element is overwritten.
// The tile is therefore one canonical [128 x D] operand, double-buffered as a whole: tile t + 1 is loaded while tile t
// is processed (D <= 128: 2 x 64 KB).
//
// Operand layout: torchflows_b200/_tcm.py (include/b2f.h, B2F_FLAG_TCM_OPERANDS).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "b2f_flow_device.cuh"
#include "b2f_philox.cuh"
#include "b2f_rqfast.cuh"
#include "b2f_umma.cuh"

namespace b2f {
namespace tcm {

constexpr int kEpiWarps = 16;
constexpr int kThreads = (kEpiWarps + 3) * 32;     // + MMA issuer, weight loader, tile IO
constexpr int kN1 = 32;                      // UMMA N of GEMM1 (hidden units, padded)
constexpr int kN2 = 48;                      // UMMA N of a GEMM2 chunk: 2 elements x 24 columns
constexpr int kSlots = 8;                    // TMEM accumulator slots: (group, buffer)
constexpr int kRing = 3;
constexpr int kColD1 = 384;
constexpr int kTmemCols = 512;
constexpr int kHdr = 8;
constexpr int kMaxLayers = 12;

struct MLayer {
    const float* blob;
    int H, K2, inverse;
    float boundary;
};

struct MArgs {
    MLayer layers[kMaxLayers];
    int n_layers, D, flags, n_tiles, use_tma, n_ring;
    long long B;
    const float* x;
    float* y;
    float* log_det;
    float* log_prob;
    const float* prog;      // program blob: [4 ints][4 consts][D x (fin_a, fin_b)][D x (in_a, in_b)]
    int philox;
    unsigned long long seed, offset;
    const float* base_loc;
    const float* base_log_scale;
};

enum { MB_X_FULL = 0, MB_TILE_DONE = 2, MB_W1_FULL, MB_W1_EMPTY, MB_A1_READY, MB_D1_FULL, MB_A2_FULL,
       MB_W2_FULL, MB_W2_EMPTY = MB_W2_FULL + kRing, MB_D2_FULL = MB_W2_EMPTY + kRing,
       MB_D2_EMPTY = MB_D2_FULL + kSlots, MB_COUNT = MB_D2_EMPTY + kSlots };

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }
__device__ __forceinline__ float tf32_rn(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ void tma_load4(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store4(const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(src) : "memory");
}
__device__ __forceinline__ int hdr(const float* blob, int i) { return __ldg(reinterpret_cast<const int*>(blob) + i); }
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 affine4(float4 v, float4 p0, float4 p1) {
    v.x = fmaf(v.x, p0.x, p0.y); v.y = fmaf(v.y, p0.z, p0.w);
    v.z = fmaf(v.z, p1.x, p1.y); v.w = fmaf(v.w, p1.z, p1.w);
    return v;
}

// Cooperative pass of the 16 epilogue warps over the tile: x <- a * x + b per column (params: [D][2]).
// Warp w owns rows 8w..8w+7, four lanes share a row and stride over its 16-byte column groups.
__device__ __forceinline__ void affine_pass(uint32_t tile_addr, uint32_t row_off, const float* __restrict__ params, int D, int kq) {
#pragma unroll 4
    for (int kc = kq; kc < D / 4; kc += 4) {
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(params + 8 * kc));
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(params + 8 * kc + 4));
        const uint32_t a = tile_addr + row_off + kc * 128;
        sts128(a, affine4(lds128(a), p0, p1));
    }
}

// One MADE layer's transformer phase for one epilogue thread (row m, group g): its chunks c = g, g + 4, ... over ALL columns
template <bool INV, bool SAFE>
__device__ __forceinline__ void chunk_loop(uint64_t* bars, const float* __restrict__ tp, float boundary, uint32_t tbase,
                                           uint32_t lane_addr, uint32_t tile_addr, int D, int n_chunks, uint32_t cc_base, int m,
                                           int g, int lane, float& ld2, float& sq) {
    const uint32_t xrow = tile_addr + (m >> 3) * (D * 32) + (m & 7) * 16;       // canon_off(m, 0, D)
    for (int c = g; c < n_chunks; c += 4) {
        const uint32_t ccl = cc_base + c, slot = ccl & (kSlots - 1);
        umma::mbar_wait(&bars[MB_D2_FULL + slot], (ccl >> 3) & 1);
        umma::tc_fence_after_sync();
        const uint32_t tcol = tbase + lane_addr + g * (2 * kN2) + ((c >> 2) & 1) * kN2;
        float ga[24], gb[24];
        umma::tmem_ld8_nowait<0>(tcol, ga);
        umma::tmem_ld8_nowait<8>(tcol + 8, ga);
        umma::tmem_ld8_nowait<16>(tcol + 16, ga);
        umma::tmem_ld8_nowait<0>(tcol + 24, gb);
        umma::tmem_ld8_nowait<8>(tcol + 32, gb);
        umma::tmem_ld8_nowait<16>(tcol + 40, gb);
        umma::tmem_ld_wait();
        umma::tc_fence_before_sync();
        if (lane == 0) umma::mbar_arrive(&bars[MB_D2_EMPTY + slot]);
        const int e0 = 2 * c;
        const uint32_t px = xrow + (e0 >> 2) * 128 + (e0 & 3) * 4;
        const float2 xv = lds64(px);
        const float4 pa = __ldg(reinterpret_cast<const float4*>(tp + e0 * 8));          // pre_a, pre_b, post_a, post_b
        const float2 fa = __ldg(reinterpret_cast<const float2*>(tp + e0 * 8 + 4));      // fin_a, fin_b
        const float4 pb = __ldg(reinterpret_cast<const float4*>(tp + e0 * 8 + 8));
        const float2 fb = __ldg(reinterpret_cast<const float2*>(tp + e0 * 8 + 12));
        float oa, ob, la, lb;
        const float va = fmaf(xv.x, pa.x, pa.y), vb = fmaf(xv.y, pb.x, pb.y);
        if (INV) {
            rqf::inverse<SAFE, 0>(va, ga, boundary, oa, la);
            rqf::inverse<SAFE, 0>(vb, gb, boundary, ob, lb);
        } else {
            rqf::forward<SAFE, 0>(va, ga, boundary, oa, la);
            rqf::forward<SAFE, 0>(vb, gb, boundary, ob, lb);
        }
        const float sa = fmaf(oa, pa.z, pa.w), sb = fmaf(ob, pb.z, pb.w);
        sts64(px, make_float2(sa, sb));
        const float ta = fmaf(sa, fa.x, fa.y), tb = fmaf(sb, fb.x, fb.y);
        sq = fmaf(ta, ta, sq);
        sq = fmaf(tb, tb, sq);
        ld2 += la + lb;
    }
}

__global__ void __launch_bounds__(kThreads, 1)
flow_tcm_kernel(const __grid_constant__ MArgs A, const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int D = A.D;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    int K2max = 8;
    for (int i = 0; i < A.n_layers; ++i) K2max = max(K2max, A.layers[i].K2);
    const uint32_t tile_bytes = 128u * D * 4;
    uint8_t* p = smem_raw;
    const uint32_t xt0 = umma::smem_u32(p); p += 2 * tile_bytes;        // tile buffer b at xt0 + b * tile_bytes
    float* w1 = reinterpret_cast<float*>(p); p += kN1 * D * 4;
    float* a2 = reinterpret_cast<float*>(p); p += 128 * K2max * 4;
    float* w2 = reinterpret_cast<float*>(p); p += A.n_ring * 4 * kN2 * K2max * 4;
    const int w2stride = 4 * kN2 * K2max;
    float2* red = reinterpret_cast<float2*>(p); p += 4 * 128 * 8;
    float* lpin = reinterpret_cast<float*>(p); p += 128 * 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(p); p += MB_COUNT * 8;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(p);

    if (tid == 0) {
        umma::mbar_init(&bars[MB_X_FULL], 1);
        umma::mbar_init(&bars[MB_X_FULL + 1], 1);
        umma::mbar_init(&bars[MB_TILE_DONE], kEpiWarps);
        umma::mbar_init(&bars[MB_W1_FULL], 1);
        umma::mbar_init(&bars[MB_W1_EMPTY], 1);
        umma::mbar_init(&bars[MB_A1_READY], kEpiWarps);
        umma::mbar_init(&bars[MB_D1_FULL], 1);
        umma::mbar_init(&bars[MB_A2_FULL], kEpiWarps);
        for (int i = 0; i < kRing; ++i) {
            umma::mbar_init(&bars[MB_W2_FULL + i], 1);
            umma::mbar_init(&bars[MB_W2_EMPTY + i], 1);
        }
        for (int i = 0; i < kSlots; ++i) {
            umma::mbar_init(&bars[MB_D2_FULL + i], 1);
            umma::mbar_init(&bars[MB_D2_EMPTY + i], 4);
        }
        umma::fence_barrier_init();
    }
    if (warp == kEpiWarps) umma::tmem_alloc(tmem_ptr, kTmemCols);
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tc_fence_after_sync();
    const uint32_t tbase = *tmem_ptr;
    const int n_chunks = D >> 1;
    const bool want_lp = A.log_prob != nullptr;
    const bool lp_in = want_lp && (A.flags & B2F_FLOW_LOGP_OF_INPUT);

    uint32_t lc = 0, cc = 0, tc = 0;          // layers / GEMM2 chunks / tiles processed so far (all roles count identically)
    uint32_t ring = 0, ring_ph = 0;

    if (warp == kEpiWarps + 2) {
        // ===================== tile IO (one thread): whole tiles, one ahead =====================
        if (lane == 0) {
            auto is_full = [&](int t) { return A.use_tma && ((long long)t * 128 + 128 <= A.B); };
            auto load = [&](int t, uint32_t buf) {
                uint64_t* bar = &bars[MB_X_FULL + buf];
                if (is_full(t) && !A.philox) {
                    umma::mbar_arrive_expect_tx(bar, tile_bytes);
                    tma_load4(xt0 + buf * tile_bytes, &map_x, 0, 0, 0, t * 16, bar);
                } else {
                    umma::mbar_arrive(bar);          // ragged tile / in-kernel noise: the epilogue warps fill it themselves
                }
            };
            int tile = blockIdx.x;
            if (tile < A.n_tiles) load(tile, 0);
            if (tile + (int)gridDim.x < A.n_tiles) load(tile + gridDim.x, 1);
            for (; tile < A.n_tiles; tile += gridDim.x, ++tc) {
                const uint32_t buf = tc & 1;
                umma::mbar_wait_backoff(&bars[MB_TILE_DONE], tc & 1);
                if (A.y && is_full(tile)) {
                    tma_store4(&map_y, 0, 0, 0, tile * 16, xt0 + buf * tile_bytes);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                const int next2 = tile + 2 * (int)gridDim.x;
                if (next2 < A.n_tiles) load(next2, buf);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        __syncwarp();
    } else if (warp == kEpiWarps + 1) {
        // ===================== weight loader (one thread) =====================
        if (lane == 0) {
            for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
                for (int li = 0; li < A.n_layers; ++li, ++lc) {
                    const MLayer& L = A.layers[li];
                    umma::mbar_wait(&bars[MB_W1_EMPTY], (lc & 1) ^ 1);
                    const uint32_t w1_bytes = kN1 * D * 4;
                    umma::mbar_arrive_expect_tx(&bars[MB_W1_FULL], w1_bytes);
                    umma::bulk_g2s(w1, L.blob + kHdr, w1_bytes, &bars[MB_W1_FULL]);
                    const float* w2g = L.blob + kHdr + kN1 * D + 32;
                    const uint32_t rd_bytes = 4 * kN2 * L.K2 * 4;
                    for (int r = 0; r < n_chunks / 4; ++r) {
                        umma::mbar_wait(&bars[MB_W2_EMPTY + ring], ring_ph ^ 1);
                        umma::mbar_arrive_expect_tx(&bars[MB_W2_FULL + ring], rd_bytes);
                        umma::bulk_g2s(w2 + ring * w2stride, w2g + (size_t)r * 4 * kN2 * L.K2, rd_bytes, &bars[MB_W2_FULL + ring]);
                        if (++ring == (uint32_t)A.n_ring) { ring = 0; ring_ph ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == kEpiWarps) {
        // ===================== MMA issuer =====================
        const uint32_t leader = umma::elect_one();
        const uint32_t idesc1 = umma::make_idesc_tf32(128, kN1), idesc2 = umma::make_idesc_tf32(128, kN2);
        const uint32_t w1a = umma::smem_u32(w1) >> 4, a2a = umma::smem_u32(a2) >> 4;
        const uint64_t d1c = umma::make_smem_desc(0, 128, D * 32);
        const uint32_t d1_lo = (uint32_t)d1c, d1_hi = (uint32_t)(d1c >> 32);
        for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++tc) {
            const uint32_t buf = tc & 1;
            for (int li = 0; li < A.n_layers; ++li, ++lc) {
                const MLayer& L = A.layers[li];
                const uint32_t ph = lc & 1;
                umma::mbar_wait(&bars[MB_W1_FULL], ph);
                umma::mbar_wait(&bars[MB_A1_READY], ph);
                umma::tc_fence_after_sync();
                if (leader) {
                    // GEMM1: D1[128 x 32] = x[128 x D] * (W1 * mask)[32 x D]^T
                    const uint32_t xa = d1_lo + ((xt0 + buf * tile_bytes) >> 4), wa = d1_lo + w1a;
                    for (int ks = 0; ks < D / 8; ++ks)
                        umma::mma_tf32_ss_parts(tbase + kColD1, xa + ks * 16, d1_hi, wa + ks * 16, d1_hi, idesc1, ks > 0);
                    umma::mma_commit(&bars[MB_D1_FULL]);
                    umma::mma_commit(&bars[MB_W1_EMPTY]);
                }
                __syncwarp();
                umma::mbar_wait(&bars[MB_A2_FULL], ph);
                umma::tc_fence_after_sync();
                const uint64_t d2c = umma::make_smem_desc(0, 128, L.K2 * 32);
                const uint32_t d2_lo = (uint32_t)d2c, d2_hi = (uint32_t)(d2c >> 32);
                const uint32_t a_lo = d2_lo + a2a;
                const int nk = L.K2 / 8;
                const uint32_t grp_units = (kN2 / 8) * (L.K2 * 32) / 16;
                for (int r = 0; r < n_chunks / 4; ++r, cc += 4) {
                    const uint32_t b = (cc >> 2) & 1, ph2 = (cc >> 3) & 1;
                    umma::mbar_wait(&bars[MB_W2_FULL + ring], ring_ph);
                    const uint32_t w_lo = d2_lo + (umma::smem_u32(w2 + ring * w2stride) >> 4);
#pragma unroll
                    for (uint32_t g = 0; g < 4; ++g) {
                        const uint32_t slot = g + 4 * b;
                        umma::mbar_wait(&bars[MB_D2_EMPTY + slot], ph2 ^ 1);
                        umma::tc_fence_after_sync();
                        if (leader) {
                            const uint32_t dcol = tbase + g * (2 * kN2) + b * kN2;
                            const uint32_t wg = w_lo + g * grp_units;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                if (ks < nk) umma::mma_tf32_ss_parts(dcol, a_lo + ks * 16, d2_hi, wg + ks * 16, d2_hi, idesc2, ks > 0);
                            umma::mma_commit(&bars[MB_D2_FULL + slot]);
                        }
                    }
                    if (leader) umma::mma_commit(&bars[MB_W2_EMPTY + ring]);
                    __syncwarp();
                    if (++ring == (uint32_t)A.n_ring) { ring = 0; ring_ph ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int q = warp & 3, g = warp >> 2;
        const int m = q * 32 + lane;
        const int m8 = warp * 8 + (lane & 7), kq = lane >> 3;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const float* prog = A.prog;
        const int fin_pass = hdr(prog, 1);
        const float const_ld = __ldg(prog + 4), const_lp = __ldg(prog + 5);
        const float* fin_params = prog + 8;
        const uint32_t row_off = (m8 >> 3) * (D * 32) + (m8 & 7) * 16;     // canon_off(m8, 0, D)
        for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++tc) {
            const long long row0 = (long long)tile * 128;
            const int rows = (int)min(128LL, A.B - row0);
            const bool full = A.use_tma && rows == 128;
            const bool tma_in = full && !A.philox;
            const uint32_t buf = tc & 1, tile_addr = xt0 + buf * tile_bytes;
            umma::mbar_wait(&bars[MB_X_FULL + buf], (tc >> 1) & 1);
            if (!tma_in) {
                const bool live = m8 < rows;
                if (A.philox) {
                    const unsigned long long g0 = (unsigned long long)(row0 + m8) * (unsigned long long)(D / 4);
#pragma unroll 4
                    for (int kc = kq; kc < D / 4; kc += 4) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (live) {
                            v = philox::normal4(g0 + kc, A.seed, A.offset);
                            if (A.base_log_scale) {
                                const float4 ls = __ldg(reinterpret_cast<const float4*>(A.base_log_scale) + kc);
                                v.x *= __expf(ls.x); v.y *= __expf(ls.y); v.z *= __expf(ls.z); v.w *= __expf(ls.w);
                            }
                            if (A.base_loc) {
                                const float4 lc4 = __ldg(reinterpret_cast<const float4*>(A.base_loc) + kc);
                                v.x += lc4.x; v.y += lc4.y; v.z += lc4.z; v.w += lc4.w;
                            }
                        }
                        sts128(tile_addr + row_off + kc * 128, v);
                    }
                } else {
                    const float4* src = reinterpret_cast<const float4*>(A.x + (row0 + m8) * D);
                    for (int kc = kq; kc < D / 4; kc += 4)
                        sts128(tile_addr + row_off + kc * 128, live ? __ldg(src + kc) : make_float4(0.f, 0.f, 0.f, 0.f));
                }
                epi_sync();
            }
            if (lp_in) {
                // Flow.sample(return_log_prob=True): base density of the INPUT rows (flows.py:710-712)
                const float* ip = prog + 8 + 2 * D;
                float acc = 0.0f;
                for (int kc = kq; kc < D / 4; kc += 4) {
                    const float4 v = lds128(tile_addr + row_off + kc * 128);
                    const float4 p0 = __ldg(reinterpret_cast<const float4*>(ip + 8 * kc));
                    const float4 p1 = __ldg(reinterpret_cast<const float4*>(ip + 8 * kc + 4));
                    const float4 t = affine4(v, p0, p1);
                    acc = fmaf(t.x, t.x, acc); acc = fmaf(t.y, t.y, acc); acc = fmaf(t.z, t.z, acc); acc = fmaf(t.w, t.w, acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 8);
                acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                if (kq == 0) lpin[m8] = acc;
            }
            float ld2 = 0.0f, sq = 0.0f;
            for (int li = 0; li < A.n_layers; ++li, ++lc) {
                const MLayer& L = A.layers[li];
                const float* blob = L.blob;
                const int src_pass = hdr(blob, 2);
                const int H = L.H, K2 = L.K2;
                const float* b1 = blob + kHdr + kN1 * D;
                const float* tp = b1 + 32 + (size_t)n_chunks * kN2 * K2;
                const float* sp = tp + D * 8;
                const float* misc = sp + D * 2;
                if (src_pass) {
                    if (li > 0) epi_sync();           // the pass touches columns other warps wrote in the previous layer
                    // elementwise layers in front of this layer: every column feeds the conditioner, so they are materialised
                    affine_pass(tile_addr, row_off, sp, D, kq);
                }
                umma::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&bars[MB_A1_READY]);
                // hidden layer: D1 -> + b1 -> tanh -> tf32 -> a2; columns H, H+1 = the 1 that multiplies the bias (hi, lo)
                umma::mbar_wait(&bars[MB_D1_FULL], lc & 1);
                umma::tc_fence_after_sync();
                if (8 * g < K2) {
                    float v[8];
                    umma::tmem_ld8(tbase + lane_addr + kColD1 + 8 * g, v);
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int j = 8 * g + i;
                        if (j < H) {
                            const float e = rqf::f_ex2((v[i] + __ldg(b1 + j)) * (2.0f * rqf::kLog2e));
                            v[i] = tf32_rn(1.0f - 2.0f * rqf::f_rcp(1.0f + e));       // tanh
                        } else {
                            v[i] = (j < H + 2) ? 1.0f : 0.0f;
                        }
                    }
                    const uint32_t a2row = umma::smem_u32(a2) + (m >> 3) * (K2 * 32) + (m & 7) * 16 + (2 * g) * 128;
                    sts128(a2row, make_float4(v[0], v[1], v[2], v[3]));
                    sts128(a2row + 128, make_float4(v[4], v[5], v[6], v[7]));
                }
                umma::tc_fence_before_sync();
                umma::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&bars[MB_A2_FULL]);
                // transformer phase (GEMM1 has read the whole tile: D1_FULL was observed above)
                const bool fast = __ldg(misc) < rqf::kNoMaxBound && __ldg(misc + 1) < rqf::kPolyBound;
                if (L.inverse) {
                    if (fast) chunk_loop<true, false>(bars, tp, L.boundary, tbase, lane_addr, tile_addr, D, n_chunks, cc, m, g, lane, ld2, sq);
                    else chunk_loop<true, true>(bars, tp, L.boundary, tbase, lane_addr, tile_addr, D, n_chunks, cc, m, g, lane, ld2, sq);
                } else {
                    if (fast) chunk_loop<false, false>(bars, tp, L.boundary, tbase, lane_addr, tile_addr, D, n_chunks, cc, m, g, lane, ld2, sq);
                    else chunk_loop<false, true>(bars, tp, L.boundary, tbase, lane_addr, tile_addr, D, n_chunks, cc, m, g, lane, ld2, sq);
                }
                cc += n_chunks;
            }
            // ---- outputs of this tile ----
            red[g * 128 + m] = make_float2(ld2, sq);
            epi_sync();
            if (tid < 128) {
                const float2 r0 = red[tid], r1 = red[128 + tid], r2 = red[256 + tid], r3 = red[384 + tid];
                const float ld = fmaf((r0.x + r1.x) + (r2.x + r3.x), rqf::kLn2, const_ld);
                float sqs = (r0.y + r1.y) + (r2.y + r3.y);
                if (lp_in) sqs = lpin[tid];
                if (tid < rows) {
                    if (A.log_det) A.log_det[row0 + tid] = ld;
                    if (want_lp) A.log_prob[row0 + tid] = fmaf(-0.5f, sqs, const_lp) + ld;
                }
            }
            if (A.y) {
                if (fin_pass) affine_pass(tile_addr, row_off, fin_params, D, kq);
                if (!full) {
                    epi_sync();
                    if (m8 < rows) {
                        float4* dst = reinterpret_cast<float4*>(A.y + (row0 + m8) * D);
                        for (int kc = kq; kc < D / 4; kc += 4) dst[kc] = lds128(tile_addr + row_off + kc * 128);
                    }
                }
            }
            umma::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bars[MB_TILE_DONE]);
        }
    }
    umma::tc_fence_before_sync();
    __syncthreads();
    if (warp == kEpiWarps) umma::tmem_dealloc(tbase, kTmemCols);
}

// ---- host side --------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// (B, D) fp32 row-major seen as {k % 4, row % 8, k / 4, row / 8}; one box = one whole 128-row tile in canonical order
static bool make_tile_map(CUtensorMap* map, const float* base, long long B, int D) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || B < 128) return false;
    const cuuint64_t gdim[4] = {4, 8, (cuuint64_t)(D / 4), (cuuint64_t)(B / 8)};
    const cuuint64_t gstride[3] = {(cuuint64_t)D * 4, 16, (cuuint64_t)D * 32};
    const cuuint32_t box[4] = {4, 8, (cuuint32_t)(D / 4), 16};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstride, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tcm

// Returns 1 if the kernel was launched, 0 if the program is not for this kernel (caller falls through), < 0 on error.
int try_launch_flow_tcm(const b2f_op_t* ops, int32_t n_ops, const float* x, float* y, float* log_det, float* log_prob,
                        int64_t B, int32_t D, int32_t flags, void* stream, const TcqNoise* noise) {
    using namespace tcm;
    if (getenv("B2F_DISABLE_TCM") || getenv("B2F_DISABLE_TC") || (flags & B2F_FLOW_MODE_PRECISE)) return 0;
    if (D % 32 != 0 || D < 32 || D > 128) return 0;
    if ((!noise && (reinterpret_cast<uintptr_t>(x) & 15)) || (y && (reinterpret_cast<uintptr_t>(y) & 15))) return 0;
    if (noise && ((reinterpret_cast<uintptr_t>(noise->base_loc) | reinterpret_cast<uintptr_t>(noise->base_log_scale)) & 15)) return 0;
    MArgs A;
    memset(&A, 0, sizeof(A));
    int flip = 0, K2max = 8;
    const float* prog = nullptr;
    for (int i = 0; i < n_ops; ++i) {
        const b2f_op_t& o = ops[i];
        if (o.kind == B2F_OP_FLIP) { flip ^= 1; continue; }
        if (o.kind == B2F_OP_ELEMENTWISE && (o.flags & B2F_FLAG_ROW_BIAS)) return 0;      // per-row parameters: generic kernel
        if (o.kind == B2F_OP_ELEMENTWISE) continue;                 // folded into the blobs by the caller
        if (o.kind != B2F_OP_MADE || !(o.flags & B2F_FLAG_TCM_OPERANDS)) return 0;
        if (o.flags & B2F_FLAG_ROW_BIAS) return 0;
        if ((o.tkind != B2F_T_RQ_FWD && o.tkind != B2F_T_RQ_INV) || o.n_bins != 8) return 0;
        if (o.n_hidden < 1 || o.n_hidden > 30 || !o.p[4] || A.n_layers >= kMaxLayers) return 0;
        if (reinterpret_cast<uintptr_t>(o.p[4]) & 15) return fail(B2F_ERR_INVALID, "op %d: tcm operand blob must be 16-byte aligned", i);
        if (!prog) prog = (const float*)o.p[5];
        MLayer& L = A.layers[A.n_layers++];
        L.blob = (const float*)o.p[4];
        L.H = o.n_hidden;
        L.K2 = (o.n_hidden + 2 + 7) / 8 * 8;
        L.inverse = o.tkind == B2F_T_RQ_INV;
        L.boundary = o.boundary;
        if (!(o.boundary > 0.0f)) return fail(B2F_ERR_INVALID, "op %d: boundary", i);
        K2max = std::max(K2max, L.K2);
    }
    if (flip != 0 || A.n_layers == 0 || !prog) return 0;
    auto smem_bytes = [&](int n_ring) {
        return (size_t)2 * 128 * D * 4 + (size_t)kN1 * D * 4 + (size_t)128 * K2max * 4 +
               (size_t)n_ring * 4 * kN2 * K2max * 4 + 4 * 128 * 8 + 128 * 4 + MB_COUNT * 8 + 16;
    };
    A.n_ring = kRing;
    while (A.n_ring > 2 && smem_bytes(A.n_ring) > 227 * 1024) --A.n_ring;
    const size_t smem = smem_bytes(A.n_ring);
    if (smem > 227 * 1024) return 0;
    A.D = D; A.flags = flags; A.B = B;
    A.n_tiles = (int)((B + 127) / 128);
    A.x = x; A.y = y; A.log_det = log_det; A.log_prob = log_prob; A.prog = prog;
    CUtensorMap map_x, map_y;
    memset(&map_x, 0, sizeof(map_x));
    memset(&map_y, 0, sizeof(map_y));
    A.use_tma = getenv("B2F_TCM_NO_TMA") ? 0 : 1;
    if (noise) {
        A.philox = 1; A.seed = noise->seed; A.offset = noise->offset;
        A.base_loc = noise->base_loc; A.base_log_scale = noise->base_log_scale;
    }
    if (A.use_tma && !noise && !make_tile_map(&map_x, x, B, D)) A.use_tma = 0;
    if (A.use_tma && y && !make_tile_map(&map_y, y, B, D)) A.use_tma = 0;
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int grid = std::min(A.n_tiles, n_sm);
    cudaError_t ce = (cudaError_t)raise_smem_limit((const void*)flow_tcm_kernel, smem);
    if (ce != cudaSuccess) return fail(B2F_ERR_CUDA, "cudaFuncSetAttribute(tcm): %s", cudaGetErrorString(ce));
    flow_tcm_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(A, map_x, map_y);
    const int rc = check_launch("b2f_flow_apply (MADE spline tensor-core kernel)");
    return rc == B2F_OK ? 1 : rc;
}

}  // namespace b2f
