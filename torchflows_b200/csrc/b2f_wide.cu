// Wide-conditioner spline coupling layer (CouplingRQNSF with n_hidden in the hundreds or thousands, e.g. the
// n_dim = 1024 / n_hidden = 1024 data-parallel training configuration): every dense contraction of the layer -- forward,
// recompute, dgrad and wgrad -- on tcgen05 (kind::tf32, fp32 accumulation in tensor memory), fed by the TMA engine.
//
// Replaces, for one layer (file:line relative to /root/reference/torchflows/bijections/finite/autoregressive):
//   layers_base.py:119-163            CouplingBijection.forward / inverse
//   conditioning/transforms.py:274-307   FeedForward: Linear(Dh, H) -> Tanh -> Linear(H, Dh * 23)
//   transformers/spline/rational_quadratic.py:45-200  the spline applied to the target half, and their autograd backward.
//
// The conditioner output h (Dh * 23 floats per sample: 47 KB at Dh = 512) is never written to HBM: the output-layer GEMM
// keeps a [256 samples x 8 elements x 24] accumulator in tensor memory and the spline (or its backward) is the GEMM's
// epilogue.  Backward recomputes h the same way; what the backward does write is dL/dh, once, because it feeds two
// contractions with different reduction axes (over the batch for dL/dW2, over the parameters for dL/dhid) whose
// accumulators ([12288 x 1024] and [B x 1024] fp32) cannot both stay on chip.
//
// Operand format ("tiled", T(R, K)): a [R x K] matrix whose K axis is the contraction axis is stored k-block-major,
//     offset(r, k) = ((k / 32) * (R / 8) + r / 8) * 256 + ((k % 32) / 4) * 32 + (r % 8) * 4 + k % 4      (floats)
// so that any run of rows (a multiple of 8) of one 32-wide k-block is ONE contiguous chunk already in the UMMA canonical
// K-major no-swizzle order (8 x 16-byte core matrices, LBO = 128 B, SBO = 1024 B): one cp.async.bulk per operand per
// pipeline stage, no tensor map.  Producers (the pack kernel and the GEMM epilogues) write this format directly.
//
// One GEMM kernel (wide_gemm_kernel): persistent, 1 CTA / SM, 18 warps = 16 epilogue + MMA issuer + loader.  A work
// item is a [256 x NT] output block (two M = 128 UMMA tiles sharing every B stage: 32 + NT / 8 KB of operands per 32-wide
// k-block, the L2 -> SM traffic is what bounds a TF32 GEMM on this part) over a k-block range (split-K for the
// gradient GEMMs whose output grid is smaller than the machine).  Epilogues: plain store / atomic add (row-major),
// spline forward / inverse, spline backward.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "b2f_common.cuh"
#include "b2f_math.cuh"
#include "b2f_rqfast.cuh"
#include "b2f_umma.cuh"

namespace b2f {
namespace wide {

constexpr int kEpiWarps = 16;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kABytes = 256 * 128;        // one k-block of the two M tiles
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 4;

enum { EPI_STORE = 0, EPI_SPLINE_FWD, EPI_SPLINE_INV, EPI_SPLINE_BWD_FWD, EPI_SPLINE_BWD_INV };

struct GArgs {
    const float* A;          // T(RA, K)
    const float* Bm;         // T(RB, K)
    int RA8, RB8;            // rows / 8 of the tiled operands
    int RB;                  // real rows of B (multiple of 16): the last N chunk may be narrower than NT
    int NT;                  // N chunk: 192 (spline epilogues: 8 elements x 24 columns) or up to 256
    int n_mt, n_nt, n_split; // items = n_mt * n_nt * n_split (nt fastest, then mt, then the k split)
    int kb_total;            // K / 32
    int stages;
    // EPI_STORE
    float* C;
    long long ldc;
    int M_real, N_real;      // rows / columns actually written
    int remap24;             // output row m -> (m / 24) * 23 + m % 24, rows with m % 24 == 23 dropped (padded parameter slot)
    int atomic;              // red.global.add instead of st.global
    // spline epilogues
    const float* x;          // layer input (B, D); the target half starts at column Dh
    float* y;                // layer output (B, D) (target half written here)
    long long ldx;           // D
    int Dh;
    long long B, Bp;
    const float* b2p;        // output-layer bias padded to 24 per element
    float boundary;
    float* ldp;              // [2 * n_nt][Bp] partial log-determinants (natural log)
    const float* gy;         // backward: dL/dy (B, D) (nullable)
    const float* gld;        // backward: dL/dlog_det (B) (nullable)
    float* gx;               // backward: dL/dx (B, D), target half written here
    float* dhA;              // backward: dL/dh as T(Bp, P)
    float* dhT;              // backward: dL/dh as T(Pp, Bp)
    int Pp8;                 // rows / 8 of dhT
    float* gb2;              // backward: dL/db2 (Dh * 23), accumulated atomically (zeroed by the caller)
};

__device__ __forceinline__ float tf32_rn(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

__host__ __device__ __forceinline__ size_t tiled_off(long long r, long long k, long long R8) {
    return (size_t)(((k >> 5) * R8 + (r >> 3)) * 256 + ((k & 31) >> 2) * 32 + (r & 7) * 4 + (k & 3));
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

enum { WB_FULL = 0, WB_EMPTY = kMaxStages, WB_D_FULL = 2 * kMaxStages, WB_D_EMPTY, WB_COUNT };

// ---- epilogues ------------------------------------------------------------------------------------------------------------
// Thread mapping: warp w owns TMEM lanes 32 (w % 4) .. +31 (hardware rule), i.e. row 32 (w % 4) + lane of M tile
// t = w / 8, and the column half (w / 4) % 2 of the accumulator [128 x NT] of that tile (columns t * NT .. ).

template <class Release>
__device__ __forceinline__ void epi_store(const GArgs& G, int mt, int nt, int ncols, uint32_t tbase, int warp, int lane,
                                          const Release& release) {
    const int q = warp & 3, t = warp >> 3, half = (warp >> 2) & 1;
    const int hc = G.NT >> 1;                                // columns per half
    const long long m = (long long)mt * 256 + t * 128 + q * 32 + lane;
    long long row = m;
    bool live = m < G.M_real;
    if (G.remap24) {
        const long long e = m / 24;
        const int i = (int)(m - e * 24);
        row = e * 23 + i;
        live = live && i < 23;
    }
    float* crow = G.C + row * G.ldc;
    const uint32_t taddr = tbase + ((uint32_t)(q * 32) << 16) + t * G.NT + half * hc;
    const int col0 = nt * G.NT + half * hc;
    const bool vec = (G.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(G.C) & 15) == 0;
    for (int c = 0; c < hc; c += 16) {
        float v[16];
        if (half * hc + c < ncols) {                         // warp-uniform
            umma::tmem_ld8_nowait<0>(taddr + c, v);
            umma::tmem_ld8_nowait<8>(taddr + c + 8, v);
            umma::tmem_ld_wait();
        }
        if (c + 16 >= hc) release();
        if (half * hc + c >= ncols || !live) continue;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const int col = col0 + c + j;
            if (vec && col + 3 < G.N_real && half * hc + c + j + 3 < ncols) {
                if (G.atomic) red_add_v4(crow + col, v[j], v[j + 1], v[j + 2], v[j + 3]);
                else *reinterpret_cast<float4*>(crow + col) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (col + u < G.N_real && half * hc + c + j + u < ncols) {
                        if (G.atomic) atomicAdd(crow + col + u, v[j + u]);
                        else crow[col + u] = v[j + u];
                    }
            }
        }
    }
}

// Backward staging (per warp, 3 blocks of 8 parameters x 32 rows): the transposed orientation of dL/dh leaves the warp as
// three contiguous 1 KB runs of T(Pp, Bp) instead of 24 scattered 4-byte stores per thread; row stride 36 floats per
// 4-row group keeps the transposing writes conflict-free.  The same tile gives the bias gradient (column sums).
constexpr int kStageBlk = 8 * 36;                    // floats per staged block [k4 = row / 4 (8)][36: n % 8 (8) x row % 4 (4), padded]
constexpr int kStageWarp = 3 * kStageBlk;

template <int EPI, class Release>
__device__ __forceinline__ void epi_spline(const GArgs& G, long long row, uint32_t taddr, int e0, int n_el, int ldp_slot,
                                           int warp, int lane, float* stage, const Release& release) {
    // row: this thread's sample; taddr: TMEM address of its first parameter column; e0, n_el: its elements of the target half
    constexpr bool BWD = EPI == EPI_SPLINE_BWD_FWD || EPI == EPI_SPLINE_BWD_INV;
    constexpr bool INV = EPI == EPI_SPLINE_INV || EPI == EPI_SPLINE_BWD_INV;
    const bool live = row < G.B;
    const float* xrow = G.x + row * G.ldx + G.Dh;
    float ldacc = 0.0f;
    float GL = 0.0f;
    if (BWD) GL = (live && G.gld) ? __ldg(G.gld + row) : 0.0f;
    // this thread's inputs of element j are requested one element ahead, the bias before the wait on tensor memory: the
    // epilogue is a long dependent chain per element and every exposed L2 round trip is paid n_el times per item
    float v_next = live ? __ldg(xrow + e0) : 0.0f;
    float gz_next = 0.0f;
    if (BWD) gz_next = (live && G.gy) ? __ldg(G.gy + row * G.ldx + G.Dh + e0) : 0.0f;
#pragma unroll 1
    for (int j = 0; j < n_el; ++j) {
        float p[24];
        umma::tmem_ld8_nowait<0>(taddr + j * 24, p);
        umma::tmem_ld8_nowait<8>(taddr + j * 24 + 8, p);
        umma::tmem_ld8_nowait<16>(taddr + j * 24 + 16, p);
        const int e = e0 + j;
        const float4* bp = reinterpret_cast<const float4*>(G.b2p + (size_t)e * 24);
        float4 bv[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) bv[c] = __ldg(bp + c);
        const float v = v_next, GZ_cur = gz_next;
        if (j + 1 < n_el) {
            v_next = live ? __ldg(xrow + e + 1) : 0.0f;
            if (BWD) gz_next = (live && G.gy) ? __ldg(G.gy + row * G.ldx + G.Dh + e + 1) : 0.0f;
        }
        umma::tmem_ld_wait();
        if (j == n_el - 1) release();
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            p[4 * c] += bv[c].x; p[4 * c + 1] += bv[c].y; p[4 * c + 2] += bv[c].z; p[4 * c + 3] += bv[c].w;
        }
        auto h = [&](int i) { return p[i]; };
        if constexpr (!BWD) {
            float out, ld;
            int k;
            rq_apply<8, INV, 1>(v, h, 8, G.boundary, out, ld, k);
            if (live) G.y[row * G.ldx + G.Dh + e] = out;
            ldacc += ld;
        } else {
            const float GZ = GZ_cur;
            float dp[24];
            dp[23] = 0.0f;
            auto g = [&](int i, float val) { dp[i] = val; };
            float dv;
            if constexpr (INV) rq_backward_inv<8, 1>(v, h, 8, G.boundary, GZ, GL, dv, g);
            else rqf::backward_fwd(v, p, G.boundary, GZ, GL, dv, dp);        // density direction: the training hot path
            if (live) G.gx[row * G.ldx + G.Dh + e] = dv;
#pragma unroll
            for (int i = 0; i < 24; ++i) dp[i] = live ? tf32_rn(dp[i]) : 0.0f;
            // dL/dh in both operand orientations: (row, k = parameter) directly, (parameter, k = row) through the staging tile
            const long long n0 = (long long)e * 24;
#pragma unroll
            for (int c = 0; c < 6; ++c)           // streaming stores: 1.6 GB of dL/dh per call must not evict the GEMM operands from L2
                __stcs(reinterpret_cast<float4*>(G.dhA + tiled_off(row, n0 + 4 * c, G.Bp >> 3)),
                       make_float4(dp[4 * c], dp[4 * c + 1], dp[4 * c + 2], dp[4 * c + 3]));
            float* sw = stage + (warp * kStageWarp);
            __syncwarp();                                    // the previous element's tile has been read
#pragma unroll
            for (int i = 0; i < 24; ++i) sw[(i >> 3) * kStageBlk + (lane >> 2) * 36 + (i & 7) * 4 + (lane & 3)] = dp[i];
            __syncwarp();
            // rows of this warp: one 32-wide k-block of T(Pp, Bp); parameters n0 .. n0 + 23 = three 8-row groups
            const long long kb = row >> 5;
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int f4 = u * 32 + lane;                // float4 slot 0 .. 191 of the 3 KB
                const int blk = f4 >> 6, k4 = (f4 >> 3) & 7, n8 = f4 & 7;
                const float4 v = *reinterpret_cast<const float4*>(sw + blk * kStageBlk + k4 * 36 + n8 * 4);
                __stcs(reinterpret_cast<float4*>(G.dhT + ((size_t)kb * G.Pp8 + (size_t)(n0 >> 3) + blk) * 256 + k4 * 32 + n8 * 4), v);
            }
            if (lane < 23) {                                 // dL/db2[e * 23 + lane] += sum over the warp's 32 rows
                const float* col = sw + (lane >> 3) * kStageBlk + (lane & 7) * 4;
                float a = 0.0f;
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 v = *reinterpret_cast<const float4*>(col + k4 * 36);
                    a += (v.x + v.y) + (v.z + v.w);
                }
                atomicAdd(G.gb2 + (size_t)e * 23 + lane, a);
            }
        }
    }
    if constexpr (!BWD) G.ldp[(size_t)ldp_slot * G.Bp + row] = ldacc;
}

// ---- the GEMM kernel -------------------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(kThreads, 1) wide_gemm_kernel(const __grid_constant__ GArgs G) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int stage_bytes = kABytes + G.NT * 128;
    constexpr bool kBwd = EPI == EPI_SPLINE_BWD_FWD || EPI == EPI_SPLINE_BWD_INV;
    float* stage = reinterpret_cast<float*>(smem_raw + (size_t)G.stages * stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)G.stages * stage_bytes + (kBwd ? kEpiWarps * kStageWarp * 4 : 0));
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + WB_COUNT);

    if (tid == 0) {
        for (int i = 0; i < kMaxStages; ++i) {
            umma::mbar_init(&bars[WB_FULL + i], 1);
            umma::mbar_init(&bars[WB_EMPTY + i], 1);
        }
        umma::mbar_init(&bars[WB_D_FULL], 1);
        umma::mbar_init(&bars[WB_D_EMPTY], kEpiWarps);
        umma::fence_barrier_init();
    }
    if (warp == kEpiWarps) umma::tmem_alloc(tmem_ptr, kTmemCols);
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tc_fence_after_sync();
    const uint32_t tbase = *tmem_ptr;

    const int n_items = G.n_mt * G.n_nt * G.n_split;
    const int kb_per = (G.kb_total + G.n_split - 1) / G.n_split;
    auto decode = [&](int item, int& mt, int& nt, int& kb0, int& kb1) {
        nt = item % G.n_nt;
        const int r = item / G.n_nt;
        mt = r % G.n_mt;
        const int sp = r / G.n_mt;
        kb0 = sp * kb_per;
        kb1 = min(G.kb_total, kb0 + kb_per);
    };

    if (warp == kEpiWarps + 1) {
        // ===================== loader: two bulk copies per stage =====================
        if (lane == 0) {
            uint32_t st = 0, ph = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                int mt, nt, kb0, kb1;
                decode(item, mt, nt, kb0, kb1);
                const int ncols = min(G.NT, G.RB - nt * G.NT);
                const uint32_t b_bytes = (uint32_t)ncols * 128;
                for (int kb = kb0; kb < kb1; ++kb) {
                    umma::mbar_wait(&bars[WB_EMPTY + st], ph ^ 1);
                    uint8_t* sa = smem_raw + (size_t)st * stage_bytes;
                    umma::mbar_arrive_expect_tx(&bars[WB_FULL + st], kABytes + b_bytes);
                    umma::bulk_g2s(sa, G.A + ((size_t)kb * G.RA8 + (size_t)mt * 32) * 256, kABytes, &bars[WB_FULL + st]);
                    umma::bulk_g2s(sa + kABytes, G.Bm + ((size_t)kb * G.RB8 + (size_t)nt * (G.NT >> 3)) * 256, b_bytes,
                                   &bars[WB_FULL + st]);
                    if (++st == (uint32_t)G.stages) { st = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == kEpiWarps) {
        // ===================== MMA issuer =====================
        const uint32_t leader = umma::elect_one();
        const uint64_t dc = umma::make_smem_desc(0, 128, 1024);
        const uint32_t d_lo = (uint32_t)dc, d_hi = (uint32_t)(dc >> 32);
        const uint32_t s0 = umma::smem_u32(smem_raw) >> 4;
        uint32_t st = 0, ph = 0, it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            int mt, nt, kb0, kb1;
            decode(item, mt, nt, kb0, kb1);
            const int ncols = min(G.NT, G.RB - nt * G.NT);
            const uint32_t idesc = umma::make_idesc_tf32(128, ncols);
            umma::mbar_wait(&bars[WB_D_EMPTY], (it & 1) ^ 1);          // the epilogue has drained the accumulators
            umma::tc_fence_after_sync();
            for (int kb = kb0; kb < kb1; ++kb) {
                umma::mbar_wait(&bars[WB_FULL + st], ph);
                umma::tc_fence_after_sync();
                if (leader) {
                    const uint32_t a_lo = d_lo + s0 + st * (stage_bytes >> 4);
                    const uint32_t b_lo = a_lo + (kABytes >> 4);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t acc = (kb > kb0 || ks > 0) ? 1u : 0u;
                        umma::mma_tf32_ss_parts(tbase, a_lo + ks * 16, d_hi, b_lo + ks * 16, d_hi, idesc, acc);
                        umma::mma_tf32_ss_parts(tbase + G.NT, a_lo + (128 * 128 >> 4) + ks * 16, d_hi, b_lo + ks * 16, d_hi, idesc, acc);
                    }
                    umma::mma_commit(&bars[WB_EMPTY + st]);
                    if (kb == kb1 - 1) umma::mma_commit(&bars[WB_D_FULL]);
                }
                __syncwarp();
                if (++st == (uint32_t)G.stages) { st = 0; ph ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            int mt, nt, kb0, kb1;
            decode(item, mt, nt, kb0, kb1);
            const int ncols = min(G.NT, G.RB - nt * G.NT);
            umma::mbar_wait_backoff(&bars[WB_D_FULL], it & 1);
            umma::tc_fence_after_sync();
            auto release = [&]() {
                umma::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&bars[WB_D_EMPTY]);
            };
            if constexpr (EPI == EPI_STORE) epi_store(G, mt, nt, ncols, tbase, warp, lane, release);
            else {
                const int q = warp & 3, t = warp >> 3, half = (warp >> 2) & 1;
                epi_spline<EPI>(G, (long long)mt * 256 + t * 128 + q * 32 + lane, tbase + ((uint32_t)(q * 32) << 16) + t * 192 + half * 96,
                                nt * 8 + half * 4, 4, nt * 2 + half, warp, lane, stage, release);
            }
        }
    }
    umma::tc_fence_before_sync();
    __syncthreads();
    if (warp == kEpiWarps) umma::tmem_dealloc(tbase, kTmemCols);
}

// ---- the spline GEMMs as a 2-CTA cluster ----------------------------------------------------------------------------------------
// The spline epilogues (above all the backward one) are long; with two M tiles per CTA the accumulators fill 384 of the 512
// TMEM columns and the epilogue cannot overlap the next item's MMAs.  Here a CTA owns ONE 128-row tile and double-buffers
// its [128 x 192] accumulator; the B operand (the element chunk's weights) is shared by the two CTAs of a cluster, each
// loading half of every stage and multicasting it to both -- the same L2 -> SM bytes per flop as the two-tile kernel.
// A stage may be refilled once BOTH CTAs have consumed it: every MMA issuer commits its stage release to both CTAs.
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(umma::smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(umma::smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void mma_commit_multicast(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(umma::smem_u32(bar)), "h"(mask) : "memory");
}

enum { CB_FULL = 0, CB_EMPTY = kMaxStages, CB_D_FULL = 2 * kMaxStages, CB_D_EMPTY = 2 * kMaxStages + 2, CB_COUNT = 2 * kMaxStages + 4 };
constexpr int kCABytes = 128 * 128;       // one k-block of this CTA's M tile

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) wide_spline_kernel(const __grid_constant__ GArgs G) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr bool kBwd = EPI == EPI_SPLINE_BWD_FWD || EPI == EPI_SPLINE_BWD_INV;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t rank = cluster_rank();
    const int stage_bytes = kCABytes + 192 * 128;
    float* stage = reinterpret_cast<float*>(smem_raw + (size_t)G.stages * stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)G.stages * stage_bytes + (kBwd ? kEpiWarps * kStageWarp * 4 : 0));
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + CB_COUNT);
    if (tid == 0) {
        for (int i = 0; i < kMaxStages; ++i) {
            umma::mbar_init(&bars[CB_FULL + i], 1);
            umma::mbar_init(&bars[CB_EMPTY + i], 2);          // this CTA's issuer and the peer's
        }
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(&bars[CB_D_FULL + i], 1);
            umma::mbar_init(&bars[CB_D_EMPTY + i], kEpiWarps);
        }
        umma::fence_barrier_init();
    }
    if (warp == kEpiWarps) umma::tmem_alloc(tmem_ptr, kTmemCols);
    umma::tc_fence_before_sync();
    __syncthreads();
    cluster_sync();                                           // the peer's barriers exist before anything is sent to them
    umma::tc_fence_after_sync();
    const uint32_t tbase = *tmem_ptr;
    const int n_clusters = gridDim.x >> 1, cid = blockIdx.x >> 1;
    const int n_items = G.n_mt * G.n_nt;                      // (pair of M tiles, element chunk); nt fastest

    if (warp == kEpiWarps + 1) {
        if (lane == 0) {
            uint32_t st = 0, ph = 0;
            for (int item = cid; item < n_items; item += n_clusters) {
                const int nt = item % G.n_nt, tile = 2 * (item / G.n_nt) + (int)rank;
                for (int kb = 0; kb < G.kb_total; ++kb) {
                    umma::mbar_wait(&bars[CB_EMPTY + st], ph ^ 1);
                    uint8_t* sa = smem_raw + (size_t)st * stage_bytes;
                    umma::mbar_arrive_expect_tx(&bars[CB_FULL + st], kCABytes + 192 * 128);
                    umma::bulk_g2s(sa, G.A + ((size_t)kb * G.RA8 + (size_t)tile * 16) * 256, kCABytes, &bars[CB_FULL + st]);
                    // this CTA's half of the chunk's weights (96 of 192 rows), to both CTAs of the cluster
                    bulk_g2s_multicast(sa + kCABytes + rank * (96 * 128), G.Bm + ((size_t)kb * G.RB8 + (size_t)nt * 24 + rank * 12) * 256,
                                       96 * 128, &bars[CB_FULL + st], (uint16_t)3);
                    if (++st == (uint32_t)G.stages) { st = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == kEpiWarps) {
        const uint32_t leader = umma::elect_one();
        const uint64_t dc = umma::make_smem_desc(0, 128, 1024);
        const uint32_t d_lo = (uint32_t)dc, d_hi = (uint32_t)(dc >> 32);
        // in a cluster launch the shared::cta address carries the CTA's rank above bit 18 (0x1000400 for rank 1): the matrix
        // descriptor wants the 14-bit (address >> 4) of the CTA-local window, anything above would spill into its LBO field
        const uint32_t s0 = (umma::smem_u32(smem_raw) >> 4) & 0x3FFFu;
        const uint32_t idesc = umma::make_idesc_tf32(128, 192);
        uint32_t st = 0, ph = 0, it = 0;
        for (int item = cid; item < n_items; item += n_clusters, ++it) {
            const uint32_t set = it & 1;
            umma::mbar_wait(&bars[CB_D_EMPTY + set], ((it >> 1) & 1) ^ 1);
            umma::tc_fence_after_sync();
            for (int kb = 0; kb < G.kb_total; ++kb) {
                umma::mbar_wait(&bars[CB_FULL + st], ph);
                umma::tc_fence_after_sync();
                if (leader) {
                    const uint32_t a_lo = d_lo + s0 + st * (stage_bytes >> 4);
                    const uint32_t b_lo = a_lo + (kCABytes >> 4);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma::mma_tf32_ss_parts(tbase + set * 192, a_lo + ks * 16, d_hi, b_lo + ks * 16, d_hi, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                    mma_commit_multicast(&bars[CB_EMPTY + st], (uint16_t)3);
                    if (kb == G.kb_total - 1) umma::mma_commit(&bars[CB_D_FULL + set]);
                }
                __syncwarp();
                if (++st == (uint32_t)G.stages) { st = 0; ph ^= 1; }
            }
        }
    } else {
        const int q = warp & 3, cg = warp >> 2;               // lane quarter; column group: 2 of the chunk's 8 elements
        uint32_t it = 0;
        for (int item = cid; item < n_items; item += n_clusters, ++it) {
            const int nt = item % G.n_nt, tile = 2 * (item / G.n_nt) + (int)rank;
            const uint32_t set = it & 1;
            umma::mbar_wait_backoff(&bars[CB_D_FULL + set], (it >> 1) & 1);
            umma::tc_fence_after_sync();
            auto release = [&]() {
                umma::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&bars[CB_D_EMPTY + set]);
            };
            epi_spline<EPI>(G, (long long)tile * 128 + q * 32 + lane, tbase + ((uint32_t)(q * 32) << 16) + set * 192 + cg * 48,
                            nt * 8 + cg * 2, 2, nt * 4 + cg, warp, lane, stage, release);
        }
    }
    umma::tc_fence_before_sync();
    __syncthreads();
    cluster_sync();                                           // nothing of the peer's is still on its way to this CTA
    if (warp == kEpiWarps) umma::tmem_dealloc(tbase, kTmemCols);
}

static int n_sm();

template <int EPI>
static int launch_spline(cudaStream_t st, GArgs& G, const char* what) {
    constexpr bool kBwd = EPI == EPI_SPLINE_BWD_FWD || EPI == EPI_SPLINE_BWD_INV;
    const size_t stage_bytes = kCABytes + 192 * 128;
    const size_t extra = (kBwd ? (size_t)kEpiWarps * kStageWarp * 4 : 0) + CB_COUNT * 8 + 16;
    G.stages = std::min(kMaxStages, (int)((227 * 1024 - extra) / stage_bytes));
    const size_t smem = (size_t)G.stages * stage_bytes + extra;
    const int items = G.n_mt * G.n_nt;
    if (items <= 0) return B2F_OK;
    cudaError_t ce = (cudaError_t)raise_smem_limit((const void*)wide_spline_kernel<EPI>, smem);
    if (ce != cudaSuccess) return fail(B2F_ERR_CUDA, "cudaFuncSetAttribute(wide spline): %s", cudaGetErrorString(ce));
    // persistent clusters: as many as can be resident at once (GPCs with an odd number of free SMs leave one SM out)
    int max_clusters = n_sm() / 2;
    {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2 * (n_sm() / 2));
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, (const void*)wide_spline_kernel<EPI>, &cfg) == cudaSuccess && n > 0) max_clusters = std::min(max_clusters, n);
        else (void)cudaGetLastError();
    }
    const int grid = 2 * std::max(1, std::min(items, max_clusters));
    wide_spline_kernel<EPI><<<grid, kThreads, smem, st>>>(G);
    return check_launch(what);
}

// ---- pack kernel: row-major -> tiled operands -------------------------------------------------------------------------------
enum { PACK_COPY = 0, PACK_TANH_BIAS = 1, PACK_DTANH = 2 };

struct PArgs {
    const float* src;        // [R x C] row-major, row stride ld (rows >= R / columns >= C read as zero)
    long long ld;
    long long R, C;
    int remap24;             // source row of logical row n: (n / 24) * 23 + n % 24, zero when n % 24 == 23
    int op;
    const float* bias;       // TANH_BIAS / DTANH: per column
    const float* aux;        // DTANH: pre-activations [R x C] row-major (row stride ld_aux); value = src * (1 - tanh(aux + bias)^2)
    long long ld_aux;
    float* outA;             // T(A_rows, A_k): rows = R axis, k = C axis (nullable); pads are written as zeros
    long long RA8, A_rows, A_k;
    float* outT;             // T(T_rows, T_k): rows = C axis, k = R axis (nullable)
    long long RT8, T_rows, T_k;
    float* colsum;           // DTANH: per-column sums of the result, atomically added (nullable)
};

// one 32 x 32 patch per block (256 threads): both output formats store a patch as 1024 contiguous floats
__global__ void __launch_bounds__(256) pack_kernel(const PArgs P) {
    __shared__ float s[32][33];
    const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rr = ty + 8 * i;
        const long long r = r0 + rr, c = c0 + tx;
        float v = 0.0f;
        if (r < P.R && c < P.C) {
            long long sr = r;
            bool ok = true;
            if (P.remap24) {
                const long long e = r / 24;
                const int k = (int)(r - e * 24);
                sr = e * 23 + k;
                ok = k < 23;
            }
            if (ok) {
                v = __ldg(P.src + sr * P.ld + c);
                if (P.op == PACK_TANH_BIAS) v = tanhf(v + __ldg(P.bias + c));
                else if (P.op == PACK_DTANH) {
                    const float th = tanhf(__ldg(P.aux + sr * P.ld_aux + c) + __ldg(P.bias + c));
                    v *= fmaf(-th, th, 1.0f);
                }
            }
        }
        s[rr][tx] = v;
    }
    __syncthreads();
    if (P.colsum && threadIdx.x < 32) {
        float a = 0.0f;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) a += s[i][threadIdx.x];
        if (c0 + threadIdx.x < P.C) atomicAdd(P.colsum + c0 + threadIdx.x, a);
    }
    const int i = threadIdx.x;                 // float4 slot of the patch: [8-group (4)][k4 (8)][row in group (8)]
    const int g = i >> 6, k4 = (i >> 3) & 7, r8 = i & 7;
    if (P.outA && r0 < P.A_rows && c0 < P.A_k) {
        const int rr = g * 8 + r8;
        float4 v = make_float4(tf32_rn(s[rr][4 * k4]), tf32_rn(s[rr][4 * k4 + 1]), tf32_rn(s[rr][4 * k4 + 2]), tf32_rn(s[rr][4 * k4 + 3]));
        *reinterpret_cast<float4*>(P.outA + tiled_off(r0 + rr, c0 + 4 * k4, P.RA8)) = v;
    }
    if (P.outT && c0 < P.T_rows && r0 < P.T_k) {
        const int cc = g * 8 + r8;
        float4 v = make_float4(tf32_rn(s[4 * k4][cc]), tf32_rn(s[4 * k4 + 1][cc]), tf32_rn(s[4 * k4 + 2][cc]), tf32_rn(s[4 * k4 + 3][cc]));
        *reinterpret_cast<float4*>(P.outT + tiled_off(c0 + cc, r0 + 4 * k4, P.RT8)) = v;
    }
}

// bias padded to 24 per element
__global__ void pad_bias_kernel(const float* __restrict__ b2, float* __restrict__ b2p, int Dh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Dh * 24) {
        const int e = i / 24, k = i % 24;
        b2p[i] = k < 23 ? b2[e * 23 + k] : 0.0f;
    }
}

// source half passes through (dst[:, :Dh] = src[:, :Dh]); optionally log_det[r] = sum of the partials
__global__ void finish_kernel(const float* __restrict__ src, float* __restrict__ dst, long long B, int D, int Dh,
                              const float* __restrict__ ldp, int n_part, long long Bp, float* __restrict__ log_det) {
    const long long n4 = B * (Dh / 4);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const long long r = i / (Dh / 4);
        const int c = (int)(i - r * (Dh / 4)) * 4;
        if (dst && dst != src)
            *reinterpret_cast<float4*>(dst + r * D + c) =
                src ? __ldg(reinterpret_cast<const float4*>(src + r * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (log_det) {
        for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < B; r += stride) {
            float a = 0.0f;
            for (int p = 0; p < n_part; ++p) a += ldp[(size_t)p * Bp + r];
            log_det[r] = a;
        }
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------------
static inline long long up(long long v, long long m) { return (v + m - 1) / m * m; }

struct Shapes {
    long long B, Bp;
    int D, Dh, H, Hk, P, Pp, Hp, Dhp;      // Hk: H rounded up to the 32-wide k-block (the padded hidden units are exact zeros)
};

static Shapes shapes(long long B, int D, int H) {
    Shapes s;
    s.B = B; s.Bp = up(std::max<long long>(B, 1), 256);
    s.D = D; s.Dh = D / 2; s.H = H; s.Hk = (int)up(H, 32);
    s.P = s.Dh * 24; s.Pp = (int)up(s.P, 256); s.Hp = (int)up(H, 256); s.Dhp = (int)up(s.Dh, 256);
    return s;
}

// Two buffers.  "keep": everything the backward can reuse from the forward of the same step (packed operands and hidden
// activations; the T-forms only when the forward was told a backward follows).  "scratch": per-call temporaries.
struct Workspace {           // offsets in floats
    size_t xaA, W1t, W2p, b2p, pre, hidA, xaT, W1T, W2pT, hidT, keep_total;
    size_t ldp, dhA, dhT, dhid, dpreA, dpreT, scratch_total;
};

static Workspace layout(const Shapes& s, bool backward) {
    Workspace w;
    size_t o = 0;
    auto take = [&](size_t n) { size_t at = o; o += (n + 63) / 64 * 64; return at; };
    w.xaA = take((size_t)s.Bp * s.Dh);
    w.W1t = take((size_t)s.Hp * s.Dh);
    w.W2p = take((size_t)s.Pp * s.Hk);
    w.b2p = take((size_t)s.P);
    w.pre = take((size_t)s.Bp * s.Hk);
    w.hidA = take((size_t)s.Bp * s.Hk);
    w.xaT = w.W1T = w.W2pT = w.hidT = 0;
    if (backward) {
        w.xaT = take((size_t)s.Dhp * s.Bp);
        w.W1T = take((size_t)s.Dhp * s.Hk);
        w.W2pT = take((size_t)s.Hp * s.P);
        w.hidT = take((size_t)s.Hp * s.Bp);
    }
    w.keep_total = o;
    o = 0;
    w.ldp = take((size_t)(s.P / 192) * 4 * s.Bp);
    w.dhA = w.dhT = w.dhid = w.dpreA = w.dpreT = 0;
    if (backward) {
        w.dhA = take((size_t)s.Bp * s.P);
        w.dhT = take((size_t)s.Pp * s.Bp);
        w.dhid = take((size_t)s.Bp * s.Hk);
        w.dpreA = take((size_t)s.Bp * s.Hk);
        w.dpreT = take((size_t)s.Hp * s.Bp);
    }
    w.scratch_total = o;
    return w;
}

static int n_sm() {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

// src [R x C] -> outA = T(A_rows, up(C, 32)) and / or outT = T(T_rows, T_k) (the transpose: k runs over the R axis)
static int pack(cudaStream_t st, const float* src, long long ld, long long R, long long C, int remap24, int op,
                const float* bias, const float* aux, long long ld_aux, float* outA, long long A_rows, float* outT,
                long long T_rows, long long T_k, float* colsum) {
    PArgs P;
    P.src = src; P.ld = ld; P.R = R; P.C = C; P.remap24 = remap24; P.op = op; P.bias = bias; P.aux = aux; P.ld_aux = ld_aux;
    P.outA = outA; P.RA8 = A_rows / 8; P.A_rows = A_rows; P.A_k = up(C, 32);
    P.outT = outT; P.RT8 = T_rows / 8; P.T_rows = T_rows; P.T_k = T_k; P.colsum = colsum;
    const long long Rgrid = std::max(outA ? A_rows : 0, outT ? T_k : 0), Cgrid = std::max(outA ? P.A_k : 0, outT ? T_rows : 0);
    dim3 grid((unsigned)(Cgrid / 32), (unsigned)(Rgrid / 32));
    pack_kernel<<<grid, 256, 0, st>>>(P);
    return check_launch("b2f_wide (pack)");
}

template <int EPI>
static int launch_gemm(cudaStream_t st, GArgs& G, const char* what) {
    constexpr bool kBwd = EPI == EPI_SPLINE_BWD_FWD || EPI == EPI_SPLINE_BWD_INV;
    const size_t extra = (kBwd ? (size_t)kEpiWarps * kStageWarp * 4 : 0) + WB_COUNT * 8 + 16;
    G.stages = std::min(kMaxStages, (int)((227 * 1024 - extra) / (kABytes + G.NT * 128)));
    const size_t smem = (size_t)G.stages * (kABytes + G.NT * 128) + extra;
    const int items = G.n_mt * G.n_nt * G.n_split;
    if (items <= 0) return B2F_OK;
    cudaError_t ce = (cudaError_t)raise_smem_limit((const void*)wide_gemm_kernel<EPI>, smem);
    if (ce != cudaSuccess) return fail(B2F_ERR_CUDA, "cudaFuncSetAttribute(wide gemm): %s", cudaGetErrorString(ce));
    wide_gemm_kernel<EPI><<<std::min(items, n_sm()), kThreads, smem, st>>>(G);
    return check_launch(what);
}

// plain GEMM C[M x N] (+)= A[T(RAp, K)] * B[T(RBp, K)]^T
static int gemm_store(cudaStream_t st, const float* A, long long RAp, const float* Bm, long long RBp, int RB, long long K,
                      float* C, long long ldc, long long M_real, int N_real, int remap24, int atomic, int n_split, int NT) {
    GArgs G;
    memset(&G, 0, sizeof(G));
    G.A = A; G.Bm = Bm; G.RA8 = (int)(RAp / 8); G.RB8 = (int)(RBp / 8); G.RB = RB; G.NT = NT;
    G.n_mt = (int)(RAp / 256); G.n_nt = (RB + NT - 1) / NT; G.kb_total = (int)(K / 32);
    G.n_split = std::max(1, std::min(n_split, G.kb_total));
    // every split must own at least one k-block (an empty item would never signal its accumulator)
    const int kb_per = (G.kb_total + G.n_split - 1) / G.n_split;
    G.n_split = (G.kb_total + kb_per - 1) / kb_per;
    G.C = C; G.ldc = ldc; G.M_real = (int)M_real; G.N_real = N_real; G.remap24 = remap24; G.atomic = atomic;
    return launch_gemm<EPI_STORE>(st, G, "b2f_wide (gemm)");
}

// k-split of a plain GEMM: minimise waves x (k-blocks per item + epilogue) -- a split adds one atomic pass over the output
// per part, so short-K GEMMs stay unsplit and long-K ones fill the last wave
static int pick_split(int items, int kb_total) {
    const int sms = n_sm();
    int best = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= 8 && s <= kb_total; ++s) {
        const double waves = (double)((items * s + sms - 1) / sms);
        const double cost = waves * ((double)((kb_total + s - 1) / s) * 1400.0 + 8000.0);
        if (cost < best_cost * 0.98) { best_cost = cost; best = s; }
    }
    return best;
}

static int check_layer(const b2f_wide_layer_t* L, long long B) {
    if (!L || !L->W1 || !L->b1 || !L->W2 || !L->b2) return fail(B2F_ERR_INVALID, "wide coupling: null layer / parameter");
    if (L->tkind != B2F_T_RQ_FWD && L->tkind != B2F_T_RQ_INV) return fail(B2F_ERR_UNSUPPORTED, "wide coupling: spline transformers only");
    if (L->n_bins != 8) return fail(B2F_ERR_UNSUPPORTED, "wide coupling: n_bins must be 8");
    if (L->D < 64 || L->D % 64 != 0) return fail(B2F_ERR_UNSUPPORTED, "wide coupling: n_dim must be a multiple of 64");
    if (L->H < 1) return fail(B2F_ERR_INVALID, "wide coupling: n_hidden");
    if (!(L->boundary > 0.0f)) return fail(B2F_ERR_INVALID, "wide coupling: boundary");
    if (B < 0 || B > (1LL << 30)) return fail(B2F_ERR_INVALID, "wide coupling: batch size");
    return B2F_OK;
}

// conditioner up to the hidden activations: xaA (and xaT), W1t, pre = xa W1^T, hidA (and hidT) = tanh(pre + b1)
static int hidden_layer(cudaStream_t st, const b2f_wide_layer_t* L, const Shapes& s, const Workspace& w, float* ws, const float* x,
                        bool backward) {
    int rc;
    if ((rc = pack(st, x, s.D, s.B, s.Dh, 0, PACK_COPY, nullptr, nullptr, 0, ws + w.xaA, s.Bp, backward ? ws + w.xaT : nullptr,
                   s.Dhp, s.Bp, nullptr)) != B2F_OK) return rc;
    if ((rc = pack(st, L->W1, s.Dh, s.H, s.Dh, 0, PACK_COPY, nullptr, nullptr, 0, ws + w.W1t, s.Hp, backward ? ws + w.W1T : nullptr,
                   s.Dhp, s.Hk, nullptr)) != B2F_OK) return rc;
    if ((rc = gemm_store(st, ws + w.xaA, s.Bp, ws + w.W1t, s.Hp, s.Hk, s.Dh, ws + w.pre, s.Hk, s.Bp, s.Hk, 0, 0, 1, 256)) != B2F_OK) return rc;
    // columns >= H (padding of the k-block) are written as exact zeros: tanh is applied to the real hidden units only
    return pack(st, ws + w.pre, s.Hk, s.B, s.H, 0, PACK_TANH_BIAS, L->b1, nullptr, 0, ws + w.hidA, s.Bp, backward ? ws + w.hidT : nullptr,
                s.Hp, s.Bp, nullptr);
}

static void spline_gemm_args(GArgs& G, const b2f_wide_layer_t* L, const Shapes& s, const Workspace& w, float* ws, const float* x) {
    memset(&G, 0, sizeof(G));
    G.A = ws + w.hidA; G.Bm = ws + w.W2p; G.RA8 = (int)(s.Bp / 8); G.RB8 = s.Pp / 8; G.RB = s.P; G.NT = 192;
    G.n_mt = (int)(s.Bp / 256); G.n_nt = s.P / 192; G.n_split = 1; G.kb_total = s.Hk / 32;
    G.x = x; G.ldx = s.D; G.Dh = s.Dh; G.B = s.B; G.Bp = s.Bp; G.b2p = ws + w.b2p; G.boundary = L->boundary;
}

}  // namespace wide
}  // namespace b2f

using namespace b2f;
using namespace b2f::wide;

extern "C" int64_t b2f_wide_coupling_workspace(int64_t B, int32_t D, int32_t H, int32_t which) {
    if (B < 0 || D < 2 || H < 1 || which < 0 || which > 3) return 0;
    const Workspace w = layout(shapes(B, D, H), (which & 1) != 0);
    return (int64_t)((which < 2 ? w.keep_total : w.scratch_total) * sizeof(float));
}

// packed weights + hidden activations into `keep` (with the transposed forms when a backward follows)
static int prepare(cudaStream_t st, const b2f_wide_layer_t* L, const Shapes& s, const Workspace& w, float* kp, const float* x,
                   bool backward) {
    int rc;
    if ((rc = hidden_layer(st, L, s, w, kp, x, backward)) != B2F_OK) return rc;
    // output-layer weights T(Pp, H) for the GEMM whose epilogue is the spline; T(Hp, P) for dL/dhid
    if ((rc = pack(st, L->W2, s.H, s.P, s.H, 1, PACK_COPY, nullptr, nullptr, 0, kp + w.W2p, s.Pp, backward ? kp + w.W2pT : nullptr,
                   s.Hp, s.P, nullptr)) != B2F_OK) return rc;
    pad_bias_kernel<<<(s.P + 255) / 256, 256, 0, st>>>(L->b2, kp + w.b2p, s.Dh);
    return check_launch("b2f_wide (bias)");
}

extern "C" int b2f_wide_coupling_forward(const b2f_wide_layer_t* L, const float* x, float* y, float* log_det, int64_t B,
                                         void* keep, int64_t keep_bytes, void* scratch, int64_t scratch_bytes, int32_t flags,
                                         void* stream) {
    int rc = check_layer(L, B);
    if (rc != B2F_OK) return rc;
    if (B == 0) return B2F_OK;
    if (!x || !y || !keep || !scratch) return fail(B2F_ERR_INVALID, "wide coupling: null buffer");
    const bool for_bwd = (flags & B2F_WIDE_FOR_BACKWARD) != 0;
    const Shapes s = shapes(B, L->D, L->H);
    const Workspace w = layout(s, for_bwd), wf = layout(s, false);
    if ((size_t)keep_bytes < w.keep_total * sizeof(float) || (size_t)scratch_bytes < wf.scratch_total * sizeof(float))
        return fail(B2F_ERR_INVALID, "wide coupling: workspace too small");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(keep) |
         reinterpret_cast<uintptr_t>(scratch)) & 15)
        return fail(B2F_ERR_INVALID, "wide coupling: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    float *kp = (float*)keep, *sc = (float*)scratch;
    if ((rc = prepare(st, L, s, w, kp, x, for_bwd)) != B2F_OK) return rc;
    GArgs G;
    spline_gemm_args(G, L, s, w, kp, x);
    G.y = y; G.ldp = sc + w.ldp;
    // 2-CTA cluster kernel (one M tile per CTA, double-buffered accumulators, B multicast); B2F_WIDE_NO_CLUSTER=1 keeps the
    // two-tile single-CTA kernel
    const bool cluster = !getenv("B2F_WIDE_NO_CLUSTER");
    if (cluster)
        rc = L->tkind == B2F_T_RQ_INV ? launch_spline<EPI_SPLINE_INV>(st, G, "b2f_wide_coupling_forward")
                                      : launch_spline<EPI_SPLINE_FWD>(st, G, "b2f_wide_coupling_forward");
    else
        rc = L->tkind == B2F_T_RQ_INV ? launch_gemm<EPI_SPLINE_INV>(st, G, "b2f_wide_coupling_forward")
                                      : launch_gemm<EPI_SPLINE_FWD>(st, G, "b2f_wide_coupling_forward");
    if (rc != B2F_OK) return rc;
    finish_kernel<<<(unsigned)std::min<long long>(4096, (B * (s.Dh / 4) + 255) / 256), 256, 0, st>>>(x, y, B, s.D, s.Dh, sc + w.ldp, (cluster ? 4 : 2) * G.n_nt, s.Bp, log_det);
    return check_launch("b2f_wide (finish)");
}

extern "C" int b2f_wide_coupling_backward(const b2f_wide_layer_t* L, const float* x, const float* gy, const float* glog_det,
                                          float* gx, float* gW1, float* gb1, float* gW2, float* gb2, int64_t B, void* keep,
                                          int64_t keep_bytes, void* scratch, int64_t scratch_bytes, int32_t flags, void* stream) {
    int rc = check_layer(L, B);
    if (rc != B2F_OK) return rc;
    if (!x || !gx || !gW1 || !gb1 || !gW2 || !gb2 || !keep || !scratch) return fail(B2F_ERR_INVALID, "wide coupling backward: null buffer");
    const Shapes s = shapes(B, L->D, L->H);
    const Workspace w = layout(s, true);
    if ((size_t)keep_bytes < w.keep_total * sizeof(float) || (size_t)scratch_bytes < w.scratch_total * sizeof(float))
        return fail(B2F_ERR_INVALID, "wide coupling backward: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float *kp = (float*)keep, *sc = (float*)scratch;
    const size_t P23 = (size_t)s.Dh * 23;
    cudaMemsetAsync(gW1, 0, (size_t)s.H * s.Dh * 4, st);
    cudaMemsetAsync(gb1, 0, (size_t)s.H * 4, st);
    cudaMemsetAsync(gW2, 0, P23 * s.H * 4, st);
    cudaMemsetAsync(gb2, 0, P23 * 4, st);
    if (B == 0) return check_launch("b2f_wide_coupling_backward");
    cudaMemsetAsync(sc + w.dhid, 0, (size_t)s.Bp * s.Hk * 4, st);
    if (s.Pp > s.P) cudaMemsetAsync(sc + w.dhT, 0, (size_t)s.Pp * s.Bp * 4, st);
    // packed operands and hidden activations: kept by the forward of this step, or rebuilt from x and the parameters
    if (!(flags & B2F_WIDE_KEPT) && (rc = prepare(st, L, s, w, kp, x, true)) != B2F_OK) return rc;
    // recompute h tile by tile, spline backward in the epilogue: gx (target half), dL/dh in both orientations, dL/db2
    GArgs G;
    spline_gemm_args(G, L, s, w, kp, x);
    G.gy = gy; G.gld = glog_det; G.gx = gx; G.dhA = sc + w.dhA; G.dhT = sc + w.dhT; G.Pp8 = s.Pp / 8; G.gb2 = gb2;
    if (!getenv("B2F_WIDE_NO_CLUSTER"))
        rc = L->tkind == B2F_T_RQ_INV ? launch_spline<EPI_SPLINE_BWD_INV>(st, G, "b2f_wide_coupling_backward (spline)")
                                      : launch_spline<EPI_SPLINE_BWD_FWD>(st, G, "b2f_wide_coupling_backward (spline)");
    else
        rc = L->tkind == B2F_T_RQ_INV ? launch_gemm<EPI_SPLINE_BWD_INV>(st, G, "b2f_wide_coupling_backward (spline)")
                                      : launch_gemm<EPI_SPLINE_BWD_FWD>(st, G, "b2f_wide_coupling_backward (spline)");
    if (rc != B2F_OK) return rc;
    // dL/dW2[n, j] = sum_b dh[b, n] hid[b, j]
    {
        const int items = (s.Pp / 256) * ((s.Hk + 255) / 256);
        if ((rc = gemm_store(st, sc + w.dhT, s.Pp, kp + w.hidT, s.Hp, s.Hk, s.Bp, gW2, s.H, s.P, s.H, 1, 1,
                             pick_split(items, (int)(s.Bp / 32)), 256)) != B2F_OK) return rc;
    }
    // dL/dhid[b, j] = sum_n dh[b, n] W2[n, j]
    {
        const int items = (int)(s.Bp / 256) * ((s.Hk + 255) / 256);
        if ((rc = gemm_store(st, sc + w.dhA, s.Bp, kp + w.W2pT, s.Hp, s.Hk, s.P, sc + w.dhid, s.Hk, s.Bp, s.Hk, 0, 1,
                             pick_split(items, s.P / 32), 256)) != B2F_OK) return rc;
    }
    // through the tanh: dpre = dhid (1 - hid^2) in both orientations, dL/db1 = column sums
    if ((rc = pack(st, sc + w.dhid, s.Hk, s.B, s.H, 0, PACK_DTANH, L->b1, kp + w.pre, s.Hk, sc + w.dpreA, s.Bp, sc + w.dpreT, s.Hp,
                   s.Bp, gb1)) != B2F_OK) return rc;
    // dL/dW1[j, i] = sum_b dpre[b, j] xa[b, i]
    {
        const int items = (s.Hp / 256) * ((s.Dh + 255) / 256);
        if ((rc = gemm_store(st, sc + w.dpreT, s.Hp, kp + w.xaT, s.Dhp, s.Dh, s.Bp, gW1, s.Dh, s.H, s.Dh, 0, 1,
                             pick_split(items, (int)(s.Bp / 32)), 256)) != B2F_OK) return rc;
    }
    // dL/dxa = gy[:, :Dh] + dpre W1
    finish_kernel<<<(unsigned)std::min<long long>(4096, (B * (s.Dh / 4) + 255) / 256), 256, 0, st>>>(gy, gx, B, s.D, s.Dh, nullptr, 0, s.Bp, nullptr);
    if ((rc = check_launch("b2f_wide (finish)")) != B2F_OK) return rc;
    {
        const int items = (int)(s.Bp / 256) * ((s.Dh + 255) / 256);
        if ((rc = gemm_store(st, sc + w.dpreA, s.Bp, kp + w.W1T, s.Dhp, s.Dh, s.Hk, gx, s.D, s.B, s.Dh, 0, 1,
                             pick_split(items, s.Hk / 32), 256)) != B2F_OK) return rc;
    }
    return B2F_OK;
}
