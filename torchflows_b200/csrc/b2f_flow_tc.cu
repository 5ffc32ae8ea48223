// Tensor-core whole-flow kernel for coupling flows: the conditioner GEMMs run on tcgen05 (kind::tf32,
// accumulators in TMEM), the weight tiles are streamed by the TMA engine (cp.async.bulk + mbarrier), and the
// transformer (rational-quadratic spline, affine or shift) is the epilogue that reads its parameters straight out of
// TMEM.  One persistent CTA per SM; a tile = 128 samples (UMMA M = 128).
//
// Replaces the same reference code as b2f_flow.cu (bijections/base.py:203-232, layers_base.py:119-163,
// transforms.py:293-307, transformers/linear/affine.py:39-59,149-159, spline/rational_quadratic.py:45-200,
// flows.py:628-648); the difference to the generic kernel is only where the two Linear layers are evaluated.
//
// Precision.  Spline layers: single-pass TF32 (log_prob stays within 1e-4 abs/rel).  Affine and
// shift layers need an fp32-faithful conditioner (their log-det is a plain sum of conditioner outputs), so they use
// the 3xTF32 split  a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi  laid out along K: A = [a | a | a_lo], B = [w_hi | w_lo |
// w_hi] (the tensor core truncates the fp32 `a` to a_hi by itself; a_lo = a - trunc(a) is exact in fp32).
//
// Shared memory (CouplingRQNSF D = 256, H = 17: 199 KB; RealNVP D = 64, H = 9: 112 KB):
//   xlo, xhi   the two halves of the sample tile, each [128 x D/2] fp32 in the canonical K-major UMMA operand
//              layout (b2f_umma.cuh).  They are BOTH the resident activations of the flow and the A operand of
//              the first GEMM (the tensor core reads fp32 bit patterns as tf32), so x is never staged twice.
//   xl3        3xTF32 only: low parts of the source half [128 x D/2]
//   w1         first Linear as B operand [32 x K1], K1 = D/2 or 3*D/2, bulk-copied per layer
//   a2         tanh(hidden) as A operand of the second GEMM [128 x K2]; two extra columns are 1.0 and multiply the
//              (hi, lo) split of the bias b2, so the bias is added by the MMA
//   w2buf[2]   second Linear, one chunk = EPC target elements x CPE parameter columns = [N2 x K2], double buffered
//              (spline: 8 x 24 = 192, affine: 64 x 2 = 128, shift: 128 x 1 = 128)
// Tensor memory (512 columns): D2[0] cols 0.., D2[1] cols 192.. (chunk accumulators, ping-pong against the
// epilogue), D1 cols 384..415 (hidden pre-activations).
//
// Warp roles: warps 0..15 epilogue (warp w owns TMEM lanes 32*(w%4).. and a quarter of each chunk's elements; a
// thread owns one sample, so log-det accumulates in a register), warp 16 issues every tcgen05.mma, warp 17 drives
// the TMA loads.  All hand-offs are mbarriers; waits are bounded (trap instead of hang).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "b2f_flow_device.cuh"
#include "b2f_umma.cuh"

namespace b2f {

constexpr int kTcEpiWarps = 16;
constexpr int kTcThreads = (kTcEpiWarps + 2) * 32;
constexpr int kTcBufCols = 192;                      // TMEM columns reserved per GEMM2 accumulator buffer
constexpr int kTcN1 = 32;                            // UMMA N of GEMM1 (hidden units, padded)
constexpr int kTcTmemCols = 512;
constexpr int kTcColD1 = 384;

// chunk geometry of GEMM2 per transformer family: CPE parameter columns per element, EPC elements per chunk
template <int TK> struct TcGeom { static constexpr int CPE = 24, EPC = 8; };                    // RQ spline
template <> struct TcGeom<B2F_T_AFFINE_FWD> { static constexpr int CPE = 2, EPC = 64; };
template <> struct TcGeom<B2F_T_AFFINE_INV> { static constexpr int CPE = 2, EPC = 64; };
template <> struct TcGeom<B2F_T_SHIFT_ADD> { static constexpr int CPE = 1, EPC = 128; };
template <> struct TcGeom<B2F_T_SHIFT_SUB> { static constexpr int CPE = 1, EPC = 128; };

struct TcOp {
    int kind, tkind, H, K2;
    int x3;                     // 3xTF32 operand split (affine / shift layers)
    int made;                   // masked autoregressive one-pass layer: source = target = all D columns
    int n_chunks, N2;           // GEMM2: chunks per layer, UMMA N of a chunk
    float boundary;
    int flip_before;            // flip state when this op runs (host-computed)
    const float* value;         // ELEMENTWISE: (D,2)
    const float* b1;            // COUPLING: (32) padded
    const float* w1c;           // canonical [32 x Ds]
    const float* w2c;           // n_chunks x canonical [192 x K2]
    long long ws_off;           // float offset of this layer's saved input in TcArgs::ws
};

struct TcArgs {
    TcOp ops[B2F_MAX_OPS];
    int n_ops, D, flags, n_tiles;
    long long B;
    const float* x;
    float* y;
    float* log_det;
    float* log_prob;
    const float* base_loc;
    const float* base_log_scale;
    float* ws;                  // optional: the tile as it enters every conditioner layer is saved here (training)
};

struct TcSmem {
    float* xlo;
    float* xhi;
    float* xl3;
    float* w1;
    float* a2;
    float* w2buf;      // two chunk buffers, w2stride floats apart
    int w2stride;
    float* ldp;     // [4][128]
    float* ldacc;   // [128]
    float* lpin;    // [128]
    float* ea;      // [3*D]
    float* ldc;     // [1]
    uint64_t* bars; // see enum
    uint32_t* tmem_ptr;
};

enum { BAR_W1_FULL = 0, BAR_W1_EMPTY, BAR_A1_READY, BAR_D1_FULL, BAR_A2_FULL, BAR_W2_FULL0, BAR_W2_FULL1,
       BAR_W2_EMPTY0, BAR_W2_EMPTY1, BAR_D2_FULL0, BAR_D2_FULL1, BAR_D2_EMPTY0, BAR_D2_EMPTY1, BAR_COUNT };

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiWarps * 32) : "memory"); }

__device__ __forceinline__ float to_tf32_rn(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// physical location of logical column j: halves are [0, Dh) -> xlo, [Dh, D) -> xhi
__device__ __forceinline__ float* xaddr(const TcSmem& s, int Dh, int m, int c) {
    float* base = c < Dh ? s.xlo : s.xhi;
    const int k = c < Dh ? c : c - Dh;
    return reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(base) + umma::canon_off(m, k, Dh));
}

// chunk loop of one coupling layer for one epilogue thread: wait for the chunk's accumulator, pull this warp's
// parameter columns out of TMEM, apply the transformer to its elements, hand the buffer back.
template <int TK, int MODE>
__device__ __forceinline__ float tc_chunk_loop(const TcSmem& s, const TcOp& op, uint32_t tbase, uint32_t lane_addr,
                                               int Dh, int m_t, int sub, int lane, uint32_t& cc) {
    constexpr int CPE = TcGeom<TK>::CPE, EPC = TcGeom<TK>::EPC, EPS = EPC / 4;   // elements per epilogue sub-warp
    const int D = 2 * Dh, n_tgt = op.made ? D : Dh, t0 = op.made ? 0 : Dh;
    // logical target element e lives at logical column t0 + e = physical column (flip ? D-1-col : col)
    auto tgt_ptr = [&](int e) {
        const int col = t0 + e, c = op.flip_before ? D - 1 - col : col;
        uint8_t* base = reinterpret_cast<uint8_t*>(c < Dh ? s.xlo : s.xhi);
        return reinterpret_cast<float*>(base + umma::canon_off(m_t, c < Dh ? c : c - Dh, Dh));
    };
    float ldpart = 0.0f;
    for (int c = 0; c < op.n_chunks; ++c, ++cc) {
        const int b = cc & 1;
        umma::mbar_wait(&s.bars[BAR_D2_FULL0 + b], (cc >> 1) & 1);
        umma::tc_fence_after_sync();
        const uint32_t tcol = tbase + lane_addr + b * kTcBufCols + sub * (EPS * CPE);
        if constexpr (CPE == 24) {
#pragma unroll
            for (int i = 0; i < EPS; ++i) {
                const int e = c * EPC + sub * EPS + i;                 // logical target element
                float acc[24];
                umma::tmem_ld8_nowait<0>(tcol + i * 24, acc);
                umma::tmem_ld8_nowait<8>(tcol + i * 24 + 8, acc);
                umma::tmem_ld8_nowait<16>(tcol + i * 24 + 16, acc);
                umma::tmem_ld_wait();
                float* px = tgt_ptr(e);
                float out, ld;
                transform_element<TK, MODE, 24>(*px, acc, op.boundary, out, ld);
                *px = out;
                ldpart += ld;
            }
        } else {
            static_assert(EPS * CPE == 32, "affine / shift chunks hand 32 TMEM columns to every epilogue warp");
            float u[32];
            umma::tmem_ld8_nowait<0>(tcol, u);
            umma::tmem_ld8_nowait<8>(tcol + 8, u);
            umma::tmem_ld8_nowait<16>(tcol + 16, u);
            umma::tmem_ld8_nowait<24>(tcol + 24, u);
            umma::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < EPS; ++i) {
                const int e = c * EPC + sub * EPS + i;
                if (e < n_tgt) {
                    float acc[CPE];
#pragma unroll
                    for (int p = 0; p < CPE; ++p) acc[p] = u[i * CPE + p];
                    float* px = tgt_ptr(e);
                    float out, ld;
                    transform_element<TK, MODE, CPE>(*px, acc, op.boundary, out, ld);
                    *px = out;
                    ldpart += ld;
                }
            }
        }
        umma::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&s.bars[BAR_D2_EMPTY0 + b]);
    }
    return ldpart;
}

template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1) flow_tc_kernel(const __grid_constant__ TcArgs A) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int D = A.D, Dh = D >> 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- carve shared memory -----------------------------------------------------------------------------
    int K2max = 8, K1max = Dh, W2max = 0, any_x3 = 0;
    for (int i = 0; i < A.n_ops; ++i)
        if (A.ops[i].kind == B2F_OP_COUPLING || A.ops[i].kind == B2F_OP_MADE) {
            K2max = max(K2max, A.ops[i].K2);
            K1max = max(K1max, (A.ops[i].x3 ? 3 : 1) * (A.ops[i].made ? D : Dh));
            W2max = max(W2max, A.ops[i].N2 * A.ops[i].K2);
            any_x3 = max(any_x3, A.ops[i].x3 ? (A.ops[i].made ? 2 : 1) : 0);
        }
    TcSmem s;
    uint8_t* p = smem_raw;
    s.xlo = reinterpret_cast<float*>(p); p += 128 * Dh * 4;
    s.xhi = reinterpret_cast<float*>(p); p += 128 * Dh * 4;
    s.xl3 = reinterpret_cast<float*>(p); p += any_x3 * 128 * Dh * 4;      // low parts of one or both halves
    s.w1 = reinterpret_cast<float*>(p); p += kTcN1 * K1max * 4;
    s.a2 = reinterpret_cast<float*>(p); p += 128 * K2max * 4;
    s.w2buf = reinterpret_cast<float*>(p); p += 2 * W2max * 4;
    s.w2stride = W2max;
    s.ldp = reinterpret_cast<float*>(p); p += 4 * 128 * 4;
    s.ldacc = reinterpret_cast<float*>(p); p += 128 * 4;
    s.lpin = reinterpret_cast<float*>(p); p += 128 * 4;
    s.ea = reinterpret_cast<float*>(p); p += 3 * D * 4;
    s.ldc = reinterpret_cast<float*>(p); p += 16;
    s.bars = reinterpret_cast<uint64_t*>(p); p += BAR_COUNT * 8;
    s.tmem_ptr = reinterpret_cast<uint32_t*>(p);

    if (tid == 0) {
        umma::mbar_init(&s.bars[BAR_W1_FULL], 1);
        umma::mbar_init(&s.bars[BAR_W1_EMPTY], 1);
        umma::mbar_init(&s.bars[BAR_A1_READY], kTcEpiWarps);
        umma::mbar_init(&s.bars[BAR_D1_FULL], 1);
        umma::mbar_init(&s.bars[BAR_A2_FULL], 4);
        for (int b = 0; b < 2; ++b) {
            umma::mbar_init(&s.bars[BAR_W2_FULL0 + b], 1);
            umma::mbar_init(&s.bars[BAR_W2_EMPTY0 + b], 1);
            umma::mbar_init(&s.bars[BAR_D2_FULL0 + b], 1);
            umma::mbar_init(&s.bars[BAR_D2_EMPTY0 + b], kTcEpiWarps);
        }
        umma::fence_barrier_init();
    }
    if (warp == kTcEpiWarps) umma::tmem_alloc(s.tmem_ptr, kTcTmemCols);
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tc_fence_after_sync();
    const uint32_t tbase = *s.tmem_ptr;

    uint32_t lc = 0;   // coupling layers processed so far (all roles count identically)
    uint32_t cc = 0;   // GEMM2 chunks processed so far

    if (warp == kTcEpiWarps + 1) {
        // ===================== TMA loader (one thread) =====================
        if (lane == 0) {
            for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
                for (int oi = 0; oi < A.n_ops; ++oi) {
                    const TcOp& op = A.ops[oi];
                    if (op.kind != B2F_OP_COUPLING && op.kind != B2F_OP_MADE) continue;
                    umma::mbar_wait_backoff(&s.bars[BAR_W1_EMPTY], (lc & 1) ^ 1);
                    const uint32_t w1_bytes = kTcN1 * (op.x3 ? 3 : 1) * (op.made ? D : Dh) * 4;
                    umma::mbar_arrive_expect_tx(&s.bars[BAR_W1_FULL], w1_bytes);
                    umma::bulk_g2s(s.w1, op.w1c, w1_bytes, &s.bars[BAR_W1_FULL]);
                    const uint32_t ch_bytes = op.N2 * op.K2 * 4;
                    for (int c = 0; c < op.n_chunks; ++c, ++cc) {
                        const int b = cc & 1;
                        umma::mbar_wait_backoff(&s.bars[BAR_W2_EMPTY0 + b], ((cc >> 1) & 1) ^ 1);
                        umma::mbar_arrive_expect_tx(&s.bars[BAR_W2_FULL0 + b], ch_bytes);
                        umma::bulk_g2s(s.w2buf + b * s.w2stride, reinterpret_cast<const uint8_t*>(op.w2c) + (size_t)c * ch_bytes, ch_bytes,
                                       &s.bars[BAR_W2_FULL0 + b]);
                    }
                    ++lc;
                }
            }
        }
        __syncwarp();
    } else if (warp == kTcEpiWarps) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
                for (int oi = 0; oi < A.n_ops; ++oi) {
                    const TcOp& op = A.ops[oi];
                    if (op.kind != B2F_OP_COUPLING && op.kind != B2F_OP_MADE) continue;
                    const uint32_t ph = lc & 1;
                    umma::mbar_wait_backoff(&s.bars[BAR_W1_FULL], ph);
                    umma::mbar_wait_backoff(&s.bars[BAR_A1_READY], ph);
                    umma::tc_fence_after_sync();
                    // GEMM1: D1[128 x 32] = x_src[128 x Ks] * W1c[32 x Ks]^T.  Source = one half (coupling) or both halves in
                    // physical order (MADE).  3xTF32: K blocks [x | x | x_lo] . [w_hi | w_lo | w_hi].
                    const uint32_t xlo_a = umma::smem_u32(s.xlo), xhi_a = umma::smem_u32(s.xhi), xl3_a = umma::smem_u32(s.xl3);
                    const uint32_t b1a = umma::smem_u32(s.w1);
                    const uint32_t idesc1 = umma::make_idesc_tf32(128, kTcN1);
                    const int nblk = op.x3 ? 3 : 1, nseg = op.made ? 2 : 1, kpb = Dh / 8;
                    const uint32_t sbo_b = nblk * nseg * Dh * 32;
                    int kidx = 0;
                    for (int blk = 0; blk < nblk; ++blk)
                        for (int seg = 0; seg < nseg; ++seg) {
                            uint32_t a_base;
                            if (blk < 2) a_base = op.made ? (seg == 0 ? xlo_a : xhi_a) : (op.flip_before ? xhi_a : xlo_a);
                            else a_base = xl3_a + seg * 128 * Dh * 4;
                            for (int ks = 0; ks < kpb; ++ks, ++kidx)
                                umma::mma_tf32_ss(tbase + kTcColD1, umma::make_smem_desc(a_base + ks * 256, 128, Dh * 32),
                                                  umma::make_smem_desc(b1a + kidx * 256, 128, sbo_b), idesc1, kidx > 0);
                        }
                    umma::mma_commit(&s.bars[BAR_D1_FULL]);
                    umma::mma_commit(&s.bars[BAR_W1_EMPTY]);
                    umma::mbar_wait(&s.bars[BAR_A2_FULL], ph);
                    umma::tc_fence_after_sync();
                    const uint32_t a2a = umma::smem_u32(s.a2);
                    const uint32_t idesc2 = umma::make_idesc_tf32(128, op.N2);
                    for (int c = 0; c < op.n_chunks; ++c, ++cc) {
                        const int b = cc & 1;
                        const uint32_t ph2 = (cc >> 1) & 1;
                        umma::mbar_wait_backoff(&s.bars[BAR_W2_FULL0 + b], ph2);
                        umma::mbar_wait_backoff(&s.bars[BAR_D2_EMPTY0 + b], ph2 ^ 1);
                        umma::tc_fence_after_sync();
                        const uint32_t wb = umma::smem_u32(s.w2buf + b * s.w2stride);
                        for (int ks = 0; ks < op.K2 / 8; ++ks)
                            umma::mma_tf32_ss(tbase + b * kTcBufCols, umma::make_smem_desc(a2a + ks * 256, 128, op.K2 * 32),
                                              umma::make_smem_desc(wb + ks * 256, 128, op.K2 * 32), idesc2, ks > 0);
                        umma::mma_commit(&s.bars[BAR_D2_FULL0 + b]);
                        umma::mma_commit(&s.bars[BAR_W2_EMPTY0 + b]);
                    }
                    ++lc;
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps =====================
        const int q = warp & 3, sub = warp >> 2;          // TMEM lane quarter, element pair inside a chunk
        const int m_t = q * 32 + lane;                    // sample owned in the transformer phase
        const int rg = warp;                              // 8-row group owned in the elementwise / IO phases
        const int r8 = lane & 7, kq = lane >> 3;          // row inside the group, 16-byte chunk inside a 64-byte run
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
            const long long row0 = (long long)tile * 128;
            const int rows = (int)min(128LL, A.B - row0);
            // ---- load: a warp moves an 8-row group; 8 lanes x 16 B fill one core matrix (conflict-free) ----
            {
                const int m = rg * 8 + r8;
                const bool live = m < rows;
                const float4* src = reinterpret_cast<const float4*>(A.x + (row0 + m) * D);
                for (int kc = kq; kc < D / 4; kc += 4) {
                    const float4 v = live ? __ldg(src + kc) : make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(xaddr(s, Dh, m, 4 * kc)) = v;
                }
            }
            if (tid < 128) s.ldacc[tid] = 0.0f;
            if (tid == 0) s.ldc[0] = 0.0f;
            epi_sync();
            int flip = 0;
            const bool want_lp = A.log_prob != nullptr;
            // Gaussian base density of rows (warp owns its 8-row group; 4 lanes per row, shuffle reduction)
            auto base_logp = [&]() {
                const int m = rg * 8 + r8;
                float acc = 0.0f;
                for (int kc = kq; kc < D / 4; kc += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(xaddr(s, Dh, m, 4 * kc));
                    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = 4 * kc + i, j = flip ? D - 1 - c : c;
                        const float loc = A.base_loc ? __ldg(A.base_loc + j) : 0.0f;
                        const float lsc = A.base_log_scale ? __ldg(A.base_log_scale + j) : 0.0f;
                        acc += gauss_logp(vv[i], loc, lsc);
                    }
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 8);
                acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                if (kq == 0) s.lpin[m] = acc;
            };
            if (want_lp && (A.flags & B2F_FLOW_LOGP_OF_INPUT)) base_logp();

            for (int oi = 0; oi < A.n_ops; ++oi) {
                const TcOp& op = A.ops[oi];
                if (op.kind == B2F_OP_FLIP) { flip ^= 1; continue; }
                if (op.kind == B2F_OP_ELEMENTWISE) {
                    auto get = [&](int i) { return EwOp{A.ops[i].kind, A.ops[i].tkind, A.ops[i].value}; };
                    const int n_run = elementwise_stage_run(s.ea, get, oi, A.n_ops, D, tid, kTcEpiWarps * 32);
                    epi_sync();
                    const int m = rg * 8 + r8;
                    for (int kc = kq; kc < D / 4; kc += 4) {
                        float4* px = reinterpret_cast<float4*>(xaddr(s, Dh, m, 4 * kc));
                        float4 v = *px;
                        float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int c = 4 * kc + i, j = flip ? D - 1 - c : c;
                            vv[i] = fmaf(s.ea[j], vv[i], s.ea[D + j]);
                        }
                        *px = make_float4(vv[0], vv[1], vv[2], vv[3]);
                    }
                    if (warp == 0) {
                        float sum = 0.0f;
                        for (int j = lane; j < D; j += 32) sum += s.ea[2 * D + j];
                        sum = warp_sum(sum);
                        if (lane == 0) s.ldc[0] += sum;
                    }
                    epi_sync();
                    oi += n_run - 1;
                    continue;
                }
                // ---------------- coupling layer ----------------
                if (A.ws) {      // training: keep this layer's input for b2f_flow_backward (physical column order)
                    const int m = rg * 8 + r8;
                    if (m < rows) {
                        float4* dst = reinterpret_cast<float4*>(A.ws + op.ws_off + (row0 + m) * D);
                        for (int kc = kq; kc < D / 4; kc += 4)
                            __stcs(dst + kc, *reinterpret_cast<const float4*>(xaddr(s, Dh, m, 4 * kc)));
                    }
                }
                const uint32_t ph = lc & 1;
                if (op.x3) {
                    // low parts of the source columns for the 3xTF32 split: x_lo = x - trunc_tf32(x), exact in fp32
                    const int nseg = op.made ? 2 : 1;
                    const int m = rg * 8 + r8;
                    for (int seg = 0; seg < nseg; ++seg) {
                        const uint8_t* src = reinterpret_cast<const uint8_t*>(
                            op.made ? (seg == 0 ? s.xlo : s.xhi) : (op.flip_before ? s.xhi : s.xlo));
                        uint8_t* dst = reinterpret_cast<uint8_t*>(s.xl3) + seg * 128 * Dh * 4;
                        for (int kc = kq; kc < Dh / 4; kc += 4) {
                            const uint32_t off = umma::canon_off(m, 4 * kc, Dh);
                            const float4 v = *reinterpret_cast<const float4*>(src + off);
                            float4 lo;
                            lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                            lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                            lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                            lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                            *reinterpret_cast<float4*>(dst + off) = lo;
                        }
                    }
                }
                umma::fence_proxy_async_smem();            // our generic-proxy writes to xlo/xhi/xl3 -> tensor core
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&s.bars[BAR_A1_READY]);
                if (warp < 4) {
                    // hidden epilogue: D1 -> +b1 -> tanh -> tf32 (hi, and lo for 3xTF32) -> a2 (A operand of GEMM2);
                    // the two bias columns are 1.0
                    umma::mbar_wait(&s.bars[BAR_D1_FULL], ph);
                    umma::tc_fence_after_sync();
                    const int K2 = op.K2, H = op.H;
                    uint8_t* a2b = reinterpret_cast<uint8_t*>(s.a2);
                    for (int k4 = 0; k4 < K2; k4 += 4)
                        *reinterpret_cast<float4*>(a2b + umma::canon_off(m_t, k4, K2)) = make_float4(0.f, 0.f, 0.f, 0.f);
                    auto put = [&](int k, float val) { *reinterpret_cast<float*>(a2b + umma::canon_off(m_t, k, K2)) = val; };
#pragma unroll
                    for (int c0 = 0; c0 < kTcN1; c0 += 8) {
                        float v[8];
                        umma::tmem_ld8(tbase + lane_addr + kTcColD1 + c0, v);
                        umma::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int j = c0 + i;
                            if (j < H) {
                                const float t = tanhf(v[i] + __ldg(op.b1 + j));
                                const float hi = to_tf32_rn(t);
                                put(j, hi);
                                if (op.x3) { put(H + j, hi); put(2 * H + j, to_tf32_rn(t - hi)); }
                            }
                        }
                    }
                    const int kb = (op.x3 ? 3 : 1) * H;
                    put(kb, 1.0f); put(kb + 1, 1.0f);
                    umma::tc_fence_before_sync();
                    umma::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) umma::mbar_arrive(&s.bars[BAR_A2_FULL]);
                }
                float ldpart = 0.0f;
#define B2F_TC_CASE(TKV) case TKV: ldpart = tc_chunk_loop<TKV, MODE>(s, op, tbase, lane_addr, Dh, m_t, sub, lane, cc); break;
                switch (op.tkind) {
                    B2F_TC_CASE(B2F_T_RQ_FWD)
                    B2F_TC_CASE(B2F_T_RQ_INV)
                    B2F_TC_CASE(B2F_T_AFFINE_FWD)
                    B2F_TC_CASE(B2F_T_AFFINE_INV)
                    B2F_TC_CASE(B2F_T_SHIFT_ADD)
                    B2F_TC_CASE(B2F_T_SHIFT_SUB)
                }
#undef B2F_TC_CASE
                s.ldp[sub * 128 + m_t] = ldpart;
                epi_sync();
                if (tid < 128) s.ldacc[tid] += (s.ldp[tid] + s.ldp[128 + tid]) + (s.ldp[256 + tid] + s.ldp[384 + tid]);
                epi_sync();
                ++lc;
            }
            // ---- outputs of this tile ----
            if (want_lp && !(A.flags & B2F_FLOW_LOGP_OF_INPUT)) base_logp();
            epi_sync();
            if (tid < rows) {
                const float ld = s.ldacc[tid] + s.ldc[0];
                if (A.log_det) A.log_det[row0 + tid] = ld;
                if (want_lp) A.log_prob[row0 + tid] = s.lpin[tid] + ld;
            }
            if (A.y) {
                const int m = rg * 8 + r8;
                if (m < rows) {
                    float4* dst = reinterpret_cast<float4*>(A.y + (row0 + m) * D);
                    for (int kc = kq; kc < D / 4; kc += 4)
                        dst[kc] = *reinterpret_cast<const float4*>(xaddr(s, Dh, m, 4 * kc));   // host guarantees flip == 0 here
                }
            }
            epi_sync();   // the tile buffers are reused by the next tile's load
        }
    }
    umma::tc_fence_before_sync();
    __syncthreads();
    if (warp == kTcEpiWarps) umma::tmem_dealloc(tbase, kTcTmemCols);
}

}  // namespace b2f

namespace b2f {

// Returns 1 if the tensor-core kernel was launched, 0 if the program is not eligible (caller falls through to the
// generic kernel -- same library, same results up to tf32 rounding of the conditioner), < 0 on error.
int try_launch_flow_tc(const b2f_op_t* ops, int32_t n_ops, const float* x, float* y, float* log_det, float* log_prob,
                       const float* base_loc, const float* base_log_scale, int64_t B, int32_t D, int32_t flags,
                       void* stream, float* ws) {
    if (getenv("B2F_DISABLE_TC")) return 0;
    if (D % 16 != 0 || D < 32 || D > 256) return 0;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (y && (reinterpret_cast<uintptr_t>(y) & 15))) return 0;
    int flip = 0, n_coupling = 0, K2max = 8, K1max = D / 2, W2max = 0, any_x3 = 0;
    TcArgs A;
    memset(&A, 0, sizeof(A));
    long long ws_next = 0;
    for (int i = 0; i < n_ops; ++i) {
        const b2f_op_t& o = ops[i];
        TcOp& t = A.ops[i];
        t.kind = o.kind; t.tkind = o.tkind; t.H = o.n_hidden; t.boundary = o.boundary; t.flip_before = flip;
        if (o.kind == B2F_OP_FLIP) { flip ^= 1; continue; }
        if (o.kind == B2F_OP_ELEMENTWISE && (o.flags & B2F_FLAG_ROW_BIAS)) return 0;     // per-row parameters: generic kernel
        if (o.kind == B2F_OP_ELEMENTWISE) { t.value = (const float*)o.p[0]; continue; }
        if (o.kind != B2F_OP_COUPLING && o.kind != B2F_OP_MADE) return 0;
        if (o.flags & B2F_FLAG_ROW_BIAS) return 0;          // context-conditioned layers: generic kernel
        t.made = o.kind == B2F_OP_MADE;
        t.ws_off = ws_next;                 // same order and stride as b2f_flow_backward_workspace
        ws_next += B * (long long)D;
        if (!(o.flags & B2F_FLAG_TC_OPERANDS) || !o.p[4] || !o.p[5] || !o.p[1]) return 0;
        const bool rq = o.tkind == B2F_T_RQ_FWD || o.tkind == B2F_T_RQ_INV;
        const int Dh_ = D / 2;
        int cpe, epc;
        if (rq) { cpe = 24; epc = 8; if (o.n_bins != 8) return 0; }
        else if (o.tkind == B2F_T_AFFINE_FWD || o.tkind == B2F_T_AFFINE_INV) { cpe = 2; epc = 64; }
        else if (o.tkind == B2F_T_SHIFT_ADD || o.tkind == B2F_T_SHIFT_SUB) { cpe = 1; epc = 128; }
        else return 0;
        t.x3 = rq ? 0 : 1;
        const int kin = (t.x3 ? 3 : 1) * o.n_hidden + 2;
        t.K2 = (kin + 7) / 8 * 8;
        if (o.n_hidden < 1 || o.n_hidden > 32 || t.K2 > 64) return 0;
        if (((o.flags & B2F_FLAG_TC_FLIPPED) != 0) != (flip != 0))
            return fail(B2F_ERR_INVALID, "op %d: tensor-core operands were laid out for the wrong flip state", i);
        t.N2 = epc * cpe;
        t.n_chunks = ((t.made ? D : Dh_) + epc - 1) / epc;
        K2max = std::max(K2max, t.K2);
        K1max = std::max(K1max, (t.x3 ? 3 : 1) * (t.made ? D : Dh_));
        W2max = std::max(W2max, t.N2 * t.K2);
        any_x3 = std::max(any_x3, t.x3 ? (t.made ? 2 : 1) : 0);
        t.b1 = (const float*)o.p[1]; t.w1c = (const float*)o.p[4]; t.w2c = (const float*)o.p[5];
        if ((reinterpret_cast<uintptr_t>(t.w1c) & 15) || (reinterpret_cast<uintptr_t>(t.w2c) & 15)) return 0;
        ++n_coupling;
    }
    if (flip != 0 || n_coupling == 0) return 0;
    const int Dh = D / 2;
    const size_t smem = (size_t)(2 + any_x3) * 128 * Dh * 4 + (size_t)kTcN1 * K1max * 4 + (size_t)128 * K2max * 4 +
                        (size_t)2 * W2max * 4 + (size_t)(4 * 128 + 128 + 128 + 3 * D) * 4 + 16 + BAR_COUNT * 8 + 16;
    if (smem > 227 * 1024) return 0;
    A.n_ops = n_ops; A.D = D; A.flags = flags; A.B = B;
    A.n_tiles = (int)((B + 127) / 128);
    A.x = x; A.y = y; A.log_det = log_det; A.log_prob = log_prob; A.base_loc = base_loc; A.base_log_scale = base_log_scale;
    A.ws = ws;
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int grid = std::min(A.n_tiles, n_sm);
    auto kern = (flags & B2F_FLOW_MODE_PRECISE) ? flow_tc_kernel<0> : ((flags & B2F_FLOW_MODE_FAST_KNOTS) ? flow_tc_kernel<2> : flow_tc_kernel<1>);
    cudaError_t ce = (cudaError_t)raise_smem_limit((const void*)kern, smem);
    if (ce != cudaSuccess) return fail(B2F_ERR_CUDA, "cudaFuncSetAttribute(tc): %s", cudaGetErrorString(ce));
    kern<<<grid, kTcThreads, smem, (cudaStream_t)stream>>>(A);
    const int rc = check_launch("b2f_flow_apply (tensor-core kernel)");
    return rc == B2F_OK ? 1 : rc;
}

}  // namespace b2f
