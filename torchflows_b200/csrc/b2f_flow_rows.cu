// Whole-flow fused kernel, row-per-thread variant (kernel K3): small conditioners (hidden width <= 32) on the
// FP32 pipe with every per-sample quantity in registers.
//
// Same program, operands and results as the generic kernel in b2f_flow.cu (which stays the fallback for shapes this
// one does not take); replaces the same reference code (file:line relative to /root/reference/torchflows):
//   bijections/base.py:203-232, .../autoregressive/layers_base.py:119-163,202-223,300-318, layers.py:19-69,
//   .../conditioning/transforms.py:197-198,259-264,293-307, matrix/permutation.py:19-37, flows.py:628-648,
//   base_distributions/gaussian.py:46-54.
//
// Why a second mapping: for RealNVP / NICE / MAF / IAF shapes (D <= a few hundred, H ~ 6..17, 1-2 parameters per
// element) the conditioner is ~1-2 kFLOP per row -- far too small for a 128-row tensor-core tile pipeline (measured:
// barrier round trips dominate) and dominated by shared-memory traffic and __syncthreads in the warp-per-element
// mapping of b2f_flow.cu.  Here a THREAD owns R rows for the whole program:
//   * the hidden activations hid[R][HB] and the transformer parameters acc[R][PP] live in registers;
//   * every weight address is uniform over the CTA, so a weight is one broadcast 16-byte load from L1 that feeds
//     4*R FFMAs in each of the 32 lanes;
//   * a warp only ever touches its own 32*R rows of the shared-memory tile, so after the tile load there is no CTA
//     barrier at all: layers, the D-step sequential inverse of a masked autoregressive layer (layers_base.py:213-223)
//     and the base log-density run back to back, warps drift freely and hide each other's latency;
//   * the tile row stride XS is a multiple of 4 floats with XS/4 odd, so a thread reads/writes 4 consecutive columns
//     of its row with one conflict-free 16-byte shared-memory access.
// HBM traffic is the algorithmic minimum (x read once, z / log-density written once, weights from L1/L2).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "b2f_flow_device.cuh"

namespace b2f {

constexpr int kRowsThreads = 128;     // four warps; between the per-layer weight staging barriers they run independently

struct RowsArgs {
    DevOp ops[B2F_MAX_OPS];
    int n_ops, D, XS, flags, n_runs, wbuf_floats;
    long long B;
    const float* x;
    float* y;
    float* log_det;
    float* log_prob;
    const float* base_loc;
    const float* base_log_scale;
};

// 4 consecutive LOGICAL columns k0..k0+3 of a row (k0 % 4 == 0, D % 4 == 0); a flipped tile stores logical column j at
// physical column D-1-j, so the same 16 bytes are read in reverse order
__device__ __forceinline__ void load4(const float* xr, int k0, int flip, int D, float (&v)[4]) {
    if (!flip) {
        const float4 q = *reinterpret_cast<const float4*>(xr + k0);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
        const float4 q = *reinterpret_cast<const float4*>(xr + (D - 4 - k0));
        v[0] = q.w; v[1] = q.z; v[2] = q.y; v[3] = q.x;
    }
}
__device__ __forceinline__ void store4(float* xr, int k0, int flip, int D, const float (&v)[4]) {
    if (!flip) *reinterpret_cast<float4*>(xr + k0) = make_float4(v[0], v[1], v[2], v[3]);
    else *reinterpret_cast<float4*>(xr + (D - 4 - k0)) = make_float4(v[3], v[2], v[1], v[0]);
}

// An elementwise run (one affine map per column, coefficients er[0..D) and er[D..2D) indexed by the LOGICAL column at the
// time the run was reached) that has not been applied yet; `rev` = an odd number of FLIP ops happened since, so today's
// logical column k takes the coefficients of column D-1-k -- which is exactly what load4's flip does.
struct PendingRun {
    const float* er;
    int rev;
};
__device__ __forceinline__ void pending_coeffs(const PendingRun& pr, int k0, int D, float (&a)[4], float (&b)[4]) {
    load4(pr.er, k0, pr.rev, D, a);
    load4(pr.er + D, k0, pr.rev, D, b);
}

// Shared-memory image of one conditioner layer's weights, zero-padded from H to HP = 4*HP4 hidden units so that every
// inner loop has a compile-time trip count and every weight access is one broadcast 16-byte load at a constant offset:
//   w1  [n_src/4][HP][4]   w1[(k/4)*HP*4 + j*4 + k%4] = W1[j][k]          (one-pass layers)
//   w1c [D][HP]            w1c[i*HP + j]              = W1[j][i]          (sequential layers: rank-1 update of step i)
//   b1  [HP]
//   w2  [n_tgt*P][HP]      w2[(e*P + p)*HP + j]       = W2tile[e][j][p] for j < H,  b2[e][p] for j == H (HP > H: the
//                          kernel keeps a constant 1 in hidden slot H, so the bias costs no extra load)
template <int HP>
struct RowsWeights {
    const float *w1, *b1, *w2;
};

template <int HP, int P>
__device__ __forceinline__ RowsWeights<HP> stage_weights(float* wbuf, const DevOp& op, int n_src, int n_tgt, bool seq) {
    const int H = op.H, tid = threadIdx.x;
    float* w1 = wbuf;
    float* b1 = w1 + n_src * HP;
    float* w2 = b1 + HP;
    if (!seq) {
        for (int d = tid; d < n_src * HP; d += kRowsThreads) {
            const int k4 = d / (HP * 4), rem = d - k4 * (HP * 4), j = rem >> 2, k = 4 * k4 + (rem & 3);
            w1[d] = (j < H) ? __ldg(op.p0 + (size_t)j * n_src + k) : 0.0f;
        }
    } else {
        for (int d = tid; d < n_src * HP; d += kRowsThreads) {
            const int i = d / HP, j = d - i * HP;
            w1[d] = (j < H) ? __ldg(op.p0 + (size_t)j * n_src + i) : 0.0f;
        }
    }
    if (tid < HP) b1[tid] = (tid < H) ? __ldg(op.p1 + tid) : 0.0f;
    for (int d = tid; d < n_tgt * P * HP; d += kRowsThreads) {
        const int ep = d / HP, j = d - ep * HP, e = ep / P, p = ep - e * P;
        w2[d] = (j < H) ? __ldg(op.p2 + ((size_t)e * H + j) * P + p) : (j == H ? __ldg(op.p3 + ep) : 0.0f);
    }
    return RowsWeights<HP>{w1, b1, w2};
}

// acc[r][p] = sum_j W2[e][p][j] * hid[r][j]  with the bias in slot H   (transforms.py:297-300, last Linear)
template <int P, int HP, int R>
__device__ __forceinline__ void row_params(float (&acc)[R][P], const RowsWeights<HP>& W, int e, const float (&hid)[R][HP]) {
    const float4* w2 = reinterpret_cast<const float4*>(W.w2) + (size_t)e * P * (HP / 4);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        float a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = 0.0f;
#pragma unroll
        for (int j4 = 0; j4 < HP / 4; ++j4) {
            const float4 w = w2[p * (HP / 4) + j4];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                a[r] = fmaf(w.x, hid[r][4 * j4 + 0], a[r]);
                a[r] = fmaf(w.y, hid[r][4 * j4 + 1], a[r]);
                a[r] = fmaf(w.z, hid[r][4 * j4 + 2], a[r]);
                a[r] = fmaf(w.w, hid[r][4 * j4 + 3], a[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r][p] = a[r];
    }
}

// one-pass conditioner layer (coupling, or masked autoregressive in its parallel direction)
template <int TK, int MODE, int HP, int R>
__device__ __forceinline__ void rows_pass(float* x0, int XS, int D, int flip, const DevOp& op,
                                          const RowsWeights<HP>& W, int n_src, int t0, const PendingRun& pr, float (&ld)[R]) {
    constexpr int P = TInfo<TK>::P;
    const int H = op.H;
    float hid[R][HP];
#pragma unroll
    for (int j = 0; j < HP; ++j) {
        const float b = W.b1[j];
#pragma unroll
        for (int r = 0; r < R; ++r) hid[r][j] = b;
    }
    // hid[r][j] = tanh(b1[j] + sum_k W1[j][k] x[r][k])                 (transforms.py:295-296 / :259-262)
    const float4* w1 = reinterpret_cast<const float4*>(W.w1);
    for (int k0 = 0; k0 < n_src; k0 += 4, w1 += HP) {
        float xv[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) load4(x0 + r * 32 * XS, k0, flip, D, xv[r]);
        if (pr.er) {   // pending elementwise run: apply it to the source columns on the way in and write them back
            float a[4], b[4];
            pending_coeffs(pr, k0, D, a, b);
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int u = 0; u < 4; ++u) xv[r][u] = fmaf(a[u], xv[r][u], b[u]);
                store4(x0 + r * 32 * XS, k0, flip, D, xv[r]);
            }
        }
#pragma unroll
        for (int j = 0; j < HP; ++j) {
            const float4 w = w1[j];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                hid[r][j] = fmaf(w.x, xv[r][0], hid[r][j]);
                hid[r][j] = fmaf(w.y, xv[r][1], hid[r][j]);
                hid[r][j] = fmaf(w.z, xv[r][2], hid[r][j]);
                hid[r][j] = fmaf(w.w, xv[r][3], hid[r][j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < HP; ++j) {
#pragma unroll
        for (int r = 0; r < R; ++r) hid[r][j] = (j == H) ? 1.0f : tanhf(hid[r][j]);   // slot H carries the bias; padded
    }                                                                                  // units: tanh(0) = 0 times zero weights
    const int n_tgt = D - t0;
    const bool ew_targets = pr.er != nullptr && t0 >= n_src;    // coupling: the targets were not touched by the loop above
    for (int e0 = 0; e0 < n_tgt; e0 += 4) {
        float xv[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) load4(x0 + r * 32 * XS, t0 + e0, flip, D, xv[r]);
        if (ew_targets) {
            float a[4], b[4];
            pending_coeffs(pr, t0 + e0, D, a, b);
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int u = 0; u < 4; ++u) xv[r][u] = fmaf(a[u], xv[r][u], b[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float acc[R][P];
            row_params<P, HP, R>(acc, W, e0 + u, hid);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float out, l;
                transform_element<TK, MODE, P>(xv[r][u], acc[r], op.boundary, out, l);
                xv[r][u] = out;
                ld[r] += l;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) store4(x0 + r * 32 * XS, t0 + e0, flip, D, xv[r]);
    }
}

// D-step sequential direction of a masked autoregressive layer (layers_base.py:213-223) at the cost of ONE conditioner
// pass: the hidden pre-activations get a rank-1 update per finished dimension and only the P parameters of dimension i
// are evaluated at step i.  Everything of a sample is in its thread's registers: no barrier in the D-step loop.
template <int TK, int MODE, int HP, int R>
__device__ __forceinline__ void rows_sequential(float* x0, int XS, int D, int flip, const DevOp& op,
                                                const RowsWeights<HP>& W, const PendingRun& pr, float (&ld)[R]) {
    constexpr int P = TInfo<TK>::P;
    const int H = op.H;
    float pre[R][HP], act[R][HP];
    int fin[HP];
#pragma unroll
    for (int j = 0; j < HP; ++j) {
        const float b = W.b1[j];
        fin[j] = (j < H) ? __ldg(op.p4 + j) : -1;
#pragma unroll
        for (int r = 0; r < R; ++r) { pre[r][j] = b; act[r][j] = (j == H) ? 1.0f : 0.0f; }   // slot H carries the bias
    }
    const float4* w1c = reinterpret_cast<const float4*>(W.w1);
    for (int i0 = 0; i0 < D; i0 += 4) {
        float xv[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) load4(x0 + r * 32 * XS, i0, flip, D, xv[r]);
        if (pr.er) {
            float a[4], b[4];
            pending_coeffs(pr, i0, D, a, b);
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int u = 0; u < 4; ++u) xv[r][u] = fmaf(a[u], xv[r][u], b[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u;
            // hidden units whose inputs x_0..x_{i-1} are now all final
#pragma unroll
            for (int j = 0; j < HP; ++j)
                if (fin[j] == i) {
#pragma unroll
                    for (int r = 0; r < R; ++r) act[r][j] = tanhf(pre[r][j]);
                }
            float acc[R][P];
            row_params<P, HP, R>(acc, W, i, act);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float out, l;
                transform_element<TK, MODE, P>(xv[r][u], acc[r], op.boundary, out, l);
                xv[r][u] = out;
                ld[r] += l;
            }
#pragma unroll
            for (int j4 = 0; j4 < HP / 4; ++j4) {
                const float4 w = w1c[i * (HP / 4) + j4];          // column i of the (masked) first layer
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    pre[r][4 * j4 + 0] = fmaf(w.x, xv[r][u], pre[r][4 * j4 + 0]);
                    pre[r][4 * j4 + 1] = fmaf(w.y, xv[r][u], pre[r][4 * j4 + 1]);
                    pre[r][4 * j4 + 2] = fmaf(w.z, xv[r][u], pre[r][4 * j4 + 2]);
                    pre[r][4 * j4 + 3] = fmaf(w.w, xv[r][u], pre[r][4 * j4 + 3]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) store4(x0 + r * 32 * XS, i0, flip, D, xv[r]);
    }
}

template <int TK, int MODE, int HP, int R>
__device__ __forceinline__ void rows_layer_tk(float* wbuf, float* x0, int XS, int D, int flip, const DevOp& op,
                                              const PendingRun& pr, float (&ld)[R]) {
    const bool seq = op.kind == B2F_OP_MADE_SEQ;
    const bool coupling = op.kind == B2F_OP_COUPLING;
    const int n_src = coupling ? D / 2 : D, t0 = coupling ? D / 2 : 0;
    __syncthreads();                        // every warp is done with the previous layer's weights
    const RowsWeights<HP> W = stage_weights<HP, TInfo<TK>::P>(wbuf, op, n_src, D - t0, seq);
    __syncthreads();
    if (seq) rows_sequential<TK, MODE, HP, R>(x0, XS, D, flip, op, W, pr, ld);
    else rows_pass<TK, MODE, HP, R>(x0, XS, D, flip, op, W, n_src, t0, pr, ld);
}

template <int MODE, int HP, int R>
__device__ __forceinline__ void rows_layer(float* wbuf, float* x0, int XS, int D, int flip, const DevOp& op,
                                           const PendingRun& pr, float (&ld)[R]) {
    switch (op.tkind) {
        case B2F_T_SHIFT_ADD: rows_layer_tk<B2F_T_SHIFT_ADD, MODE, HP, R>(wbuf, x0, XS, D, flip, op, pr, ld); break;
        case B2F_T_SHIFT_SUB: rows_layer_tk<B2F_T_SHIFT_SUB, MODE, HP, R>(wbuf, x0, XS, D, flip, op, pr, ld); break;
        case B2F_T_AFFINE_FWD: rows_layer_tk<B2F_T_AFFINE_FWD, MODE, HP, R>(wbuf, x0, XS, D, flip, op, pr, ld); break;
        case B2F_T_AFFINE_INV: rows_layer_tk<B2F_T_AFFINE_INV, MODE, HP, R>(wbuf, x0, XS, D, flip, op, pr, ld); break;
        default: break;
    }
}

template <int MODE, int HP4, int R>
__global__ void __launch_bounds__(kRowsThreads) flow_rows_kernel(const __grid_constant__ RowsArgs A) {
    extern __shared__ __align__(16) float smem[];
    constexpr int NW = kRowsThreads / 32, TMW = 32 * R, HP = 4 * HP4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = A.D, XS = A.XS, D4 = D >> 2;
    float* xw = smem + (size_t)warp * TMW * XS;          // this warp's [32*R][XS] rows
    float* ea = smem + (size_t)NW * TMW * XS;            // [n_runs][2][D]  A_j, B_j of every elementwise run
    float* gb = ea + (size_t)A.n_runs * 2 * D;           // [2][D]          base loc_j, 1/scale_j
    float* red = gb + 2 * D;                             // [2][NW]         CTA-uniform constants (partials per warp)
    float* wbuf = red + 2 * NW;                          // staged weights of the current conditioner layer

    // ---- this warp's rows: asynchronous 16-byte copies global -> shared (rows beyond B are zero-filled); they are in
    //      flight while the batch-independent constants below are computed ---------------------------------------------
    const long long wrow0 = ((long long)blockIdx.x * NW + warp) * TMW;
    const int rows = (int)max(0LL, min((long long)TMW, A.B - wrow0));
    {
        const float4* src = reinterpret_cast<const float4*>(A.x + wrow0 * D);
        int m = 0, c4 = lane;
        while (c4 >= D4) { c4 -= D4; ++m; }
        for (int idx = lane; m < TMW; idx += 32) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(xw + m * XS + 4 * c4);
            const int nbytes = (m < rows) ? 16 : 0;                     // src-size 0: the 16 bytes are zero-filled
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src + (m < rows ? idx : 0)), "r"(nbytes)
                         : "memory");
            c4 += 32;
            while (c4 >= D4) { c4 -= D4; ++m; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // ---- batch-independent part: every run of consecutive elementwise layers is one affine map per column -----------
    {
        float lsum = 0.0f, gsum = 0.0f;
        int run = 0;
        for (int oi = 0; oi < A.n_ops; ++oi) {
            if (A.ops[oi].kind != B2F_OP_ELEMENTWISE) continue;
            int n_run = 1;
            while (oi + n_run < A.n_ops && A.ops[oi + n_run].kind == B2F_OP_ELEMENTWISE) ++n_run;
            float* er = ea + (size_t)run * 2 * D;
            for (int j = tid; j < D; j += kRowsThreads) {
                float Aj = 1.0f, Bj = 0.0f;
                for (int r = 0; r < n_run; ++r) {
                    const DevOp& o = A.ops[oi + r];
                    float a, la;
                    affine_scale<0>(__ldg(o.p0 + 2 * j), a, la);
                    const float b = __ldg(o.p0 + 2 * j + 1);
                    if (o.tkind == B2F_T_AFFINE_FWD) { Aj *= a; Bj = fmaf(a, Bj, b); lsum += la; }
                    else { const float ia = 1.0f / a; Aj *= ia; Bj = (Bj - b) * ia; lsum -= la; }
                }
                er[j] = Aj; er[D + j] = Bj;
            }
            oi += n_run - 1;
            ++run;
        }
        if (A.log_prob) {
            for (int j = tid; j < D; j += kRowsThreads) {
                const float lsc = A.base_log_scale ? __ldg(A.base_log_scale + j) : 0.0f;
                gb[j] = A.base_loc ? __ldg(A.base_loc + j) : 0.0f;
                gb[D + j] = expf(-lsc);
                gsum += 0.91893853320467274178f + lsc;
            }
        }
        lsum = warp_sum(lsum);
        gsum = warp_sum(gsum);
        if (lane == 0) { red[warp] = lsum; red[NW + warp] = gsum; }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();   // staged constants + tile visible
    float ldc = 0.0f, gconst = 0.0f;
#pragma unroll
    for (int w = 0; w < NW; ++w) { ldc += red[w]; gconst += red[NW + w]; }

    float* x0 = xw + lane * XS;                          // row r of this thread: x0 + r*32*XS
    float ld[R], lp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { ld[r] = 0.0f; lp[r] = 0.0f; }
    const bool want_lp = A.log_prob != nullptr;
    int flip = 0;
    PendingRun pr{nullptr, 0};      // elementwise run that has been reached but not applied yet: the next conditioner layer
                                    // (or the epilogue) applies it on the fly to the values it loads anyway
    // DiagonalGaussian.log_prob (gaussian.py:46-54) of the thread's rows as they stand (after the pending run, if any)
    auto base_logp = [&]() {
        float s[R];
#pragma unroll
        for (int r = 0; r < R; ++r) s[r] = 0.0f;
        for (int c0 = 0; c0 < D; c0 += 4) {
            float loc[4], isc[4], a[4], b[4];
            load4(gb, c0, 0, D, loc);
            load4(gb + D, c0, 0, D, isc);
            if (pr.er) pending_coeffs(pr, c0, D, a, b);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float xv[4];
                load4(x0 + r * 32 * XS, c0, flip, D, xv);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float v = pr.er ? fmaf(a[u], xv[u], b[u]) : xv[u];
                    const float t = (v - loc[u]) * isc[u];
                    s[r] = fmaf(0.5f * t, t, s[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) lp[r] = -(s[r] + gconst);
    };
    if (want_lp && (A.flags & B2F_FLOW_LOGP_OF_INPUT)) base_logp();

    // ---- the layers ------------------------------------------------------------------------------------------------
    int run = 0;
    for (int oi = 0; oi < A.n_ops; ++oi) {
        const DevOp& op = A.ops[oi];
        if (op.kind == B2F_OP_FLIP) { flip ^= 1; pr.rev ^= 1; continue; }
        if (op.kind == B2F_OP_ELEMENTWISE) {
            while (oi + 1 < A.n_ops && A.ops[oi + 1].kind == B2F_OP_ELEMENTWISE) ++oi;
            if (pr.er) {                            // two runs separated only by FLIPs (no preset does this): apply the first
                for (int c0 = 0; c0 < D; c0 += 4) {
                    float a[4], b[4];
                    pending_coeffs(pr, c0, D, a, b);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float xv[4];
                        load4(x0 + r * 32 * XS, c0, flip, D, xv);
#pragma unroll
                        for (int u = 0; u < 4; ++u) xv[u] = fmaf(a[u], xv[u], b[u]);
                        store4(x0 + r * 32 * XS, c0, flip, D, xv);
                    }
                }
            }
            pr.er = ea + (size_t)run * 2 * D;       // every run is followed by a conditioner layer or the epilogue
            pr.rev = 0;
            ++run;
            continue;
        }
        rows_layer<MODE, HP, R>(wbuf, x0, XS, D, flip, op, pr, ld);
        pr.er = nullptr;
    }

    // ---- epilogue ------------------------------------------------------------------------------------------------
    if (want_lp && !(A.flags & B2F_FLOW_LOGP_OF_INPUT)) base_logp();
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int m = r * 32 + lane;
        if (m < rows) {
            const float l = ld[r] + ldc;
            if (A.log_det) A.log_det[wrow0 + m] = l;
            if (want_lp) A.log_prob[wrow0 + m] = lp[r] + l;
        }
    }
    if (A.y) {
        __syncwarp();
        float4* dst = reinterpret_cast<float4*>(A.y + wrow0 * D);
        int m = 0, c4 = lane;
        while (c4 >= D4) { c4 -= D4; ++m; }
        for (int idx = lane; m < rows; idx += 32) {
            float v[4];
            load4(xw + m * XS, 4 * c4, flip, D, v);
            if (pr.er) {
                float a[4], b[4];
                pending_coeffs(pr, 4 * c4, D, a, b);
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = fmaf(a[u], v[u], b[u]);
            }
            __stcs(dst + idx, make_float4(v[0], v[1], v[2], v[3]));
            c4 += 32;
            while (c4 >= D4) { c4 -= D4; ++m; }
        }
    }
}

template <int MODE, int HP4>
static cudaError_t launch_rows(const RowsArgs& A, int R, unsigned grid, size_t smem, cudaStream_t st) {
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return ce;
        kern<<<grid, kRowsThreads, smem, st>>>(A);
        return cudaSuccess;
    };
    (void)R;     // two rows per thread measured slower at every preset shape (occupancy beats weight reuse): one row
    return go(flow_rows_kernel<MODE, HP4, 1>);
}
template <int MODE>
static cudaError_t launch_rows_h(const RowsArgs& A, int hp4, int R, unsigned grid, size_t smem, cudaStream_t st) {
    switch (hp4) {
        case 1: return launch_rows<MODE, 1>(A, R, grid, smem, st);
        case 2: return launch_rows<MODE, 2>(A, R, grid, smem, st);
        case 3: return launch_rows<MODE, 3>(A, R, grid, smem, st);
        case 4: return launch_rows<MODE, 4>(A, R, grid, smem, st);
        case 5: return launch_rows<MODE, 5>(A, R, grid, smem, st);
        case 6: return launch_rows<MODE, 6>(A, R, grid, smem, st);
        case 7: return launch_rows<MODE, 7>(A, R, grid, smem, st);
        default: return launch_rows<MODE, 8>(A, R, grid, smem, st);
    }
}

// Returns 0 = not eligible (caller falls through), 1 = launched, negative = error.
int try_launch_flow_rows(const b2f_op_t* ops, int32_t n_ops, const float* x, float* y, float* log_det, float* log_prob,
                         const float* base_loc, const float* base_log_scale, int64_t B, int32_t D, int32_t flags,
                         void* stream) {
    if (getenv("B2F_DISABLE_ROWS")) return 0;
    if (D % 8 != 0 || D > 1024) return 0;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (y && (reinterpret_cast<uintptr_t>(y) & 15))) return 0;
    RowsArgs A;
    memset(&A, 0, sizeof(A));
    int Hmax = 0, n_runs = 0, n_cond = 0;
    bool prev_ew = false;
    for (int i = 0; i < n_ops; ++i) {
        const b2f_op_t& o = ops[i];
        DevOp& d = A.ops[i];
        d.kind = o.kind; d.tkind = o.tkind; d.H = o.n_hidden; d.flags = o.flags; d.boundary = o.boundary;
        d.p0 = (const float*)o.p[0]; d.p1 = (const float*)o.p[1]; d.p2 = (const float*)o.p[2];
        d.p3 = (const float*)o.p[3]; d.p4 = (const int*)o.p[4];
        const bool ew = o.kind == B2F_OP_ELEMENTWISE;
        if (ew && !prev_ew) ++n_runs;
        prev_ew = ew;
        switch (o.kind) {
            case B2F_OP_FLIP: break;
            case B2F_OP_ELEMENTWISE:
                if (!o.p[0] || (o.tkind != B2F_T_AFFINE_FWD && o.tkind != B2F_T_AFFINE_INV)) return 0;
                break;
            case B2F_OP_COUPLING: case B2F_OP_MADE: case B2F_OP_MADE_SEQ: {
                if (!o.p[0] || !o.p[1] || !o.p[2] || !o.p[3] || o.n_hidden <= 0 || o.n_hidden > 32) return 0;
                // 1-2 parameters per element only: a spline's 23-parameter output layer belongs on the tensor cores
                if (o.tkind != B2F_T_SHIFT_ADD && o.tkind != B2F_T_SHIFT_SUB && o.tkind != B2F_T_AFFINE_FWD &&
                    o.tkind != B2F_T_AFFINE_INV) return 0;
                if (o.kind == B2F_OP_MADE_SEQ && !o.p[4]) return 0;
                Hmax = std::max(Hmax, o.n_hidden);
                ++n_cond;
                break;
            }
            default: return 0;
        }
    }
    if (n_cond == 0) return 0;     // purely elementwise programs: the generic kernel is already bandwidth-bound
    const int hp4 = (Hmax + 1 + 3) / 4, HP = 4 * hp4;      // + 1: hidden slot H carries the output bias
    if (hp4 > 8) return 0;
    int XS = D + 4;
    if (((XS >> 2) & 1) == 0) XS += 4;     // XS/4 odd: conflict-free 16-byte row accesses across a warp
    // staged weights of the largest layer: w1 [n_src][HP] + b1 [HP] + w2 [n_tgt*P][HP], P <= 2, n_src, n_tgt <= D
    const size_t wbuf = (size_t)D * HP + HP + (size_t)2 * D * HP;
    const size_t fixed = sizeof(float) * ((size_t)n_runs * 2 * D + 2 * D + 8 + wbuf);
    auto smem_bytes = [&](int r) { return sizeof(float) * (size_t)(kRowsThreads / 32) * 32 * r * XS + fixed; };
    const int R = 1;
    const size_t smem = smem_bytes(R);
    if (smem > 110 * 1024) return 0;
    A.n_ops = n_ops; A.D = D; A.XS = XS; A.B = B; A.flags = flags; A.n_runs = n_runs; A.wbuf_floats = (int)wbuf;
    A.x = x; A.y = y; A.log_det = log_det; A.log_prob = log_prob; A.base_loc = base_loc; A.base_log_scale = base_log_scale;
    const long long rows_per_cta = (long long)(kRowsThreads / 32) * 32 * R;
    const long long grid = (B + rows_per_cta - 1) / rows_per_cta;
    if (grid > 0x7fffffffLL) return fail(B2F_ERR_UNSUPPORTED, "b2f_flow_apply: batch too large for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    const cudaError_t ce = (flags & B2F_FLOW_MODE_PRECISE) ? launch_rows_h<0>(A, hp4, R, (unsigned)grid, smem, st)
                                                           : launch_rows_h<1>(A, hp4, R, (unsigned)grid, smem, st);
    if (ce != cudaSuccess) return fail(B2F_ERR_CUDA, "flow_rows_kernel: %s", cudaGetErrorString(ce));
    const int rc = check_launch("b2f_flow_apply (rows kernel)");
    return rc == B2F_OK ? 1 : rc;
}

}  // namespace b2f
