// Per-element arithmetic of the elementwise transformers (Affine / InverseAffine / Shift /
// rational-quadratic spline), forward, analytic inverse and backward.  Shared by every kernel in
// this directory.  The functions are __host__ __device__ so that tests/host_math can compile the
// very same header with g++ and check it against the oracle without a GPU; the product only ever
// calls them from device code.
//
// Reference being replaced (file:line relative to /root/reference/torchflows/bijections/finite/autoregressive):
//   transformers/linear/affine.py:33-59,149-159      Affine / Shift
//   transformers/spline/base.py:29-72                in-bounds mask, identity tails
//   transformers/spline/rational_quadratic.py:45-200 knots, bin search, RQ forward / inverse, log-det
//
// Determinism contract ("bit-exact bin indices"): everything that decides the bin index k -- the two
// softmaxes, the cumulative sums, the knot positions and the search -- uses only IEEE-754
// correctly-rounded operations (add, mul, fma, reciprocal) in a fixed order plus exp_det(), a
// polynomial exponential built from the same operations.  The same sequence is restated in plain C
// in oracle/b2f_oracle.c, so k agrees bit for bit between the CUDA kernels and the oracle.  What
// follows the search (softplus of two derivatives, three logs, the rational function) may use the
// SFU approximations (MODE >= 1); those are tolerance-checked, not bit-checked.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define B2F_HD __host__ __device__ __forceinline__
#else
#define B2F_HD inline
#endif

namespace b2f {

// ---- correctly rounded primitives (never contracted by the compiler) ---------------------------
#if defined(__CUDA_ARCH__)
B2F_HD float r_add(float a, float b) { return __fadd_rn(a, b); }
B2F_HD float r_mul(float a, float b) { return __fmul_rn(a, b); }
B2F_HD float r_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
B2F_HD float r_rcp(float a) { return __frcp_rn(a); }
B2F_HD float r_div(float a, float b) { return __fdiv_rn(a, b); }
B2F_HD int32_t f2i(float a) { return __float_as_int(a); }
B2F_HD float i2f(int32_t a) { return __int_as_float(a); }
#else
// host build: compile with -ffp-contract=off so that a*b+c is never fused behind our back
B2F_HD float r_add(float a, float b) { volatile float r = a + b; return r; }
B2F_HD float r_mul(float a, float b) { volatile float r = a * b; return r; }
B2F_HD float r_fma(float a, float b, float c) { return fmaf(a, b, c); }
B2F_HD float r_rcp(float a) { volatile float r = 1.0f / a; return r; }
B2F_HD float r_div(float a, float b) { volatile float r = a / b; return r; }
B2F_HD int32_t f2i(float a) { int32_t r; memcpy(&r, &a, 4); return r; }
B2F_HD float i2f(int32_t a) { float r; memcpy(&r, &a, 4); return r; }
#endif

// ---- deterministic exponential for t <= 0 ---------------------------------------------------------
// exp(t) = 2^n * exp(f), n = rint(t*log2e) via the 1.5*2^23 trick, f = fma(n, -ln2, t) (one Cody-Waite
// step: exact enough because the terms that matter in a softmax have |n| <= 10), degree-6 minimax
// polynomial with c0 = c1 = 1.  Error: <= 0.98 ulp for t in [-3, 0], <= 1.5 ulp down to -20, <= 4.5 ulp at
// -86 where the value is ~1e-38 relative to the maximum term 1.0 (tests/test_c_oracle_and_hostmath.py).
// Inputs below -86 are clamped (exp(-86) = 4.5e-38 is still a normal number).
B2F_HD float exp_det(float t) {
    t = fmaxf(t, -86.0f);
    const float kMagic = 12582912.0f;                       // 1.5 * 2^23
    const float r = r_fma(t, 0x1.715476p+0f, kMagic);       // low mantissa bits of r hold n
    const float n = r_add(r, -kMagic);
    const float f = r_fma(n, -0x1.62e430p-1f, t);
    float p = 0x1.6ada7ap-10f;
    p = r_fma(p, f, 0x1.127528p-7f);
    p = r_fma(p, f, 0x1.55585ep-5f);
    p = r_fma(p, f, 0x1.5554p-3f);
    p = r_fma(p, f, 0x1.fffffcp-2f);
    p = r_fma(p, f, 1.0f);
    p = r_fma(p, f, 1.0f);
    return i2f(f2i(p) + (int32_t)((uint32_t)f2i(r) << 23)); // multiply by 2^n through the exponent field
}

// ---- packed pairs: Blackwell executes two fp32 operations per lane in one FFMA2 / FADD2 / FMUL2 instruction
// (PTX fma.rn.f32x2 ...).  Each half is the same correctly rounded IEEE operation as the scalar form, so results are
// bit-identical to r_fma / r_add / r_mul; the spline evaluates its two softmaxes (widths, heights) as such pairs.
struct f2 { float x, y; };
B2F_HD f2 mk2(float x, float y) { f2 r; r.x = x; r.y = y; return r; }
B2F_HD f2 pk_fma(f2 a, f2 b, f2 c) {
#if defined(__CUDA_ARCH__)
    f2 d;
    asm("{ .reg .b64 ra, rb, rc, rd;\n\t mov.b64 ra, {%2,%3};\n\t mov.b64 rb, {%4,%5};\n\t mov.b64 rc, {%6,%7};\n\t"
        " fma.rn.f32x2 rd, ra, rb, rc;\n\t mov.b64 {%0,%1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
#else
    return mk2(r_fma(a.x, b.x, c.x), r_fma(a.y, b.y, c.y));
#endif
}
B2F_HD f2 pk_add(f2 a, f2 b) {
#if defined(__CUDA_ARCH__)
    f2 d;
    asm("{ .reg .b64 ra, rb, rd;\n\t mov.b64 ra, {%2,%3};\n\t mov.b64 rb, {%4,%5};\n\t add.rn.f32x2 rd, ra, rb;\n\t"
        " mov.b64 {%0,%1}, rd; }" : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
#else
    return mk2(r_add(a.x, b.x), r_add(a.y, b.y));
#endif
}
B2F_HD f2 pk_mul(f2 a, f2 b) {
#if defined(__CUDA_ARCH__)
    f2 d;
    asm("{ .reg .b64 ra, rb, rd;\n\t mov.b64 ra, {%2,%3};\n\t mov.b64 rb, {%4,%5};\n\t mul.rn.f32x2 rd, ra, rb;\n\t"
        " mov.b64 {%0,%1}, rd; }" : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
#else
    return mk2(r_mul(a.x, b.x), r_mul(a.y, b.y));
#endif
}
// exp_det on both halves (same operation sequence as the scalar exp_det above)
B2F_HD f2 exp_det2(f2 t) {
    t.x = fmaxf(t.x, -86.0f); t.y = fmaxf(t.y, -86.0f);
    const float kMagic = 12582912.0f;
    const f2 r = pk_fma(t, mk2(0x1.715476p+0f, 0x1.715476p+0f), mk2(kMagic, kMagic));
    const f2 n = pk_add(r, mk2(-kMagic, -kMagic));
    const f2 f = pk_fma(n, mk2(-0x1.62e430p-1f, -0x1.62e430p-1f), t);
    f2 p = mk2(0x1.6ada7ap-10f, 0x1.6ada7ap-10f);
    p = pk_fma(p, f, mk2(0x1.127528p-7f, 0x1.127528p-7f));
    p = pk_fma(p, f, mk2(0x1.55585ep-5f, 0x1.55585ep-5f));
    p = pk_fma(p, f, mk2(0x1.5554p-3f, 0x1.5554p-3f));
    p = pk_fma(p, f, mk2(0x1.fffffcp-2f, 0x1.fffffcp-2f));
    p = pk_fma(p, f, mk2(1.0f, 1.0f));
    p = pk_fma(p, f, mk2(1.0f, 1.0f));
    p.x = i2f(f2i(p.x) + (int32_t)((uint32_t)f2i(r.x) << 23));
    p.y = i2f(f2i(p.y) + (int32_t)((uint32_t)f2i(r.y) << 23));
    return p;
}

// ---- math that is tolerance-checked only -------------------------------------------------------------
// MODE 0: accurate library functions.  MODE 1: SFU approximations (ex2/lg2/rcp.approx) on the device.
template <int MODE> B2F_HD float m_exp(float x) {
#if defined(__CUDA_ARCH__)
    if (MODE >= 1) return __expf(x);
#endif
    return expf(x);
}
template <int MODE> B2F_HD float m_log(float x) {
#if defined(__CUDA_ARCH__)
    if (MODE >= 1) return __logf(x);
#endif
    return logf(x);
}
template <int MODE> B2F_HD float m_rcp(float x) {
#if defined(__CUDA_ARCH__)
    if (MODE >= 1) {                     // rcp.approx + one Newton step: ~1 ulp
        float y;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return fmaf(y, fmaf(-x, y, 1.0f), y);
    }
#endif
    return 1.0f / x;
}
template <int MODE> B2F_HD float m_sqrt(float x) {
#if defined(__CUDA_ARCH__)
    if (MODE >= 1) return __fsqrt_rn(x);
#endif
    return sqrtf(x);
}

// softplus with torch's default threshold (F.softplus: x > 20 -> x)
template <int MODE> B2F_HD float m_softplus(float x) {
    const float e = m_exp<MODE>(fminf(x, 20.0f));
    const float sp = (e < 1e-4f) ? e * (1.0f - 0.5f * e) : m_log<MODE>(1.0f + e);
    return (x > 20.0f) ? x : sp;
}

// ---- Affine / Shift (affine.py:33-59,149-159) ---------------------------------------------------------------
constexpr float kAffineM = 1e-10f;                       // min_scale, affine.py:19
constexpr float kAffineC0 = -1.00000000005e-10f;         // log(1 - 1e-10), affine.py:22

template <int MODE> B2F_HD void affine_scale(float u0, float& alpha, float& log_alpha) {
    alpha = m_exp<MODE>(kAffineC0 + 0.5f * u0) + kAffineM;   // affine.py:33-34
    log_alpha = m_log<MODE>(alpha);                          // affine.py:42 (exp then log, not u0/2)
}
template <int MODE> B2F_HD void affine_fwd(float x, float u0, float u1, float& z, float& ld) {
    float a; affine_scale<MODE>(u0, a, ld);
    z = fmaf(a, x, u1);
}
template <int MODE> B2F_HD void affine_inv(float z, float u0, float u1, float& x, float& ld) {
    float a, la; affine_scale<MODE>(u0, a, la);
    x = (z - u1) / a;
    ld = -la;
}

// ---- rational-quadratic spline -------------------------------------------------------------------------------
constexpr float kRqMinBin = 0x1.0624dep-10f;     // 1e-3   rational_quadratic.py:36
constexpr float kRqMinDelta = 0x1.4f8b58p-17f;   // 1e-5   rational_quadratic.py:37
constexpr float kRqEdgeU = 0x1.152676p-1f;       // log(expm1(1 - 1e-5))  rational_quadratic.py:38
constexpr int kRqMaxBins = 64;                   // runtime-n_bins fallback capacity

struct RqSel {           // what the evaluation needs from the bin that was found
    float xk, xk1, yk, yk1, ud0, ud1;
    int k;
};

// Knots + search.  H is any callable h(i) -> float returning parameter i of this element
// (i in [0, 3*nb-1): widths logits, heights offsets, interior derivative logits).
// NB > 0: compile-time bin count (fully unrolled, registers only); NB == 0: run-time nb <= kRqMaxBins.
// INV selects the search key: knots_x for the forward map, knots_y for the inverse (rational_quadratic.py:82,147).
// MODE 2 ("fast knots", opt-in): the softmax exponentials come from the SFU (ex2.approx) and the normalisation from
// rcp.approx + Newton; knots then agree with the deterministic ones to ~2 ulp of the boundary and the bin index can
// differ at exact ties -- tolerance-checked, not bit-checked.
template <int MODE> B2F_HD float knot_rcp(float x) {
#if defined(__CUDA_ARCH__)
    if (MODE >= 2) return m_rcp<1>(x);
#endif
    return r_rcp(x);
}

template <int MODE> B2F_HD f2 knot_exp2(f2 t) {
#if defined(__CUDA_ARCH__)
    if (MODE >= 2) return mk2(__expf(t.x), __expf(t.y));
#endif
    return exp_det2(t);
}

// Step 1 of the knots: the two softmaxes' exponentials e[] (.x widths, .y heights) and the factor g that turns them into
// bin sizes -- everything that does not depend on the evaluation point.
template <int NB, int MODE, class H>
B2F_HD void rq_knot_weights(const H& h, int nb_rt, f2 (&e)[NB > 0 ? NB : kRqMaxBins], f2& g) {
    const int nb = NB > 0 ? NB : nb_rt;
    // logits: widths u_x; heights u_x + u_y/1000 (rational_quadratic.py:75-76)
    float mx = -INFINITY, my = -INFINITY;
#pragma unroll
    for (int j = 0; j < nb; ++j) {
        const float ux = h(j);
        const float ty = r_fma(h(nb + j), 0x1.0624dep-10f, ux);
        e[j] = mk2(ux, ty);
        mx = fmaxf(mx, ux); my = fmaxf(my, ty);
    }
    const f2 negm = mk2(-mx, -my);
    f2 sum = mk2(0.0f, 0.0f);
#pragma unroll
    for (int j = 0; j < nb; ++j) {
        e[j] = knot_exp2<MODE>(pk_add(e[j], negm));
        sum = pk_add(sum, e[j]);
    }
    // sizes_j = 1e-3 + (1 - 1e-3*nb) * softmax_j   (rational_quadratic.py:46-47)
    const float c1 = (float)(1.0 - 1e-3 * (double)nb);   // Python double, cast once (rational_quadratic.py:47)
    g = mk2(r_mul(c1, knot_rcp<MODE>(sum.x)), r_mul(c1, knot_rcp<MODE>(sum.y)));
}

// Step 2: cumulative sums -> knots, and the search for the bin of v.
template <int NB, bool INV, class H>
B2F_HD void rq_search(float v, const H& h, int nb_rt, float lo, float hi, const f2 (&e)[NB > 0 ? NB : kRqMaxBins], f2 g,
                      RqSel& s) {
    const int nb = NB > 0 ? NB : nb_rt;
    const float span = r_add(hi, -lo);
    const f2 span2 = mk2(span, span), lo2 = mk2(lo, lo), minb2 = mk2(kRqMinBin, kRqMinBin);
    f2 c = mk2(0.0f, 0.0f);
    bool prev_below = true;                         // knot_0 = lo < v always (strict in-bounds test)
    s.xk = lo; s.yk = lo; s.xk1 = hi; s.yk1 = hi; s.ud0 = kRqEdgeU; s.ud1 = kRqEdgeU; s.k = 0;
#pragma unroll
    for (int j = 0; j < nb; ++j) {
        // cumsum, then (hi-lo)*c + lo as two rounded steps, ends pinned (rational_quadratic.py:48-52)
        c = pk_add(c, pk_fma(e[j], g, minb2));
        const bool last = (j == nb - 1);
        const f2 kn = pk_add(pk_mul(span2, c), lo2);
        const float kx = last ? hi : kn.x;                           // knot_{j+1}
        const float ky = last ? hi : kn.y;
        const float ud = last ? kRqEdgeU : h(2 * nb + j);           // derivative logit of knot_{j+1}
        // searchsorted(right=False) - 1  ==  #{knots < v} - 1      (rational_quadratic.py:82)
        const bool below = (INV ? ky : kx) < v;
        const bool take = prev_below && !below;                      // knot_{j+1} is the upper knot
        if (below) { s.xk = kx; s.yk = ky; s.ud0 = ud; s.k = j + 1; }
        if (take) { s.xk1 = kx; s.yk1 = ky; s.ud1 = ud; }
        prev_below = below;
    }
}

template <int NB, bool INV, int MODE, class H>
B2F_HD void rq_select(float v, const H& h, int nb_rt, float lo, float hi, RqSel& s) {
    f2 e[NB > 0 ? NB : kRqMaxBins];            // .x: widths softmax, .y: heights softmax, evaluated in lock-step
    f2 g;
    rq_knot_weights<NB, MODE>(h, nb_rt, e, g);
    rq_search<NB, INV>(v, h, nb_rt, lo, hi, e, g, s);
}

struct RqEval {          // quantities shared by value, log-det and backward
    float w, hgt, s, d0, d1, t1, xi, q, den, M;
    bool clipped;
};

// derivative at a knot: 1e-5 + softplus(c + u/1000)   (rational_quadratic.py:77)
template <int MODE> B2F_HD float rq_delta(float u) {
    return kRqMinDelta + m_softplus<MODE>(fmaf(u, 1e-3f, kRqEdgeU));
}

template <int MODE> B2F_HD float rq_logdet(const RqEval& e) {
    // 2 log s + log M - 2 log den   (rational_quadratic.py:56-63), merged into one logarithm
    const float r = e.s * m_rcp<MODE>(e.den);
    return m_log<MODE>(r * r * e.M);
}

// forward evaluation inside bin k (rational_quadratic.py:88-109)
template <int MODE> B2F_HD void rq_eval_fwd(float v, const RqSel& s, float& out, float& ld, RqEval& e) {
    // true (IEEE) divisions as in the reference: out = y_k + num/den cancels against y_k ~ -b, so an
    // ulp of num/den is worth ulp(b) in the output
    e.w = s.xk1 - s.xk; e.hgt = s.yk1 - s.yk;
    float xr;
    if (MODE >= 1) {                 // one reciprocal (rcp.approx + Newton, ~1 ulp) shared by both quotients
        const float iw = m_rcp<MODE>(e.w);
        e.s = e.hgt * iw;
        xr = (v - s.xk) * iw;
    } else {
        e.s = e.hgt / e.w;
        xr = (v - s.xk) / e.w;
    }
    e.d0 = rq_delta<MODE>(s.ud0); e.d1 = rq_delta<MODE>(s.ud1);
    e.t1 = e.d1 + e.d0 - 2.0f * e.s;
    e.xi = fminf(fmaxf(xr, 0.0f), 1.0f);
    e.clipped = (xr < 0.0f) || (xr > 1.0f);
    e.q = e.xi * (1.0f - e.xi);
    const float num = e.hgt * (e.s * e.xi * e.xi + e.d0 * e.q);
    e.den = e.s + e.t1 * e.q;
    const float om = 1.0f - e.xi;
    e.M = e.d1 * e.xi * e.xi + 2.0f * e.s * e.q + e.d0 * om * om;
    if (MODE >= 1) {                 // one reciprocal of the denominator for the value and the log-det
        const float iden = m_rcp<MODE>(e.den);
        out = fmaf(num, iden, s.yk);
        const float r = e.s * iden;
        ld = m_log<MODE>(r * r * e.M);
    } else {
        out = s.yk + num / e.den;
        ld = rq_logdet<MODE>(e);
    }
}

// inverse evaluation inside bin k (rational_quadratic.py:153-181)
template <int MODE> B2F_HD void rq_eval_inv(float v, const RqSel& s, float& out, float& ld, RqEval& e) {
    e.w = s.xk1 - s.xk; e.hgt = s.yk1 - s.yk;
    e.s = e.hgt / e.w;
    e.d0 = rq_delta<MODE>(s.ud0); e.d1 = rq_delta<MODE>(s.ud1);
    e.t1 = e.d1 + e.d0 - 2.0f * e.s;
    const float t0 = v - s.yk;
    const float t2 = e.hgt * e.d0;
    const float a = (e.hgt * e.s - t2) + t0 * e.t1;
    const float b = t2 - t0 * e.t1;
    const float c = -e.s * t0;
    const float sq = m_sqrt<MODE>(fmaxf(b * b - 4.0f * a * c, 0.0f));
    const float xr = 2.0f * c / (-b - sq);
    e.xi = fminf(fmaxf(xr, 0.0f), 1.0f);
    e.clipped = !(xr >= 0.0f && xr <= 1.0f);
    e.q = e.xi * (1.0f - e.xi);
    out = fmaf(e.xi, e.w, s.xk);
    e.den = e.s + e.t1 * e.q;
    const float om = 1.0f - e.xi;
    e.M = e.d1 * e.xi * e.xi + 2.0f * e.s * e.q + e.d0 * om * om;
    ld = -rq_logdet<MODE>(e);
}

// Full element: bounds test (strict, spline/base.py:29-33), search, evaluation.  k = -1 outside.
template <int NB, bool INV, int MODE, class H>
B2F_HD void rq_apply(float v, const H& h, int nb_rt, float boundary, float& out, float& ld, int& k) {
    if (!(v > -boundary && v < boundary)) { out = v; ld = 0.0f; k = -1; return; }
    RqSel s; RqEval e;
    rq_select<NB, INV, MODE>(v, h, nb_rt, -boundary, boundary, s);
    if (INV) rq_eval_inv<MODE>(v, s, out, ld, e); else rq_eval_fwd<MODE>(v, s, out, ld, e);
    k = s.k;
}

// out = T(v) and, on top, the log-det of the SAME spline at the point `out` -- the term the reference's sequential
// inverse reports for every dimension but the last (layers_base.py:218-223).  Identical arithmetic
// to two rq_apply calls with the same parameters, but the softmax exponentials are computed once.
template <int NB, bool INV, int MODE, class H>
B2F_HD void rq_apply_then_logdet_at_output(float v, const H& h, int nb_rt, float boundary, float& out, float& ld_at_out) {
    if (!(v > -boundary && v < boundary)) { out = v; ld_at_out = 0.0f; return; }     // then `out` is out of bounds too
    f2 e[NB > 0 ? NB : kRqMaxBins];
    f2 g;
    rq_knot_weights<NB, MODE>(h, nb_rt, e, g);
    RqSel s; RqEval ev;
    float ld;
    rq_search<NB, INV>(v, h, nb_rt, -boundary, boundary, e, g, s);
    if (INV) rq_eval_inv<MODE>(v, s, out, ld, ev); else rq_eval_fwd<MODE>(v, s, out, ld, ev);
    if (!(out > -boundary && out < boundary)) { ld_at_out = 0.0f; return; }
    float out2;
    rq_search<NB, INV>(out, h, nb_rt, -boundary, boundary, e, g, s);
    if (INV) rq_eval_inv<MODE>(out, s, out2, ld_at_out, ev); else rq_eval_fwd<MODE>(out, s, out2, ld_at_out, ev);
}

// ---- backward of the forward-direction spline ---------------------------------------------
// Inputs: v, parameters h(i), upstream GZ = dL/dout and GL = dL/dlogdet.  Outputs dL/dv and, through
// the callable G(i, value), dL/dh_i for every parameter (all 3*nb-1 are written exactly once).
// Out-of-bounds elements pass GZ through and have zero parameter gradient.
template <int NB, int MODE, class H, class G>
B2F_HD void rq_backward_fwd(float v, const H& h, int nb_rt, float boundary, float GZ, float GL, float& dv,
                            const G& gout) {
    const int nb = NB > 0 ? NB : nb_rt;
    if (!(v > -boundary && v < boundary)) {
        dv = GZ;
#pragma unroll
        for (int i = 0; i < 3 * nb - 1; ++i) gout(i, 0.0f);
        return;
    }
    const float lo = -boundary, hi = boundary;
    RqSel s; RqEval e; float out, ld;
    rq_select<NB, false, 0>(v, h, nb, lo, hi, s);
    rq_eval_fwd<MODE>(v, s, out, ld, e);
    const int k = s.k;
    const float xi = e.xi, q = e.q, sk = e.s, d0 = e.d0, d1 = e.d1, t1 = e.t1, Dn = e.den, M = e.M;
    const float iDn = 1.0f / Dn, iM = 1.0f / M, iw = 1.0f / e.w;
    const float A = sk * xi * xi + d0 * q;
    const float omq = 1.0f - 2.0f * q, omx = 1.0f - xi;
    // partials (Appendix D)
    const float out_xi = e.hgt * sk * M * iDn * iDn;
    const float out_s = e.hgt * (xi * xi * Dn - A * omq) * iDn * iDn;
    const float out_d0 = e.hgt * q * (Dn - A) * iDn * iDn;
    const float out_d1 = -e.hgt * A * q * iDn * iDn;
    const float ld_xi = (2.0f * d1 * xi + 2.0f * sk * (1.0f - 2.0f * xi) - 2.0f * d0 * omx) * iM
                        - 2.0f * t1 * (1.0f - 2.0f * xi) * iDn;
    const float ld_s = 2.0f / sk + 2.0f * q * iM - 2.0f * omq * iDn;
    const float ld_d0 = omx * omx * iM - 2.0f * q * iDn;
    const float ld_d1 = xi * xi * iM - 2.0f * q * iDn;
    float Gxi = GZ * out_xi + GL * ld_xi;
    if (e.clipped) Gxi = 0.0f;
    const float Gs = GZ * out_s + GL * ld_s;
    dv = Gxi * iw;
    const float Gxk = -Gxi * iw;
    const float Gw = -Gxi * xi * iw - Gs * sk * iw;
    const float Ghgt = GZ * A * iDn + Gs * iw;
    const float Gyk = GZ;
    const float Gd0 = GZ * out_d0 + GL * ld_d0;
    const float Gd1 = GZ * out_d1 + GL * ld_d1;
    // softmax probabilities of both logit vectors (recomputed; gradients need p, not the knots)
    float px[NB > 0 ? NB : kRqMaxBins], py[NB > 0 ? NB : kRqMaxBins];
    float mx = -INFINITY, my = -INFINITY;
#pragma unroll
    for (int j = 0; j < nb; ++j) {
        px[j] = h(j); py[j] = fmaf(h(nb + j), 1e-3f, px[j]);
        mx = fmaxf(mx, px[j]); my = fmaxf(my, py[j]);
    }
    float sx = 0.0f, sy = 0.0f;
#pragma unroll
    for (int j = 0; j < nb; ++j) {
        px[j] = exp_det(px[j] - mx); py[j] = exp_det(py[j] - my);
        sx += px[j]; sy += py[j];
    }
    const float isx = 1.0f / sx, isy = 1.0f / sy;
    const float span = hi - lo;
    const float c1 = (float)(1.0 - 1e-3 * (double)nb);
    // knots: gx_j = span*([j<k] Gxk + [j==k] Gw), gy_j likewise; J(p,g)_i = c1 p_i (g_i - sum_j p_j g_j)
    float dotx = 0.0f, doty = 0.0f;
#pragma unroll
    for (int j = 0; j < nb; ++j) {
        px[j] *= isx; py[j] *= isy;
        const float gxj = span * ((j < k ? Gxk : 0.0f) + (j == k ? Gw : 0.0f));
        const float gyj = span * ((j < k ? Gyk : 0.0f) + (j == k ? Ghgt : 0.0f));
        dotx = fmaf(px[j], gxj, dotx); doty = fmaf(py[j], gyj, doty);
    }
#pragma unroll
    for (int j = 0; j < nb; ++j) {
        const float gxj = span * ((j < k ? Gxk : 0.0f) + (j == k ? Gw : 0.0f));
        const float gyj = span * ((j < k ? Gyk : 0.0f) + (j == k ? Ghgt : 0.0f));
        const float jx = c1 * px[j] * (gxj - dotx);
        const float jy = c1 * py[j] * (gyj - doty);
        gout(j, jx + jy);                 // dL/du_x
        gout(nb + j, jy * 1e-3f);         // dL/du_y
    }
    // interior derivative logits: knot index jj = 1..nb-1 <-> parameter 2nb + jj - 1
#pragma unroll
    for (int jj = 1; jj < nb; ++jj) {
        const float u = h(2 * nb + jj - 1);
        const float a = fmaf(u, 1e-3f, kRqEdgeU);
        const float sg = 1.0f / (1.0f + m_exp<MODE>(-a));      // softplus' = sigmoid
        const float g = (jj == k ? Gd0 : 0.0f) + (jj == k + 1 ? Gd1 : 0.0f);
        gout(2 * nb + jj - 1, g * sg * 1e-3f);
    }
}

// ---- backward of the inverse-direction spline (implicit function theorem) ------------------------
// x = f^{-1}(z; h), ld_inv = -ld_f(x; h).  With upstream GX = dL/dx and GL = dL/dld_inv, f' = d f / d x and
// l_v = d ld_f / d x evaluated at the recovered x (ld_inv = -ld_f):  Gtot = GX - GL*l_v, dL/dz = Gtot / f'; the
// parameter gradients are the forward-direction expressions evaluated at x with
// (GZ, GL_f) := (-Gtot / f', -GL).
template <int NB, int MODE, class H, class G>
B2F_HD void rq_backward_inv(float z, const H& h, int nb_rt, float boundary, float GX, float GL, float& dz, const G& gout) {
    const int nb = NB > 0 ? NB : nb_rt;
    if (!(z > -boundary && z < boundary)) {
        dz = GX;
#pragma unroll
        for (int i = 0; i < 3 * nb - 1; ++i) gout(i, 0.0f);
        return;
    }
    float x, ldi; int k;
    rq_apply<NB, true, 0>(z, h, nb, boundary, x, ldi, k);
    if (!(x > -boundary && x < boundary)) {           // rounding pushed x onto the boundary: identity tail
        dz = GX;
#pragma unroll
        for (int i = 0; i < 3 * nb - 1; ++i) gout(i, 0.0f);
        return;
    }
    RqSel s; RqEval e; float out, ldf;
    rq_select<NB, false, 0>(x, h, nb, -boundary, boundary, s);
    rq_eval_fwd<0>(x, s, out, ldf, e);
    const float iDn = 1.0f / e.den, iM = 1.0f / e.M, iw = 1.0f / e.w, om = 1.0f - e.xi;
    const float fprime = e.hgt * e.s * e.M * iDn * iDn * iw;
    const float ld_xi = (2.0f * e.d1 * e.xi + 2.0f * e.s * (1.0f - 2.0f * e.xi) - 2.0f * e.d0 * om) * iM
                        - 2.0f * e.t1 * (1.0f - 2.0f * e.xi) * iDn;
    const float Gtot = GX - GL * ld_xi * iw;
    dz = Gtot / fprime;
    float dv_unused;
    rq_backward_fwd<NB, MODE>(x, h, nb, boundary, -dz, -GL, dv_unused, gout);
}

// ---- backward of Affine forward / inverse and Shift (Appendix D, last paragraph) ------------------------------
template <int MODE>
B2F_HD void affine_fwd_backward(float x, float u0, float GZ, float GL, float& dx, float& du0, float& du1) {
    float a, la; affine_scale<MODE>(u0, a, la);
    dx = GZ * a;
    du1 = GZ;
    du0 = (GZ * x + GL / a) * (a - kAffineM) * 0.5f;
}
template <int MODE>
B2F_HD void affine_inv_backward(float z, float u0, float u1, float GX, float GL, float& dz, float& du0, float& du1) {
    float a, la; affine_scale<MODE>(u0, a, la);
    const float ia = 1.0f / a;
    dz = GX * ia;
    du1 = -GX * ia;
    du0 = (-GX * (z - u1) * ia * ia - GL * ia) * (a - kAffineM) * 0.5f;
}

// standard-normal / diagonal-Gaussian log density of one coordinate (base_distributions/gaussian.py:46-54)
B2F_HD float gauss_logp(float z, float loc, float log_scale) {
    const float t = (z - loc) * expf(-log_scale);
    return -(0.5f * t * t + 0.91893853320467274178f + log_scale);
}

}  // namespace b2f
