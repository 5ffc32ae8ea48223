// Linear rational spline transformer (Dolatabadi et al. 2020) and the Scale transformer given materialised parameters:
// forward, inverse and backward per element.  Restates, relative to /root/reference/torchflows/bijections/finite/autoregressive:
// transformers/spline/linear_rational.py:9-182 (parameter split :96-100, compute_knots :82-90, compute_bins :67-75,
// compute_parameters :30-65, forward_1d :92-135, inverse_1d :137-182), spline/base.py:53-72 (strict in-bounds mask, identity
// tails) and transformers/linear/affine.py:160-200 (Scale).
//
// Per element the 4K parameters are [u_x (K) | u_y (K) | u_lambda (K) | u_d (K-1) | u_w0].  Only the bin that holds the
// evaluation point matters, so the kernel computes the two softmaxes once, walks the cumulative sums to find the bin, and
// evaluates the rational-linear piece from nine scalars (x_k, x_k+1, y_k, y_k+1, d_k, d_k+1, lambda_k, w0 and the point).
// Backward: the nine-scalar piece is differentiated in forward mode (dual numbers with nine tangents: exact, no hand
// derivation to get wrong), the knot / softmax / softplus / sigmoid chains behind the nine scalars by hand.
//
// Off the benchmarked path: accurate libm arithmetic throughout, tolerance-checked against the reference's
// outputs and autograd gradients (tests/golden/lrs.pt).
#pragma once
#include "b2f_math.cuh"

namespace b2f {

constexpr float kLrsMinBin = 1e-2f;              // linear_rational.py:19-20
constexpr float kLrsMinD = 1e-5f;                // linear_rational.py:21
constexpr float kLrsEps = 5e-10f;                // linear_rational.py:23
constexpr int kLrsMaxBins = 64;
constexpr int kLrsTangents = 9;                  // point, x_k, x_k+1, y_k, y_k+1, d_k, d_k+1, lambda_k, w0

// ---- dual numbers -------------------------------------------------------------------------------------------------------
template <int N> struct Dual {
    float v;
    float d[N];
};
template <int N> B2F_HD Dual<N> dual_var(float v, int i) {
    Dual<N> r;
    r.v = v;
#pragma unroll
    for (int j = 0; j < N; ++j) r.d[j] = (j == i) ? 1.0f : 0.0f;
    return r;
}
template <int N> B2F_HD Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v + b.v;
#pragma unroll
    for (int j = 0; j < N; ++j) r.d[j] = a.d[j] + b.d[j];
    return r;
}
template <int N> B2F_HD Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v - b.v;
#pragma unroll
    for (int j = 0; j < N; ++j) r.d[j] = a.d[j] - b.d[j];
    return r;
}
template <int N> B2F_HD Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v * b.v;
#pragma unroll
    for (int j = 0; j < N; ++j) r.d[j] = a.d[j] * b.v + a.v * b.d[j];
    return r;
}
template <int N> B2F_HD Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v / b.v;
    const float ib = 1.0f / b.v;
#pragma unroll
    for (int j = 0; j < N; ++j) r.d[j] = (a.d[j] - r.v * b.d[j]) * ib;
    return r;
}
template <int N> B2F_HD Dual<N> operator-(float a, const Dual<N>& b) {
    Dual<N> r; r.v = a - b.v;
#pragma unroll
    for (int j = 0; j < N; ++j) r.d[j] = -b.d[j];
    return r;
}
template <int N> B2F_HD Dual<N> operator+(const Dual<N>& a, float b) { Dual<N> r = a; r.v = a.v + b; return r; }
B2F_HD float t_log(float a) { return logf(a); }
B2F_HD float t_sqrt(float a) { return sqrtf(a); }
B2F_HD float t_val(float a) { return a; }
template <int N> B2F_HD Dual<N> t_log(const Dual<N>& a) {
    Dual<N> r; r.v = logf(a.v);
    const float ia = 1.0f / a.v;
#pragma unroll
    for (int j = 0; j < N; ++j) r.d[j] = a.d[j] * ia;
    return r;
}
template <int N> B2F_HD Dual<N> t_sqrt(const Dual<N>& a) {
    Dual<N> r; r.v = sqrtf(a.v);
    const float h = 0.5f / r.v;
#pragma unroll
    for (int j = 0; j < N; ++j) r.d[j] = a.d[j] * h;
    return r;
}
template <int N> B2F_HD float t_val(const Dual<N>& a) { return a.v; }
B2F_HD float t_rcp(float a) { return 1.0f / a; }
template <int N> B2F_HD Dual<N> t_rcp(const Dual<N>& a) {
    Dual<N> r; r.v = 1.0f / a.v;
    const float m = -r.v * r.v;
#pragma unroll
    for (int j = 0; j < N; ++j) r.d[j] = a.d[j] * m;
    return r;
}

struct LrsSel {
    float xk, xk1, yk, yk1, dk, dk1, lam, w0;
    float sx_k, sx_k1, sy_k, sy_k1;      // softmax mass below knot k / k+1 (widths, heights): the knots' softmax chain
    float arg_d0, arg_d1, u_w0;          // softplus arguments of d_k, d_k+1 and of w0
    int k;
};

B2F_HD float lrs_softplus(float x) { return (x > 20.0f) ? x : log1pf(expf(x)); }       // F.softplus defaults
B2F_HD float lrs_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// Softmax exponentials of the width / height logits and their sums (linear_rational.py:68, :86-87).
template <int NB, class H>
B2F_HD void lrs_softmax(const H& h, int nb_rt, float (&ex)[NB > 0 ? NB : kLrsMaxBins], float (&ey)[NB > 0 ? NB : kLrsMaxBins],
                        float& sum_x, float& sum_y) {
    const int nb = NB > 0 ? NB : nb_rt;
    float mx = -INFINITY, my = -INFINITY;
#pragma unroll
    for (int j = 0; j < nb; ++j) {
        ex[j] = h(j);
        ey[j] = h(j) + h(nb + j) / 100.0f;
        mx = fmaxf(mx, ex[j]); my = fmaxf(my, ey[j]);
    }
    sum_x = 0.0f; sum_y = 0.0f;
#pragma unroll
    for (int j = 0; j < nb; ++j) {
        ex[j] = expf(ex[j] - mx); ey[j] = expf(ey[j] - my);
        sum_x += ex[j]; sum_y += ey[j];
    }
}

// Knots by cumulative sums, the bin of v (searchsorted(knots, v) - 1 = #{knots < v} - 1, :102 / :148) and what the piece needs.
template <int NB, bool INV, class H>
B2F_HD void lrs_select(float v, const H& h, int nb_rt, float lo, float hi, const float (&ex)[NB > 0 ? NB : kLrsMaxBins],
                       const float (&ey)[NB > 0 ? NB : kLrsMaxBins], float sum_x, float sum_y, LrsSel& s) {
    const int nb = NB > 0 ? NB : nb_rt;
    const float c1 = (float)(1.0 - 1e-2 * (double)nb);       // Python double, cast once (linear_rational.py:69)
    const float span = hi - lo;
    float cx = 0.0f, cy = 0.0f, mass_x = 0.0f, mass_y = 0.0f;
    bool prev_below = true;
    s.xk = lo; s.yk = lo; s.xk1 = hi; s.yk1 = hi; s.k = 0;
    s.sx_k = 0.0f; s.sy_k = 0.0f; s.sx_k1 = 1.0f; s.sy_k1 = 1.0f;
#pragma unroll
    for (int j = 0; j < nb; ++j) {
        const float px = ex[j] / sum_x, py = ey[j] / sum_y;
        cx += kLrsMinBin + c1 * px;
        cy += kLrsMinBin + c1 * py;
        mass_x += px; mass_y += py;
        const bool last = (j == nb - 1);
        const float kx = last ? hi : span * cx + lo;           // knot_{j+1}; ends pinned (:73-74)
        const float ky = last ? hi : span * cy + lo;
        const bool below = (INV ? ky : kx) < v;
        const bool take = prev_below && !below;
        if (below) { s.xk = kx; s.yk = ky; s.k = j + 1; s.sx_k = mass_x; s.sy_k = mass_y; }
        if (take) { s.xk1 = kx; s.yk1 = ky; s.sx_k1 = mass_x; s.sy_k1 = mass_y; }
        prev_below = below;
    }
    if (s.k > nb - 1) s.k = nb - 1;                              // cannot happen for lo < v < hi; keeps the indices below in range
    const int k = s.k;
    s.lam = lrs_sigmoid(h(2 * nb + k));                          // :88
    const float c = kRqEdgeU;                                    // log(exp(1 - 1e-5) - 1), :22 (same constant as the RQ spline)
    s.arg_d0 = (k >= 1) ? c + h(3 * nb + k - 1) / 100.0f : 0.0f;
    s.arg_d1 = (k + 1 <= nb - 1) ? c + h(3 * nb + k) / 100.0f : 0.0f;
    s.dk = (k >= 1) ? lrs_softplus(s.arg_d0) + kLrsMinD : 1.0f;                 // :77-80, padded with 1.0
    s.dk1 = (k + 1 <= nb - 1) ? lrs_softplus(s.arg_d1) + kLrsMinD : 1.0f;
    s.u_w0 = h(4 * nb - 1);
    s.w0 = lrs_softplus(s.u_w0);                                 // :38
}

template <bool INV, class T>
B2F_HD void lrs_eval(const T& v, const T& xk, const T& xk1, const T& yk, const T& yk1, const T& dk, const T& dk1, const T& lam,
                     const T& w0, T& out, T& ld) {
    // w = w0 * sqrt(d_0 / d) with d_0 = 1, the padded edge derivative (:39-40)
    const T wk = w0 * t_sqrt(t_rcp(dk)), wk1 = w0 * t_sqrt(t_rcp(dk1));
    const T oml = 1.0f - lam;
    const T ym = (oml * wk * yk + lam * wk1 * yk1) / (oml * wk + lam * wk1);                       // :53-56
    const T wm = (lam * wk * dk + oml * wk1 * dk1) * ((xk1 - xk) / (yk1 - yk));                     // :57-63
    const T w = xk1 - xk;
    if (!INV) {
        const T phi = (v - xk) / w;                                                                   // :107
        if (t_val(phi) > t_val(lam)) {                                                                // :119-127
            const T den = wm * (1.0f - phi) + wk1 * (phi - lam);
            out = (wm * ym * (1.0f - phi) + wk1 * yk1 * (phi - lam)) / den;
            ld = t_log(oml * wm * wk1 * (yk1 - ym)) - t_log(den * den + kLrsEps) - t_log(w);
        } else {                                                                                      // :110-117
            const T den = wk * (lam - phi) + wm * phi;
            out = (wk * yk * (lam - phi) + wm * ym * phi) / den;
            ld = t_log(lam * wk * wm * (ym - yk)) - t_log(den * den + kLrsEps) - t_log(w);
        }
    } else {
        if (t_val(v) > t_val(ym)) {                                                                   // :165-173
            const T den = wk1 * (yk1 - v) + wm * (v - ym);
            out = (lam * wk1 * (yk1 - v) + wm * (v - ym)) / den * w + xk;
            ld = t_log(oml * wm * wk1 * (yk1 - ym)) - t_log(den * den + kLrsEps) + t_log(w);
        } else {                                                                                      // :155-163
            const T den = wk * (yk - v) + wm * (v - ym);
            out = (lam * wk * (yk - v)) / den * w + xk;
            ld = t_log(lam * wk * wm * (ym - yk)) - t_log(den * den + kLrsEps) + t_log(w);
        }
    }
}

template <int NB, bool INV, class H>
B2F_HD void lrs_apply(float v, const H& h, int nb_rt, float boundary, float& out, float& ld) {
    if (!(v > -boundary && v < boundary)) { out = v; ld = 0.0f; return; }            // spline/base.py:29-33,53-72
    float ex[NB > 0 ? NB : kLrsMaxBins], ey[NB > 0 ? NB : kLrsMaxBins], sum_x, sum_y;
    lrs_softmax<NB>(h, nb_rt, ex, ey, sum_x, sum_y);
    LrsSel s;
    lrs_select<NB, INV>(v, h, nb_rt, -boundary, boundary, ex, ey, sum_x, sum_y, s);
    lrs_eval<INV, float>(v, s.xk, s.xk1, s.yk, s.yk1, s.dk, s.dk1, s.lam, s.w0, out, ld);
}

// d(GZ * out + GL * ld) / d(v, parameters).  g(i, value) receives the gradient of parameter i (every i in [0, 4K) once).
template <int NB, bool INV, class H, class G>
B2F_HD void lrs_backward(float v, const H& h, int nb_rt, float boundary, float GZ, float GL, float& dv, const G& g) {
    const int nb = NB > 0 ? NB : nb_rt;
    if (!(v > -boundary && v < boundary)) {
        dv = GZ;
        for (int i = 0; i < 4 * nb; ++i) g(i, 0.0f);
        return;
    }
    float ex[NB > 0 ? NB : kLrsMaxBins], ey[NB > 0 ? NB : kLrsMaxBins], sum_x, sum_y;
    lrs_softmax<NB>(h, nb_rt, ex, ey, sum_x, sum_y);
    LrsSel s;
    lrs_select<NB, INV>(v, h, nb_rt, -boundary, boundary, ex, ey, sum_x, sum_y, s);
    typedef Dual<kLrsTangents> D;
    D out, ld;
    lrs_eval<INV, D>(dual_var<kLrsTangents>(v, 0), dual_var<kLrsTangents>(s.xk, 1), dual_var<kLrsTangents>(s.xk1, 2),
                     dual_var<kLrsTangents>(s.yk, 3), dual_var<kLrsTangents>(s.yk1, 4), dual_var<kLrsTangents>(s.dk, 5),
                     dual_var<kLrsTangents>(s.dk1, 6), dual_var<kLrsTangents>(s.lam, 7), dual_var<kLrsTangents>(s.w0, 8), out, ld);
    float gs[kLrsTangents];
#pragma unroll
    for (int i = 0; i < kLrsTangents; ++i) gs[i] = GZ * out.d[i] + GL * ld.d[i];
    dv = gs[0];
    const int k = s.k;
    const bool lo_free = k >= 1, hi_free = k + 1 <= nb - 1;        // pinned end knots carry no gradient (:73-74)
    const float c1 = (float)(1.0 - 1e-2 * (double)nb);
    const float scale = 2.0f * boundary * c1;
    const float gxk = lo_free ? gs[1] : 0.0f, gxk1 = hi_free ? gs[2] : 0.0f;
    const float gyk = lo_free ? gs[3] : 0.0f, gyk1 = hi_free ? gs[4] : 0.0f;
#pragma unroll
    for (int j = 0; j < nb; ++j) {
        // knot_m = span * sum_{i<m} (min + c1 * softmax_i) + lo:  d knot_m / d logit_j = span * c1 * p_j * ([j < m] - mass_m)
        const float px = ex[j] / sum_x, py = ey[j] / sum_y;
        const float gx = scale * px * (gxk * ((j < k ? 1.0f : 0.0f) - s.sx_k) + gxk1 * ((j < k + 1 ? 1.0f : 0.0f) - s.sx_k1));
        const float gy = scale * py * (gyk * ((j < k ? 1.0f : 0.0f) - s.sy_k) + gyk1 * ((j < k + 1 ? 1.0f : 0.0f) - s.sy_k1));
        g(j, gx + gy);                                               // height logits are u_x + u_y / 100 (:87)
        g(nb + j, gy / 100.0f);
        g(2 * nb + j, j == k ? gs[7] * s.lam * (1.0f - s.lam) : 0.0f);
    }
    for (int j = 0; j < nb - 1; ++j) {
        float gd = 0.0f;
        if (lo_free && j == k - 1) gd += gs[5] * (s.arg_d0 > 20.0f ? 1.0f : lrs_sigmoid(s.arg_d0)) / 100.0f;
        if (hi_free && j == k) gd += gs[6] * (s.arg_d1 > 20.0f ? 1.0f : lrs_sigmoid(s.arg_d1)) / 100.0f;
        g(3 * nb + j, gd);
    }
    g(4 * nb - 1, gs[8] * (s.u_w0 > 20.0f ? 1.0f : lrs_sigmoid(s.u_w0)));
}

// ---- Scale (affine.py:160-200): z = alpha * x, alpha = exp(log(1 - m) + u / 2) + m ------------------------------------------
template <int MODE> B2F_HD void scale_fwd(float x, float u0, float& z, float& ld) {
    float a; affine_scale<MODE>(u0, a, ld);
    z = a * x;
}
template <int MODE> B2F_HD void scale_inv(float z, float u0, float& x, float& ld) {
    float a, la; affine_scale<MODE>(u0, a, la);
    x = z / a;
    ld = -la;
}

}  // namespace b2f
