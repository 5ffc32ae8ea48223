/* Plain-C CPU oracle for the elementwise transformers of torchflows v1.2.0 (scalar, single thread).
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ and by __graft_entry__.smoke() as the checker.  Never
 * linked into or loaded by the product library.
 *
 * What it restates (file:line relative to /root/reference/torchflows/bijections/finite/autoregressive):
 *   transformers/spline/rational_quadratic.py:45-54   compute_bins: softmax -> affine -> cumsum -> pad
 *                                                      -> scale -> pin ends -> sizes
 *   transformers/spline/rational_quadratic.py:65-110  rqs_forward_1d
 *   transformers/spline/rational_quadratic.py:130-182 rqs_inverse_1d
 *   transformers/spline/base.py:29-72                 strict bounds mask, identity tails
 *   transformers/linear/affine.py:33-59,149-159       Affine, Shift
 *
 * It is written in the reference's array order (all bins first, then searchsorted, then gathers).
 * The arithmetic that decides the bin index uses only correctly rounded IEEE-754 operations and the
 * polynomial exponential oexp() below, in this fixed association:
 *      t_j    = u_x[j]                      (widths)   |   fma(u_y[j], 1e-3f, u_x[j])   (heights)
 *      E_j    = oexp(t_j - max_j t_j);  S = ((0 + E_0) + E_1) + ...
 *      size_j = fma(E_j, c1 * (1/S), 1e-3f)            c1 = (float)(1 - 1e-3*n_bins)
 *      c_j    = c_{j-1} + size_j
 *      knot_{j+1} = (2b * c_j) + (-b)      two rounded steps;  knot_0 = -b, knot_K = +b pinned
 * which is the specification the CUDA kernels implement as well (torchflows_b200/csrc/b2f_math.cuh),
 * so bin indices can be compared bit for bit.  Against the reference (torch CPU: Sleef expf, vector
 * reduction order) the knots agree to an ulp or two; tests/test_oracle_golden.py pins this file
 * against the reference's own outputs (values <= 1e-5, bin indices equal away from exact ties).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared oracle/b2f_oracle.c -o oracle/_build/libb2f_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MIN_BIN 0x1.0624dep-10f   /* 1e-3  rational_quadratic.py:36 */
#define MIN_DELTA 0x1.4f8b58p-17f /* 1e-5  rational_quadratic.py:37 */
#define EDGE_U 0x1.152676p-1f     /* log(expm1(1-1e-5))  rational_quadratic.py:38 */
#define MAX_BINS 64

static float oexp(float t) {
    /* exp for t <= 0: n = rint(t*log2e), f = t - n*ln2 (one fma), degree-6 polynomial, exponent add */
    if (t < -86.0f) t = -86.0f;
    const float magic = 12582912.0f;
    float r = fmaf(t, 0x1.715476p+0f, magic);
    float n = r - magic;
    float f = fmaf(n, -0x1.62e430p-1f, t);
    float p = 0x1.6ada7ap-10f;
    p = fmaf(p, f, 0x1.127528p-7f);
    p = fmaf(p, f, 0x1.55585ep-5f);
    p = fmaf(p, f, 0x1.5554p-3f);
    p = fmaf(p, f, 0x1.fffffcp-2f);
    p = fmaf(p, f, 1.0f);
    p = fmaf(p, f, 1.0f);
    uint32_t pb, rb;
    memcpy(&pb, &p, 4);
    memcpy(&rb, &r, 4);
    pb += rb << 23;
    memcpy(&p, &pb, 4);
    return p;
}

/* compute_bins (rational_quadratic.py:45-54): logits t[0..nb) -> knots[0..nb], sizes[0..nb) */
static void compute_bins(const float *t, int nb, float lo, float hi, float *knots, float *sizes) {
    float m = t[0];
    for (int j = 1; j < nb; ++j) m = t[j] > m ? t[j] : m;
    float e[MAX_BINS], sum = 0.0f;
    for (int j = 0; j < nb; ++j) { e[j] = oexp(t[j] - m); sum = sum + e[j]; }
    const float c1 = (float)(1.0 - 1e-3 * (double)nb);
    const float g = c1 * (1.0f / sum);
    const float span = hi - lo;
    float c = 0.0f;
    knots[0] = lo;
    for (int j = 0; j < nb; ++j) {
        c = c + fmaf(e[j], g, MIN_BIN);     /* cumsum of (min_bin + c1 * softmax) */
        float scaled = span * c;            /* (maximum - minimum) * bins ... */
        knots[j + 1] = scaled + lo;         /* ... + minimum */
    }
    knots[0] = lo;
    knots[nb] = hi;
    for (int j = 0; j < nb; ++j) sizes[j] = knots[j + 1] - knots[j];
}

static float softplus(float x) { return x > 20.0f ? x : log1pf(expf(x)); } /* F.softplus defaults */

static float log_det(float s, float d0, float d1, float xi, float q, float t1) {
    /* rational_quadratic.py:56-63 */
    float ln = 2.0f * logf(s) + logf(d1 * xi * xi + 2.0f * s * q + d0 * (1.0f - xi) * (1.0f - xi));
    float ld = 2.0f * logf(s + t1 * q);
    return ln - ld;
}

/* One element.  h: 3*nb-1 parameters.  inverse = 0: rqs_forward_1d, 1: rqs_inverse_1d.
 * Outside (-b, b) the element is returned unchanged with log-det 0 and k = -1 (spline/base.py:53-72). */
static void rq_element(float v, const float *h, int nb, float b, int inverse, float *out, float *ld, int *k_out) {
    if (!(v > -b && v < b)) { *out = v; *ld = 0.0f; *k_out = -1; return; }
    float tx[MAX_BINS], ty[MAX_BINS], bin_x[MAX_BINS + 1], bin_y[MAX_BINS + 1], w[MAX_BINS], hg[MAX_BINS];
    float delta[MAX_BINS + 1];
    for (int j = 0; j < nb; ++j) { tx[j] = h[j]; ty[j] = fmaf(h[nb + j], MIN_BIN, h[j]); }
    compute_bins(tx, nb, -b, b, bin_x, w);
    compute_bins(ty, nb, -b, b, bin_y, hg);
    for (int j = 0; j <= nb; ++j) {
        float u = (j == 0 || j == nb) ? EDGE_U : h[2 * nb + j - 1];  /* F.pad with the edge constant */
        delta[j] = MIN_DELTA + softplus(fmaf(u, 1e-3f, EDGE_U));
    }
    const float *key = inverse ? bin_y : bin_x;
    int cnt = 0;                                   /* searchsorted(right=False): #knots < v */
    for (int j = 0; j <= nb; ++j) cnt += key[j] < v;
    int k = cnt - 1;
    float y_k = bin_y[k], x_k = bin_x[k], h_k = hg[k], w_k = w[k], d0 = delta[k], d1 = delta[k + 1];
    float s = h_k / w_k;
    float t1 = d1 + d0 - 2.0f * s;
    float xi;
    if (!inverse) {
        xi = (v - x_k) / w_k;
        xi = xi < 0.0f ? 0.0f : (xi > 1.0f ? 1.0f : xi);
        float q = xi * (1.0f - xi);
        float num = h_k * (s * xi * xi + d0 * q);
        float den = s + t1 * q;
        *out = y_k + num / den;
        *ld = log_det(s, d0, d1, xi, q, t1);
    } else {
        float t0 = v - y_k, t2 = h_k * d0;
        float a = (h_k * s - t2) + t0 * t1;
        float bq = t2 - t0 * t1;
        float c = -s * t0;
        float disc = bq * bq - 4.0f * a * c;
        float sq = disc > 0.0f ? sqrtf(disc) : 0.0f;
        xi = 2.0f * c / (-bq - sq);
        xi = xi < 0.0f ? 0.0f : (xi > 1.0f ? 1.0f : xi);
        float q = xi * (1.0f - xi);
        *out = xi * w_k + x_k;
        *ld = -log_det(s, d0, d1, xi, q, t1);
    }
    *k_out = k;
}

/* x, out, ld_elem: n elements; h: n * (3*nb-1); k: n ints.  ld_elem is per element (not row-summed). */
void b2f_oracle_rq(const float *x, const float *h, float *out, float *ld_elem, int32_t *k, int64_t n, int nb,
                   float boundary, int inverse) {
    const int P = 3 * nb - 1;
    for (int64_t i = 0; i < n; ++i) {
        int kk;
        rq_element(x[i], h + i * P, nb, boundary, inverse, out + i, ld_elem + i, &kk);
        k[i] = kk;
    }
}

/* knots only, for diagnostics: u (n*nb logits) -> knots (n*(nb+1)) */
void b2f_oracle_rq_knots(const float *u, float *knots, int64_t n, int nb, float boundary) {
    float sizes[MAX_BINS];
    for (int64_t i = 0; i < n; ++i) compute_bins(u + i * nb, nb, -boundary, boundary, knots + i * (nb + 1), sizes);
}

/* Affine (affine.py:33-59): h = (u_alpha, u_beta) per element.  inverse = 0: z = a*x + b; 1: (z-b)/a. */
void b2f_oracle_affine(const float *x, const float *h, float *out, float *ld_elem, int64_t n, int inverse) {
    const float m = 1e-10f, c0 = -1.00000000005e-10f;
    for (int64_t i = 0; i < n; ++i) {
        float a = expf(c0 + h[2 * i] / 2.0f) + m;
        float la = logf(a);
        if (!inverse) { out[i] = a * x[i] + h[2 * i + 1]; ld_elem[i] = la; }
        else { out[i] = (x[i] - h[2 * i + 1]) / a; ld_elem[i] = -la; }
    }
}

/* Shift (affine.py:149-159): sign = +1 forward, -1 inverse; log-det 0. */
void b2f_oracle_shift(const float *x, const float *h, float *out, int64_t n, int sign) {
    for (int64_t i = 0; i < n; ++i) out[i] = sign > 0 ? x[i] + h[i] : x[i] - h[i];
}

float b2f_oracle_exp(float t) { return oexp(t); }
