// TEST-ONLY host build of torchflows_b200/csrc/b2f_math.cuh (g++, -ffp-contract=off): lets the CPU test
// suite check the kernels' per-element arithmetic against the oracle without a GPU.  Never part of
// the product library.
#include "../torchflows_b200/csrc/b2f_math.cuh"
#include "../torchflows_b200/csrc/b2f_rqfast.cuh"

using namespace b2f;

struct HPtr {
    const float* p;
    float operator()(int i) const { return p[i]; }
};
struct GPtr {
    float* p;
    void operator()(int i, float v) const { p[i] = v; }
};

template <int NB>
static void rq_run(const float* x, const float* h, float* out, float* ld, int32_t* k, int64_t n, int nb, float b,
                   int inverse) {
    const int P = 3 * nb - 1;
    for (int64_t i = 0; i < n; ++i) {
        HPtr hp{h + i * P};
        int kk;
        if (inverse) rq_apply<NB, true, 0>(x[i], hp, nb, b, out[i], ld[i], kk);
        else rq_apply<NB, false, 0>(x[i], hp, nb, b, out[i], ld[i], kk);
        k[i] = kk;
    }
}

// folded-parameter spline of the tensor-core kernel (b2f_rqfast.cuh): fold the 23 raw parameters of an element exactly like
// torchflows_b200/_tcq.py folds the weights (log2e scaling, /1000, differences of the padded derivative logits)
template <bool SAFE, int NY>
static void rqfast_run(const float* x, const float* h, float* out, float* ld, int64_t n, float b, int inverse) {
    for (int64_t i = 0; i < n; ++i) {
        const float* u = h + i * 23;
        float g[24];
        for (int j = 0; j < 8; ++j) {
            g[j] = rqf::kLog2e * u[j];
            g[8 + j] = rqf::kLog2e * u[8 + j] / 1000.0f;
        }
        float a[9];
        a[0] = a[8] = kRqEdgeU / 1000.0f;
        for (int j = 1; j < 8; ++j) a[j] = u[16 + j - 1] / 1000.0f;
        for (int j = 0; j < 8; ++j) g[16 + j] = a[j + 1] - a[j];
        float ld2;
        if (inverse) rqf::inverse<SAFE, NY>(x[i], g, b, out[i], ld2);
        else rqf::forward<SAFE, NY>(x[i], g, b, out[i], ld2);
        ld[i] = ld2 * rqf::kLn2;
    }
}

template <bool SAFE, int NY>
static void rqfast_g_run(const float* x, const float* g, float* out, float* ld2, int64_t n, float b, int inverse) {
    for (int64_t i = 0; i < n; ++i) {
        float gg[24];
        for (int j = 0; j < 24; ++j) gg[j] = g[i * 24 + j];
        if (inverse) rqf::inverse<SAFE, NY>(x[i], gg, b, out[i], ld2[i]);
        else rqf::forward<SAFE, NY>(x[i], gg, b, out[i], ld2[i]);
    }
}

extern "C" {
// folded columns given directly (what the GEMM of the tensor-core kernel produces); ld2 in log2 units
void hm_rqfast_g(const float* x, const float* g, float* out, float* ld2, int64_t n, float b, int inverse, int safe) {
    if (safe) rqfast_g_run<true, 0>(x, g, out, ld2, n, b, inverse);
    else rqfast_g_run<false, 2>(x, g, out, ld2, n, b, inverse);
}

void hm_rqfast(const float* x, const float* h, float* out, float* ld, int64_t n, float b, int inverse, int safe, int ny) {
    if (safe) rqfast_run<true, 0>(x, h, out, ld, n, b, inverse);
    else if (ny == 0) rqfast_run<false, 0>(x, h, out, ld, n, b, inverse);
    else if (ny == 3) rqfast_run<false, 4>(x, h, out, ld, n, b, inverse);
    else rqfast_run<false, 8>(x, h, out, ld, n, b, inverse);
}

float hm_exp_det(float t) { return exp_det(t); }

// fast backward of the wide-conditioner kernel (b2f_rqfast.cuh rqf::backward_fwd): h (n, 23) raw parameters
void hm_rq_backward_fast(const float* x, const float* h, const float* gz, const float* gl, float* dx, float* dh, int64_t n, float b) {
    for (int64_t i = 0; i < n; ++i) {
        float p[24], dp[24];
        for (int j = 0; j < 23; ++j) p[j] = h[i * 23 + j];
        p[23] = 0.0f;
        rqf::backward_fwd(x[i], p, b, gz[i], gl[i], dx[i], dp);
        for (int j = 0; j < 23; ++j) dh[i * 23 + j] = dp[j];
    }
}

void hm_rq(const float* x, const float* h, float* out, float* ld, int32_t* k, int64_t n, int nb, float b, int inverse,
           int templated) {
    if (templated && nb == 8) rq_run<8>(x, h, out, ld, k, n, nb, b, inverse);
    else if (templated && nb == 4) rq_run<4>(x, h, out, ld, k, n, nb, b, inverse);
    else rq_run<0>(x, h, out, ld, k, n, nb, b, inverse);
}

void hm_rq_backward(const float* x, const float* h, const float* gz, const float* gl, float* dv, float* dh, int64_t n,
                    int nb, float b) {
    const int P = 3 * nb - 1;
    for (int64_t i = 0; i < n; ++i) {
        HPtr hp{h + i * P};
        GPtr gp{dh + i * P};
        if (nb == 8) rq_backward_fwd<8, 0>(x[i], hp, nb, b, gz[i], gl[i], dv[i], gp);
        else rq_backward_fwd<0, 0>(x[i], hp, nb, b, gz[i], gl[i], dv[i], gp);
    }
}

void hm_rq_backward_inv(const float* z, const float* h, const float* gx, const float* gl, float* dz, float* dh, int64_t n,
                        int nb, float b) {
    const int P = 3 * nb - 1;
    for (int64_t i = 0; i < n; ++i) {
        HPtr hp{h + i * P};
        GPtr gp{dh + i * P};
        if (nb == 8) rq_backward_inv<8, 0>(z[i], hp, nb, b, gx[i], gl[i], dz[i], gp);
        else rq_backward_inv<0, 0>(z[i], hp, nb, b, gx[i], gl[i], dz[i], gp);
    }
}

void hm_affine(const float* x, const float* h, float* out, float* ld, int64_t n, int inverse) {
    for (int64_t i = 0; i < n; ++i) {
        if (inverse) affine_inv<0>(x[i], h[2 * i], h[2 * i + 1], out[i], ld[i]);
        else affine_fwd<0>(x[i], h[2 * i], h[2 * i + 1], out[i], ld[i]);
    }
}

void hm_affine_backward(const float* x, const float* h, const float* gz, const float* gl, float* dx, float* dh,
                        int64_t n, int inverse) {
    for (int64_t i = 0; i < n; ++i) {
        if (inverse) affine_inv_backward<0>(x[i], h[2 * i], h[2 * i + 1], gz[i], gl[i], dx[i], dh[2 * i], dh[2 * i + 1]);
        else affine_fwd_backward<0>(x[i], h[2 * i], gz[i], gl[i], dx[i], dh[2 * i], dh[2 * i + 1]);
    }
}
}
