// TEST-ONLY host build of torchflows_b200/csrc/b2f_math.cuh (g++, -ffp-contract=off): lets the CPU test
// suite check the kernels' per-element arithmetic against the oracle without a GPU.  Never part of
// the product library.
#include "../torchflows_b200/csrc/b2f_math.cuh"

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

extern "C" {
float hm_exp_det(float t) { return exp_det(t); }

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
