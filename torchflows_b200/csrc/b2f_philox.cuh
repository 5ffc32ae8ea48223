// Counter-based normal noise for Flow.sample (reference: base_distributions/gaussian.py:41-44 draws torch.randn on the CPU and
// copies it; flows.py:660-713 pushes it through the inverse bijection).  Philox4x32-10 (Salmon et al., SC'11; the same
// generator family curand / torch use) keyed by a 64-bit seed: flat element e = row * D + column of the noise matrix is
// lane e % 4 of counter e / 4 + offset, so ANY kernel can regenerate any 4-aligned group of columns of any row without
// reading memory -- the fused spline kernel draws its tiles in registers (csrc/b2f_flow_tcq.cu), every other path
// materialises the same stream with philox_normal_kernel (csrc/b2f_philox.cu).
// Box-Muller on (a, b): r = sqrt(-2 ln u1), u1 = (a >> 8 + 0.5) 2^-24 in (0, 1); theta = 2 pi (b >> 8) 2^-24.
#pragma once
#include <stdint.h>

namespace b2f {
namespace philox {

__host__ __device__ __forceinline__ void round4(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[0] = n0; c[1] = (uint32_t)p1; c[2] = n2; c[3] = (uint32_t)p0;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint64_t counter, uint64_t seed, uint32_t (&out)[4]) {
    uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        round4(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    const float u1 = fmaf((float)(a >> 8), 5.9604644775390625e-08f, 2.98023223876953125e-08f);     // (k + 0.5) 2^-24
    const float th = (float)(b >> 8) * (6.283185307179586f * 5.9604644775390625e-08f);
    // SFU forms: lg2 / sqrt / sin / cos .approx (theta in [0, 2 pi): the range the approximations are specified for)
    float l2, r, s, c;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l2 * -1.3862943611198906f));      // -2 ln u1 = -2 ln2 * lg2 u1
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
    n0 = r * c;
    n1 = r * s;
}

// four standard normals: flat elements 4 * group .. 4 * group + 3 of the stream (seed, offset)
__device__ __forceinline__ float4 normal4(uint64_t group, uint64_t seed, uint64_t offset) {
    uint32_t u[4];
    philox4x32_10(group + offset, seed, u);
    float4 v;
    box_muller(u[0], u[1], v.x, v.y);
    box_muller(u[2], u[3], v.z, v.w);
    return v;
}

}  // namespace philox
}  // namespace b2f
