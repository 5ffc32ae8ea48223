// Micro-benchmarks that shape the spline epilogue (run on the B200, results in profiles/r2_ubench.txt):
// per-SMSP issue cost of the instruction classes the epilogue is made of, and TMEM read throughput.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048

template <int OP>
__global__ void k(float* out, long long* cyc, float a0, float b0) {
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = a0 + i + threadIdx.x;
    float b = b0, c = a0 * 0.5f;
    unsigned long long q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = ((unsigned long long)__float_as_uint(r[2 * i]) << 32) | __float_as_uint(r[2 * i + 1]);
    unsigned long long bb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    unsigned long long cc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(b), "f"(c));
            if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(r[i]) : "f"(b));
            if (OP == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(b));
            if (OP == 3) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(b));
            if (OP == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
            if (OP == 5) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
            if (OP == 6) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
            if (OP == 7) asm volatile("max.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(b));
            if (OP == 8) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(r[i]) : "f"(b), "f"(c));
            if (OP == 9 && i < 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(bb), "l"(cc));
            if (OP == 10 && i < 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(bb));
            if (OP == 11 && i < 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(bb));
            if (OP == 12) {   // 1 MUFU + 7 FFMA interleaved: does MUFU hide under FMA issue?
                if (i == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
                else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(b), "f"(c));
            }
            if (OP == 13) {   // 2 MUFU + 6 FFMA
                if (i < 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
                else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(b), "f"(c));
            }
            if (OP == 14) {   // FFMA + FSETP/SEL alternating (fma pipe + alu pipe)
                if (i & 1) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(r[i]) : "f"(b), "f"(c));
                else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(b), "f"(c));
            }
            if (OP == 15) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; @p mov.f32 %0, %2;}" : "+f"(r[i]) : "f"(b), "f"(c));
            if (OP == 16) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; @p add.f32 %0, %0, %2;}" : "+f"(r[i]) : "f"(b), "f"(c));
            if (OP == 17) asm volatile("max.f32 %0, %0, %1; max.f32 %0, %0, %2;" : "+f"(r[i]) : "f"(b), "f"(c));   // FMNMX3 fusion?
            if (OP == 18) {   // FFMA2 + MUFU: 1 MUFU per 3 FFMA2
                if (i == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
                else if (i < 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(bb), "l"(cc));
            }
            if (OP == 19) {   // FFMA2 and FFMA interleaved 1:1
                if (i < 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(bb), "l"(cc));
                else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(b), "f"(c));
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += __uint_as_float((unsigned)q[i]) + __uint_as_float((unsigned)(q[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// TMEM read throughput: `nw` warps each issue tcgen05.ld 32x32b.xN over 512 columns repeatedly
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int X>
__global__ void tmem_rd(float* out, long long* cyc, int reps) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll 1
        for (int c = 0; c < 512; c += 64) {
            uint32_t v[32], w[32];
            if (X == 32) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),"=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31])
                    : "r"(tb + c));
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(w[0]),"=r"(w[1]),"=r"(w[2]),"=r"(w[3]),"=r"(w[4]),"=r"(w[5]),"=r"(w[6]),"=r"(w[7]),"=r"(w[8]),"=r"(w[9]),"=r"(w[10]),"=r"(w[11]),"=r"(w[12]),"=r"(w[13]),"=r"(w[14]),"=r"(w[15]),"=r"(w[16]),"=r"(w[17]),"=r"(w[18]),"=r"(w[19]),"=r"(w[20]),"=r"(w[21]),"=r"(w[22]),"=r"(w[23]),"=r"(w[24]),"=r"(w[25]),"=r"(w[26]),"=r"(w[27]),"=r"(w[28]),"=r"(w[29]),"=r"(w[30]),"=r"(w[31])
                    : "r"(tb + c + 32));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += __uint_as_float(v[0] ^ v[31] ^ w[0] ^ w[31]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                        : "=r"(v[4*(j&3)+0]),"=r"(v[4*(j&3)+1]),"=r"(v[4*(j&3)+2]),"=r"(v[4*(j&3)+3]),"=r"(w[4*(j&3)+0]),"=r"(w[4*(j&3)+1]),"=r"(w[4*(j&3)+2]),"=r"(w[4*(j&3)+3])
                        : "r"(tb + c + 8 * j));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += __uint_as_float(v[0] ^ v[15] ^ w[0] ^ w[15]);
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512) : "memory");
}

template <int OP> void run(const char* name, int per_iter, int threads) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<OP><<<1, threads>>>(out, cyc, 1.0f, 1.0001f);
    cudaDeviceSynchronize();
    k<OP><<<1, threads>>>(out, cyc, 1.0f, 1.0001f);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const int warps_per_smsp = threads / 32 / 4;
    printf("%-34s threads=%4d  cycles/warp-instr/SMSP = %.3f   (err=%s)\n", name, threads,
           (double)c / ((double)ITERS * per_iter * warps_per_smsp), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {128, 512, 1024}) {
        run<0>("FFMA r,r,r", 8, threads);
        run<1>("FFMA r,r,imm", 8, threads);
        run<2>("FADD", 8, threads);
        run<3>("FMUL", 8, threads);
        run<4>("MUFU.EX2", 8, threads);
        run<5>("MUFU.RCP", 8, threads);
        run<6>("MUFU.LG2", 8, threads);
        run<7>("FMNMX", 8, threads);
        run<8>("FSETP+SEL (2 instr)", 8, threads);
        run<9>("FFMA2 (per instr)", 4, threads);
        run<10>("FADD2 (per instr)", 4, threads);
        run<11>("FMUL2 (per instr)", 4, threads);
        run<12>("1 EX2 + 7 FFMA (per instr)", 8, threads);
        run<13>("2 EX2 + 6 FFMA (per instr)", 8, threads);
        run<14>("FFMA / FSETP+SEL alternating (12 instr)", 8, threads);
        run<15>("FSETP + @p MOV (2 instr)", 8, threads);
        run<16>("FSETP + @p FADD (2 instr)", 8, threads);
        run<17>("max,max (FMNMX3?) per pair", 8, threads);
        run<18>("1 EX2 + 3 FFMA2 (per instr)", 4, threads);
        run<19>("4 FFMA2 + 4 FFMA (per instr)", 8, threads);
    }
    for (int threads : {128, 256, 512}) {
        float* out; long long* cyc;
        cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
        const int reps = 256;
        tmem_rd<32><<<1, threads>>>(out, cyc, reps); cudaDeviceSynchronize();
        tmem_rd<32><<<1, threads>>>(out, cyc, reps); cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double bytes = (double)reps * 512 * 4 * threads;
        printf("TMEM ld x32: threads=%d  %.1f B/clk/SM  (err=%s)\n", threads, bytes / c, cudaGetErrorString(cudaGetLastError()));
        tmem_rd<8><<<1, threads>>>(out, cyc, reps); cudaDeviceSynchronize();
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("TMEM ld x8 : threads=%d  %.1f B/clk/SM  (err=%s)\n", threads, bytes / c, cudaGetErrorString(cudaGetLastError()));
        cudaFree(out); cudaFree(cyc);
    }
    return 0;
}
