// Issue-rate micro-benchmark for the instruction mix of the sliding-window ZNCC kernels (sm_100a).
// Prints warp-instructions per clock per SM for scalar and packed fp32 ops, MUFU.RSQ, SHFL and LDS.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v){ float a,b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a+b; }
#define ITERS 4096
template <int MODE> __global__ void __launch_bounds__(256) k(float *out, float x, float y, long long *cyc) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = x * i;
    __syncthreads();
    float a[16]; u64 p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = x + i + threadIdx.x; p[i] = pk(a[i], a[i] + 1.f); }
    const u64 px = pk(x, x), py = pk(y, y);
    unsigned b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) b[i] = threadIdx.x * 7 + i;
    const unsigned bc = (unsigned)(x * 1000.f);
    long long t0 = clock64();
    int idx = threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], x, y);
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(px), "l"(py));
            if (MODE == 2) a[i] = a[i] + x;
            if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(px));
            if (MODE == 4) a[i] = a[i] * x;
            if (MODE == 5) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 6) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);
            if (MODE == 7) { a[i] += sm[(idx + i * 33) & 1023]; }
            if (MODE == 8) { a[i] = fmaf(a[i], x, a[(i + 1) & 15]); }            // 3 distinct regs
            if (MODE == 9) { a[i] = fmaxf(a[i], x); }
            if (MODE == 10) { if (i & 1) a[i] = fmaf(a[i], x, y); else a[i] = a[i] + x; }
            // issue-slot test: does a packed FFMA2 leave the second cycle's issue slot to another pipe?
            if (MODE == 12) { a[i] = fmaf(a[i], x, y); if (i & 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i]) : "r"(bc)); }   // 16 FFMA + 8 LOP3
            if (MODE == 13) { if (i < 8) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(px), "l"(py));
                              else asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i]) : "r"(bc)); }                                     // 8 FFMA2 + 8 LOP3
            if (MODE == 14) { if (i & 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i]) : "r"(bc)); }                               // 8 LOP3
            if (MODE == 15) { if (i < 8) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(px), "l"(py));
                              else { float4 v = *reinterpret_cast<float4 *>(&sm[((idx + i * 8) & 255) * 4]); a[i] += v.x; } }        // 8 FFMA2 + 8 (LDS.128 + FADD)
            if (MODE == 16) { a[i] = fmaf(a[i], x, y); if (i >= 8) { float4 v = *reinterpret_cast<float4 *>(&sm[((idx + i * 8) & 255) * 4]); a[i] += v.x; } }  // 16 FFMA + 8 (LDS.128 + FADD)
            if (MODE == 11) { float4 v = *reinterpret_cast<float4 *>(&sm[((idx + i * 8) & 255) * 4]); a[i] += v.x + v.y + v.z + v.w; }
        }
        if (MODE == 7 || MODE == 11) idx = (idx + 17) & 1023;
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + lo(p[i]) + (float)b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name, int instr_per_iter, int blocks_per_sm) {
    int sms = 148; float *out; long long *cyc; int nb = sms * blocks_per_sm;
    cudaMalloc(&out, (size_t)nb * 256 * 4); cudaMalloc(&cyc, nb * 8);
    k<MODE><<<nb, 256>>>(out, 1.0001f, 0.5f, cyc); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<nb, 256>>>(out, 1.0001f, 0.5f, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 8]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nb; ++i) avg += h[i]; avg /= nb;
    double winstr = (double)ITERS * instr_per_iter * 8 * blocks_per_sm;   // warp-instr per SM
    printf("%-28s blocks/SM=%d  %.3f warp-instr/clk/SM (cycles %.0f)  %.3f ms  eff_clk=%.0f MHz\n", name, blocks_per_sm,
           winstr / avg, avg, ms, avg / ms / 1e3);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int b : {1, 2, 4}) {
        if (b == 1) { run<0>("FFMA r,r,c", 16, 1); run<1>("FFMA2", 16, 1); run<2>("FADD", 16, 1); run<3>("FADD2", 16, 1); run<4>("FMUL", 16, 1);
                      run<5>("MUFU.RSQ", 16, 1); run<6>("SHFL", 16, 1); run<7>("LDS.32+FADD", 32, 1); run<8>("FFMA 3reg", 16, 1); run<9>("FMNMX", 16, 1); run<10>("FFMA/FADD mix", 16, 1); run<11>("LDS.128+4FADD", 80, 1);
                      run<12>("16 FFMA + 8 LOP3", 24, 1); run<13>("8 FFMA2 + 8 LOP3", 16, 1); run<14>("8 LOP3", 8, 1);
                      run<15>("8 FFMA2 + 8 LDS.128 + 8 FADD", 24, 1); run<16>("16 FFMA + 8 LDS.128 + 8 FADD", 32, 1); }
        if (b == 2) { run<0>("FFMA r,r,c", 16, 2); run<1>("FFMA2", 16, 2); run<2>("FADD", 16, 2); run<3>("FADD2", 16, 2); run<5>("MUFU.RSQ", 16, 2); run<6>("SHFL", 16, 2); run<8>("FFMA 3reg", 16, 2); run<10>("FFMA/FADD mix", 16, 2); run<7>("LDS.32+FADD", 32, 2); run<11>("LDS.128+4FADD", 80, 2);}
        if (b == 4) { run<0>("FFMA r,r,c", 16, 4); run<1>("FFMA2", 16, 4); run<2>("FADD", 16, 4); run<8>("FFMA 3reg", 16, 4); run<10>("FFMA/FADD mix", 16, 4); run<6>("SHFL", 16, 4); run<7>("LDS.32+FADD", 32, 4);
                      run<12>("16 FFMA + 8 LOP3", 24, 4); run<13>("8 FFMA2 + 8 LOP3", 16, 4); run<14>("8 LOP3", 8, 4);
                      run<15>("8 FFMA2 + 8 LDS.128 + 8 FADD", 24, 4); run<16>("16 FFMA + 8 LDS.128 + 8 FADD", 32, 4);}
    }
    return 0;
}
