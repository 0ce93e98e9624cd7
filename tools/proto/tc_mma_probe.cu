// Probe: tcgen05.mma kind::tf32 with the A operand in TMEM and an MN-major, unswizzled B operand in shared memory
// (the second GEMM of the tensor-core backward).  Fills A and B with small integers and compares D with the host.
// usage: tc_mma_probe [sbo] [lbo] [b_major_bit]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int MT = 128, NT = 192, TAPS = 32, CH = 8;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | (uint64_t)(lbo >> 4) << 16 | (uint64_t)(sbo >> 4) << 32 | 1ull << 46 | (uint64_t)layout << 61;
}
__host__ __device__ inline float fa(int m, int k) { return (float)((m * 3 + k * 7) % 11 - 5); }
__host__ __device__ inline float fb(int k, int n) { return (float)((k * 5 + n * 13) % 7 - 3); }

__global__ void __launch_bounds__(128) probe(float *out, int sbo, int lbo, int bmaj, int ksteps, int mode) {
    extern __shared__ __align__(1024) float Bt[];   // [CH][NT][4]: element (k = row, n = tap) at chunk n/4, row k, n%4
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (mode == 0) {
        for (int i = tid; i < CH * NT * 4; i += 128) {
            const int c = i / (NT * 4), k = (i / 4) % NT, n = 4 * c + (i % 4);
            Bt[i] = fb(k, n);
        }
    } else if (mode == 3) {
        for (int i = tid; i < NT * 32; i += 128) {
            const int r = i / 32, t = i % 32;
            Bt[r * 32 + (((t >> 3) ^ (r & 3)) << 3) + (t & 7)] = fb(r, t);
        }
    } else {
        // 128-byte swizzled rows: row = projector column (192 of them), 32 taps per row; 16-byte chunk c of row r sits at c ^ (r & 7)
        for (int i = tid; i < NT * 32; i += 128) {
            const int r = i / 32, t = i % 32;
            Bt[r * 32 + (((t >> 2) ^ (r & 7)) << 2) + (t & 3)] = fb(r, t);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_s, lane_addr = tm + ((uint32_t)(32 * warp) << 16);
    for (int c0 = 0; c0 < NT; c0 += 4) {
        uint32_t r[4];
        for (int j = 0; j < 4; ++j) r[j] = __float_as_uint(fa(tid, c0 + j));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lane_addr + c0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)bmaj << 16) | ((uint32_t)(TAPS >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
        if (mode == 2) {
            // GEMM1 style: D[128 x 192] = A[128 x 32] * B[192 x 32]^T, B K-major with 128-byte swizzle
            const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
            for (int kk = 0; kk < 4; ++kk) {
                const uint64_t bd = make_desc(smem_u32(Bt) + kk * 32, lbo, sbo, 2);
                const uint32_t acc = kk > 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                             ::"r"(tm + 256), "r"(tm + kk * 8), "l"(bd), "r"(idesc1), "r"(acc) : "memory");
            }
        } else
        for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t bd = mode == 0 ? make_desc(smem_u32(Bt) + kk * 128, lbo, sbo) : make_desc(smem_u32(Bt) + kk * 1024, lbo, sbo, mode == 3 ? 1 : 2);
            const uint32_t acc = kk > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                         ::"r"(tm + 256), "r"(tm + kk * 8), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t d[32];
    for (int c0 = 0; c0 < 32; c0 += 4) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(d[c0]), "=r"(d[c0 + 1]), "=r"(d[c0 + 2]), "=r"(d[c0 + 3]) : "r"(lane_addr + 256 + c0));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[tid * 32 + j] = __uint_as_float(d[j]);
    // read A back as well
    uint32_t a0[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0[0]), "=r"(a0[1]), "=r"(a0[2]), "=r"(a0[3]) : "r"(lane_addr + 8));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 4; ++j) out[128 * 32 + tid * 4 + j] = __uint_as_float(a0[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

int main(int argc, char **argv) {
    const int sbo = argc > 1 ? atoi(argv[1]) : NT * 16, lbo = argc > 2 ? atoi(argv[2]) : 128, bmaj = argc > 3 ? atoi(argv[3]) : 1;
    const int ksteps = argc > 4 ? atoi(argv[4]) : 24;
    const int mode = argc > 5 ? atoi(argv[5]) : 0;
    float *out;
    cudaMalloc(&out, (128 * 32 + 128 * 4) * 4);
    cudaMemset(out, 0, (128 * 32 + 128 * 4) * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * NT * 16);
    probe<<<1, 128, CH * NT * 16>>>(out, sbo, lbo, bmaj, ksteps, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    static float h[128 * 32 + 128 * 4];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0, a_bad = 0;
    for (int m = 0; m < 128; ++m) {
        for (int n = 0; n < 32; ++n) {
            float ref = 0;
            if (mode == 2) { for (int k = 0; k < 32; ++k) ref += fa(m, k) * fb(n, k); }
            else for (int k = 0; k < 8 * ksteps; ++k) ref += fa(m, k) * fb(k, n);
            if (h[m * 32 + n] != ref) ++bad;
        }
        for (int j = 0; j < 4; ++j) if (h[128 * 32 + m * 4 + j] != fa(m, 8 + j)) ++a_bad;
    }
    printf("mode=%d sbo=%d lbo=%d b_major=%d ksteps=%d: %d of 4096 D entries wrong, %d of 512 A read-backs wrong;  D[0][0..3] = %g %g %g %g, D[5][7] = %g\n", mode, sbo, lbo,
           bmaj, ksteps, bad, a_bad, h[0], h[1], h[2], h[3], h[5 * 32 + 7]);
    if (mode == 2) return 0;
    float r0 = 0, r57 = 0;
    for (int k = 0; k < 8 * ksteps; ++k) { r0 += fa(0, k) * fb(k, 0); r57 += fa(5, k) * fb(k, 7); }
    printf("   expected D[0][0] = %g, D[5][7] = %g\n", r0, r57);
    return 0;
}
