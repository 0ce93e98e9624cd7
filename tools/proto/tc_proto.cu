// Feasibility prototype (not part of the product): the ZNCC forward as a dense contraction on the 5th-generation
// tensor cores.  For one image row, the centred window sum
//     exy[x, p] = sum_tap (cam_patch[x][tap] - mean_cam[x]) * (proj_patch[p][tap] - mean_proj[p])
// is a GEMM  CC[128 x 32] * PC[320 x 32]^T  with K = k*k = 25 taps padded to 32.  It is evaluated with three
// tcgen05.mma kind::tf32 passes (hi*hi + hi*lo + lo*hi, "3xTF32"), accumulators in TMEM, and the ZNCC epilogue
// (reference stereo_matching_kernel.cu:71) applied on the way out.  Because both patches are centred per window, the
// arithmetic is as well conditioned as the reference's two-pass sums - no pivots, no verdict, no fallback.
//
// This file answers two questions for the next round: (1) does 3xTF32 meet the 1e-5 tolerance on textured and on
// low-texture images, (2) what does a row tile cost.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int KW = 5, R = 2, TAPS = 32, MT = 128, D = 192, NP = 320;
constexpr float kEps = 1e-8f;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// K-major, no swizzle: element (row, k) of a [rows x 32] fp32 tile lives at chunk (k/4) * rows*16 B + row*16 B + (k%4)*4 B,
// so a core matrix (8 rows x 16 B) is 128 contiguous bytes, SBO (next 8 rows) = 128 B, LBO (next 16 B of K) = rows*16 B.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3ffffu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= 1ull << 46;   // descriptor version (Blackwell)
    return d;          // layout_type = 0 (no swizzle), base_offset = 0
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

// one patch row: 25 centred taps (+7 zeros) split into tf32 hi / lo parts, and the centred second moment
__device__ __forceinline__ void build_row(const float *img, int H, int W, int y, int col, float *hi, float *lo, int rows,
                                          int row, float *e2) {
    float v[TAPS];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < KW; ++i)
#pragma unroll
        for (int j = 0; j < KW; ++j) {
            const int yy = y + i - R, xx = col + j - R;
            const float t = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img[(size_t)yy * W + xx] : 0.f;
            v[i * KW + j] = t;
            sum += t;
        }
    const float mean = sum / (float)(KW * KW);
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
        v[t] = t < KW * KW ? v[t] - mean : 0.f;
        q = fmaf(v[t], v[t], q);
    }
    *e2 = q;
#pragma unroll
    for (int c = 0; c < TAPS / 4; ++c) {
        float4 h, l;
        h.x = tf32_hi(v[4 * c]); h.y = tf32_hi(v[4 * c + 1]); h.z = tf32_hi(v[4 * c + 2]); h.w = tf32_hi(v[4 * c + 3]);
        l.x = v[4 * c] - h.x; l.y = v[4 * c + 1] - h.y; l.z = v[4 * c + 2] - h.z; l.w = v[4 * c + 3] - h.w;
        *reinterpret_cast<float4 *>(hi + ((size_t)c * rows + row) * 4) = h;
        *reinterpret_cast<float4 *>(lo + ((size_t)c * rows + row) * 4) = l;
    }
}

// v2: one CTA = 128 camera columns x a band of rows.  Per row the projector range is cut into NBLK blocks of NB columns;
// a "job" is (row, block).  Iteration j: all warps build the operand tiles of job j in shared memory, one thread issues
// the MMAs of job j into TMEM buffer j&1, then all warps run the epilogue of job j-1 out of buffer (j-1)&1 while the
// tensor core works.  The epilogue thread owns one camera column (TMEM lane), so the WTA is a thread-local running
// maximum; costs are transposed through a per-warp shared-memory stage and written as contiguous runs.
constexpr int NB = 160, NBLK = NP / NB, HALF = NB / 2, STAGE_LD = HALF + 4;

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

#define TMEM_LD16(r, taddr)                                                                                                   \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),   \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                       \
                 : "r"(taddr))

__global__ void __launch_bounds__(256, 1) tc_forward(const float *__restrict__ cam, const float *__restrict__ proj,
                                                     float *__restrict__ out, float *__restrict__ best, int *__restrict__ index,
                                                     int H, int W, int RB, int passes, long long *dbg) {
    long long t_wait = 0, t_build = 0, t_mma = 0, t_math = 0, t_out = 0, t_sync = 0, tc;
#define TICK() tc = clock64()
#define TOCK(acc) acc += clock64() - tc
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *Ahi = reinterpret_cast<float *>(smem_raw);          // [8][128][4]
    float *Alo = Ahi + MT * TAPS;
    float *Bhi = Alo + MT * TAPS;                               // [8][NB][4]
    float *Blo = Bhi + NB * TAPS;
    float *stage = Blo + NB * TAPS;                             // [8 warps][32][STAGE_LD]
    float *ex2 = stage + 8 * 32 * STAGE_LD;                     // [2][128]   (row parity)
    float *ey2 = ex2 + 2 * MT;                                  // [2][NB]    (job parity)
    float *wbest = ey2 + 2 * NB;                                // [128] WTA hand-over between the two column halves
    int *wbs = reinterpret_cast<int *>(wbest + MT);             // [128]
    __shared__ __align__(8) unsigned long long mbar[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x0 = blockIdx.x * MT, h0 = blockIdx.y * RB, b = blockIdx.z;
    const int rows = min(RB, H - h0);
    cam += (size_t)b * H * W; proj += (size_t)b * H * W;
    const int p_lo = x0 - (D - 1);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);

    const int q = warp & 3, hh = warp >> 2, m = 32 * q + lane, x = x0 + m;
    float *my_stage = stage + warp * 32 * STAGE_LD;
    float bv = -INFINITY;
    int bs = 0;
    const int njobs = rows * NBLK;

    for (int j = 0; j <= njobs; ++j) {
        TICK();
        if (j > 0) mbar_wait(smem_u32(&mbar[(j - 1) & 1]), ((j - 1) >> 1) & 1);   // MMA j-1 done: operand tiles are free
        TOCK(t_wait);
        TICK();
        if (j < njobs) {
            const int y = h0 + j / NBLK, blk = j % NBLK;
            const int nrows_a = blk == 0 ? MT : 0;
            for (int r = tid; r < nrows_a + NB; r += 256) {
                if (r < nrows_a) build_row(cam, H, W, y, x0 + r, Ahi, Alo, MT, r, &ex2[((j / NBLK) & 1) * MT + r]);
                else build_row(proj, H, W, y, p_lo + blk * NB + (r - nrows_a), Bhi, Blo, NB, r - nrows_a, &ey2[(j & 1) * NB + r - nrows_a]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        TOCK(t_build);
        TICK();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        TOCK(t_sync);
        TICK();
        if (j < njobs && tid == 0) {
            const uint32_t a_hi = smem_u32(Ahi), a_lo = smem_u32(Alo), b_hi = smem_u32(Bhi), b_lo = smem_u32(Blo);
            uint32_t acc = 0;
            for (int pass = 3 - passes; pass < 3; ++pass) {
                const uint32_t a = pass == 0 ? a_lo : a_hi, bb = pass == 1 ? b_lo : b_hi;
#pragma unroll
                for (int kk = 0; kk < TAPS / 8; ++kk) {
                    mma_tf32(tmem_base + (j & 1) * 256, make_desc(a + kk * 2 * (MT * 16), MT * 16, 128),
                             make_desc(bb + kk * 2 * (NB * 16), NB * 16, 128), idesc, acc);
                    acc = 1;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[j & 1])) : "memory");
        }
        TOCK(t_mma);
        TICK();
        if (j > 0) {
            // ---- epilogue of job j-1 (its MMAs completed before this iteration's build) ----
            const int je = j - 1, y = h0 + je / NBLK, blk = je % NBLK;
            const int P0 = p_lo + blk * NB + hh * HALF;                       // projector column of this warp's first TMEM column
            const float e2 = ex2[((je / NBLK) & 1) * MT + m];
            const float *ey = ey2 + (je & 1) * NB + hh * HALF;
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (je & 1) * 256 + hh * HALF;
            if (blk == 0) { bv = -INFINITY; bs = 0; }
#pragma unroll
            for (int c0 = 0; c0 < HALF; c0 += 16) {
                uint32_t r[16];
                TMEM_LD16(r, taddr + c0);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float v[16];
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    const int p = P0 + c0 + jj, s = x - p;
                    const float val = (__uint_as_float(r[jj]) + kEps) * rsqrtf(fmaf(e2, ey[c0 + jj], kEps));
                    v[jj] = p >= 0 ? val : -2.f;
                    if (s >= 0 && s < D && p >= 0 && val > bv) { bv = val; bs = s; }
                }
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    *reinterpret_cast<float4 *>(my_stage + lane * STAGE_LD + c0 + 4 * g) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
            }
            __syncwarp();
            TOCK(t_math);
            TICK();
            // write-out: row i of the stage holds camera column x0 + 32q + i, position n <-> projector column P0 + n
            for (int i = 0; i < 32; ++i) {
                const int xr = x0 + 32 * q + i;
                if (xr >= W) break;
                const int s_hi = min(D - 1, xr - P0), s_lo = max(0, xr - P0 - (HALF - 1));
                float *orow = out + (((size_t)b * H + y) * W + xr) * D;
                for (int s = s_lo + lane; s <= s_hi; s += 32) __stcs(orow + s, my_stage[i * STAGE_LD + (xr - P0 - s)]);
            }
            __syncwarp();
            if (blk == NBLK - 1) {   // row complete: merge the two column halves of every camera column
                if (hh == 1) { wbest[m] = bv; wbs[m] = bs; }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (hh == 0 && x < W) {
                    const float ov = wbest[m];
                    const int os = wbs[m];
                    if (ov > bv || (ov == bv && os > bs)) { bv = ov; bs = os; }
                    best[((size_t)b * H + y) * W + x] = bv;
                    index[((size_t)b * H + y) * W + x] = bs;
                }
            }
        }
        TOCK(t_out);
        TICK();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        TOCK(t_sync);
    }
    if (dbg && blockIdx.x == 3 && blockIdx.y == 3 && blockIdx.z == 0 && (tid == 0 || tid == 200)) {
        long long *d = dbg + (tid ? 8 : 0);
        d[0] = t_wait; d[1] = t_build; d[2] = t_mma; d[3] = t_math; d[4] = t_out; d[5] = t_sync; d[6] = njobs;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// fp64 two-pass reference of the same cells (reference stereo_matching_kernel.cu:39-71 in double)
__global__ void ref_forward(const float *cam, const float *proj, double *out, int H, int W, int y0, int rows) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)rows * W * D) return;
    const int s = id % D, x = (id / D) % W, y = y0 + id / ((long long)D * W);
    const int p = x - s;
    if (p < 0) { out[id] = -2.0; return; }
    double cm = 0, pm = 0;
    for (int i = 0; i < KW; ++i)
        for (int j = 0; j < KW; ++j) {
            const int yy = y + i - R, xc = x + j - R, xp = p + j - R;
            const bool oky = yy >= 0 && yy < H;
            cm += (oky && xc >= 0 && xc < W) ? cam[(size_t)yy * W + xc] : 0.f;
            pm += (oky && xp >= 0 && xp < W) ? proj[(size_t)yy * W + xp] : 0.f;
        }
    cm /= KW * KW; pm /= KW * KW;
    double exy = 0, ex2 = 0, ey2 = 0;
    for (int i = 0; i < KW; ++i)
        for (int j = 0; j < KW; ++j) {
            const int yy = y + i - R, xc = x + j - R, xp = p + j - R;
            const bool oky = yy >= 0 && yy < H;
            const double c = ((oky && xc >= 0 && xc < W) ? cam[(size_t)yy * W + xc] : 0.f) - cm;
            const double q = ((oky && xp >= 0 && xp < W) ? proj[(size_t)yy * W + xp] : 0.f) - pm;
            exy += c * q; ex2 += c * c; ey2 += q * q;
        }
    out[id] = (exy + 1e-8) / sqrt(ex2 * ey2 + 1e-8);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main() {
    const int H = 375, W = 1242;
    const size_t npx = (size_t)H * W, ncell = npx * D;
    std::vector<float> hc(npx), hp(npx);
    float *cam, *proj, *out, *best;
    int *index;
    double *ref;
    const int check_rows = 24;
    CK(cudaMalloc(&cam, npx * 4)); CK(cudaMalloc(&proj, npx * 4)); CK(cudaMalloc(&out, ncell * 4)); CK(cudaMalloc(&best, npx * 4)); CK(cudaMalloc(&index, npx * 4));
    CK(cudaMalloc(&ref, (size_t)check_rows * W * D * 8));
    const size_t smem = (size_t)(2 * MT * TAPS + 2 * NB * TAPS + 8 * 32 * STAGE_LD + 2 * MT + 2 * NB + 2 * MT) * 4;
    const int RB = 25;
    std::vector<float> hb(npx);
    std::vector<int> hi(npx);
    CK(cudaFuncSetAttribute(tc_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> ho(ncell);
    std::vector<double> hr((size_t)check_rows * W * D);
    for (int family = 0; family < 3; ++family) {
        srand(1234 + family);
        for (size_t i = 0; i < npx; ++i) {
            const float u = rand() / (float)RAND_MAX, v = rand() / (float)RAND_MAX;
            const int x = i % W, y = i / W;
            if (family == 0) { hc[i] = u; hp[i] = v; }                                                     // uniform random
            else if (family == 1) { hc[i] = 0.2f + 0.6f * x / W + 0.05f * (u - 0.5f); hp[i] = 0.2f + 0.6f * x / W + 0.05f * (v - 0.5f); }  // ramp, weak texture
            else { hc[i] = 0.5f + 0.4f * sinf(x * 0.01f) * cosf(y * 0.02f) + 0.01f * (u - 0.5f); hp[i] = 0.5f + 0.4f * sinf((x + 40) * 0.01f) * cosf(y * 0.02f) + 0.01f * (v - 0.5f); }  // smooth, 1 % noise
        }
        CK(cudaMemcpy(cam, hc.data(), npx * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(proj, hp.data(), npx * 4, cudaMemcpyHostToDevice));
        for (int passes = 3; passes >= 1; passes -= 2) {
            CK(cudaMemset(out, 0xff, ncell * 4));
            dim3 grid((W + MT - 1) / MT, (H + RB - 1) / RB, 1);
            tc_forward<<<grid, 256, smem>>>(cam, proj, out, best, index, H, W, RB, passes, nullptr);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            double worst = 0; size_t bad = 0, wta_bad = 0;
            CK(cudaMemcpy(hb.data(), best, npx * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hi.data(), index, npx * 4, cudaMemcpyDeviceToHost));
            for (int y0 : {0, 180, H - check_rows}) {
                ref_forward<<<(unsigned)(((size_t)check_rows * W * D + 255) / 256), 256>>>(cam, proj, ref, H, W, y0, check_rows);
                CK(cudaMemcpy(hr.data(), ref, hr.size() * 8, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(ho.data(), out + (size_t)y0 * W * D, hr.size() * 4, cudaMemcpyDeviceToHost));
                for (size_t i = 0; i < hr.size(); ++i) {
                    const double e = fabs((double)ho[i] - hr[i]);
                    if (!(e <= 1e-5)) ++bad;
                    if (e > worst || e != e) worst = e;
                }
                for (size_t px = 0; px < (size_t)check_rows * W; ++px) {   // WTA = arg-max of the kernel's own volume, ties -> largest s
                    float m = -INFINITY; int ms = 0;
                    for (int s = D - 1; s >= 0; --s) if (ho[px * D + s] > m && ho[px * D + s] != -2.f) { m = ho[px * D + s]; ms = s; }
                    const size_t gp = (size_t)y0 * W + px;
                    if (hb[gp] != m || hi[gp] != ms) ++wta_bad;
                }
            }
            printf("family %d  %d-pass tf32: max |cost - fp64| = %.3e over %zu cells, %zu above 1e-5, %zu WTA mismatches\n", family, passes,
                   worst, 3 * hr.size(), bad, wta_bad);
        }
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dim3 grid((W + MT - 1) / MT, (H + RB - 1) / RB, 1);
    long long *dbg; CK(cudaMalloc(&dbg, 16 * 8)); CK(cudaMemset(dbg, 0, 16 * 8));
    cudaEventRecord(e0);
    for (int it = 0; it < 10; ++it) tc_forward<<<grid, 256, smem>>>(cam, proj, out, best, index, H, W, RB, 3, dbg);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long hd[16]; CK(cudaMemcpy(hd, dbg, sizeof(hd), cudaMemcpyDeviceToHost));
    for (int t = 0; t < 2; ++t)
        printf("thread %d clocks per job: wait %lld  build %lld  mma-issue %lld  math %lld  write-out %lld  sync %lld  (jobs %lld)\n", t ? 200 : 0,
               hd[8 * t] / hd[8 * t + 6], hd[8 * t + 1] / hd[8 * t + 6], hd[8 * t + 2] / hd[8 * t + 6], hd[8 * t + 3] / hd[8 * t + 6], hd[8 * t + 4] / hd[8 * t + 6], hd[8 * t + 5] / hd[8 * t + 6], hd[8 * t + 6]);
    printf("prototype v2 (banded, MMA overlapped with the epilogue, staged stores, WTA): %.3f ms per pair = %.1f Gcell/s\n", ms / 10, ncell / (ms / 10) * 1e-6);
    return 0;
}
