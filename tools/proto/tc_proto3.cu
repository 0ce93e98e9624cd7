// Feasibility prototype v3 (not part of the product): ZNCC forward + WTA as a 3xTF32 tcgen05 contraction, organised
// the way a product kernel would be.  See tools/tc_proto.cu for the idea and the first numerics check.
//
//   CTA            128 camera columns x a band of rows; 16 worker warps + 1 MMA-issuing warp
//   job            (row y, projector block blk): D[128 x 192] = CC[128 x 32] * PC[192 x 32]^T, three tf32 passes
//   iteration j    workers: wait MMA j-1 | build operand tiles of job j from a 6-row image ring in shared memory |
//                  arrive on ops_bar | epilogue of job j-1 out of TMEM buffer (j-1)&1 | staged, coalesced stores
//                  MMA warp: wait ops_bar | 12 x tcgen05.mma into TMEM buffer j&1 | commit -> mma_bar[j&1]
//   TMEM lane L    camera column x0 + 4*(L%32) + L/32: all 32 lanes of a warp share x mod 4, so one warp-uniform
//                  column shift makes every thread's accumulator registers start at a disparity s = 0 (mod 4) and
//                  128-bit shared / global accesses line up although s = x - p shears the tile
//   WTA            the epilogue thread owns one camera column: a thread-local running maximum, no shuffles, no atomics
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int KW = 5, R = 2, TAPS = 32, MT = 128, D = 192;
constexpr int PW = 176, NT = 192, NQ = 44, SLD = 180;            // processed / computed columns per job, per warp, stage pitch
constexpr int NBLK = (D + 131 + PW - 1) / PW;                     // projector blocks per row
constexpr int CAMW = 136, PRW = 376, RING = 6;                    // image-row ring pitches
constexpr int P_SPAN = (NBLK - 1) * PW + NT - 1;                  // P_top + 4 - P_SPAN = lowest projector column built
constexpr int NWORK = 512, NTHREADS = NWORK + 32;
constexpr float kEps = 1e-8f;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | (uint64_t)(lbo_bytes >> 4) << 16 | (uint64_t)(sbo_bytes >> 4) << 32 | 1ull << 46;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
#define TMEM_LD(NREG, r, taddr) tmem_ld##NREG(r, taddr)
__device__ __forceinline__ void tmem_ld16(uint32_t *r, uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t *r, uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t *r, uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}

struct Smem {
    float Ahi[TAPS / 4][MT][4], Alo[TAPS / 4][MT][4];
    float Bhi[TAPS / 4][NT][4], Blo[TAPS / 4][NT][4];
    float stage[MT][SLD];
    float camring[RING][CAMW], prjring[RING][PRW];
    float ex2[2][MT];
    float ey2[2][4][NT + 4];          // [job parity][shift][column]: ey2[.][f][c] = second moment of column c + f
    float wv[4][MT];
    int ws[4][MT];
    unsigned long long ops_bar, mma_bar[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ void load_ring_row(Smem &S, const float *cam, const float *proj, int H, int W, int yy, int x0, int tid) {
    const int slot = (yy + RING) % RING;
    const bool row_ok = yy >= 0 && yy < H;
    for (int i = tid; i < 132 + 372; i += NWORK) {
        if (i < 132) {
            const int xc = x0 - 2 + i;
            S.camring[slot][i] = (row_ok && xc >= 0 && xc < W) ? __ldg(cam + (size_t)yy * W + xc) : 0.f;
        } else {
            const int pc = x0 - 238 + (i - 132);
            S.prjring[slot][i - 132] = (row_ok && pc >= 0 && pc < W) ? __ldg(proj + (size_t)yy * W + pc) : 0.f;
        }
    }
}

// one patch row out of the ring: 25 centred taps split into tf32 hi / lo parts (chunks 0..6; chunk 7 stays zero)
template <int ROWS>
__device__ __forceinline__ float build_patch(const float *ring, int pitch, int y, int col0, float (*hi)[ROWS][4], float (*lo)[ROWS][4], int row) {
    float v[28];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < KW; ++i) {
        const float *src = ring + ((y + i + RING - R) % RING) * pitch + col0;
#pragma unroll
        for (int j = 0; j < KW; ++j) { v[i * KW + j] = src[j]; sum += v[i * KW + j]; }
    }
    const float mean = sum / (float)(KW * KW);
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < 28; ++t) {
        v[t] = t < KW * KW ? v[t] - mean : 0.f;
        q = fmaf(v[t], v[t], q);
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        float4 h, l;
        h.x = tf32_hi(v[4 * c]); h.y = tf32_hi(v[4 * c + 1]); h.z = tf32_hi(v[4 * c + 2]); h.w = tf32_hi(v[4 * c + 3]);
        l.x = v[4 * c] - h.x; l.y = v[4 * c + 1] - h.y; l.z = v[4 * c + 2] - h.z; l.w = v[4 * c + 3] - h.w;
        *reinterpret_cast<float4 *>(hi[c][row]) = h;
        *reinterpret_cast<float4 *>(lo[c][row]) = l;
    }
    return q;
}

__global__ void __launch_bounds__(NTHREADS, 1) tc_forward(const float *__restrict__ cam, const float *__restrict__ proj,
                                                          float *__restrict__ out, float *__restrict__ best, int *__restrict__ index,
                                                          int H, int W, int RB, long long *dbg) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    long long t_wait = 0, t_build = 0, t_math = 0, t_bar = 0, t_out = 0, tc;
#define TICK() tc = clock64()
#define TOCK(acc) acc += clock64() - tc

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x0 = blockIdx.x * MT, h0 = blockIdx.y * RB, b = blockIdx.z;
    const int rows = min(RB, H - h0), njobs = rows * NBLK;
    cam += (size_t)b * H * W; proj += (size_t)b * H * W;
    const int P_top4 = x0 + 127 + 4;                        // projector column of TMEM column 0 of block 0

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&S.ops_bar)), "r"(NWORK));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.mma_bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.mma_bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    // zero the operand tiles once (the K padding, chunk 7, is never written again)
    for (int i = tid; i < (int)(sizeof(S.Ahi) + sizeof(S.Alo) + sizeof(S.Bhi) + sizeof(S.Blo)) / 16; i += NTHREADS)
        reinterpret_cast<float4 *>(smem_raw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < NWORK)
        for (int yy = h0 - R; yy <= h0 + R; ++yy) load_ring_row(S, cam, proj, H, W, yy, x0, tid);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = S.tmem_base;

    if (warp == NWORK / 32) {
        // ================= MMA warp =================
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
        const uint32_t a_hi = smem_u32(S.Ahi), a_lo = smem_u32(S.Alo), b_hi = smem_u32(S.Bhi), b_lo = smem_u32(S.Blo);
        for (int j = 0; j < njobs; ++j) {
            mbar_wait(smem_u32(&S.ops_bar), j & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                uint32_t acc = 0;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {        // small terms first: lo*hi, hi*lo, hi*hi
                    const uint32_t a = pass == 0 ? a_lo : a_hi, bb = pass == 1 ? b_lo : b_hi;
#pragma unroll
                    for (int kk = 0; kk < TAPS / 8; ++kk) {
                        mma_tf32(tmem_base + (j & 1) * 256, make_desc(a + kk * 2 * (MT * 16), MT * 16, 128),
                                 make_desc(bb + kk * 2 * (NT * 16), NT * 16, 128), idesc, acc);
                        acc = 1;
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&S.mma_bar[j & 1])) : "memory");
            }
            __syncwarp();
        }
    } else {
        // ================= worker warps =================
        const int q = warp & 3, cq = warp >> 2, L = 32 * q + lane, mp = 4 * lane + q, x = x0 + mp;
        const int phi = (3 - q) & 3, col_first = phi + NQ * cq;
        float bv = -INFINITY;
        int bs = 0;
        for (int j = 0; j <= njobs; ++j) {
            TICK();
            if (j > 0) mbar_wait(smem_u32(&S.mma_bar[(j - 1) & 1]), ((j - 1) >> 1) & 1);   // MMA j-1 done
            TOCK(t_wait);
            TICK();
            if (j < njobs) {
                const int y = h0 + j / NBLK, blk = j % NBLK;
                const int nA = blk == 0 ? MT : 0;
                if (tid < nA) {
                    const int col0 = 4 * (tid & 31) + (tid >> 5);          // camera column x0 + col0, ring origin x0 - 2
                    S.ex2[(j / NBLK) & 1][tid] = build_patch<MT>(&S.camring[0][0], CAMW, y, col0, S.Ahi, S.Alo, tid);
                } else if (tid < nA + NT) {
                    const int n = tid - nA;
                    const int col0 = P_SPAN - blk * PW - n;                // projector column P_top4 - (blk*PW + n), ring origin x0 - 238
                    const float e = build_patch<NT>(&S.prjring[0][0], PRW, y, col0, S.Bhi, S.Blo, n);
#pragma unroll
                    for (int f = 0; f < 4; ++f)
                        if (n - f >= 0) S.ey2[j & 1][f][n - f] = e;
                }
                if (blk == NBLK - 1) load_ring_row(S, cam, proj, H, W, y + R + 1, x0, tid);   // next row's incoming image row
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&S.ops_bar)) : "memory");
            }
            TOCK(t_build);
            if (j > 0) {
                TICK();
                // ---- epilogue of job j-1 ----
                const int je = j - 1, y = h0 + je / NBLK, blk = je % NBLK;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (je & 1) * 256 + col_first;
                const float e2 = S.ex2[(je / NBLK) & 1][L];
                const float4 *ey = reinterpret_cast<const float4 *>(&S.ey2[je & 1][phi][NQ * cq]);
                const int s_first = mp - 131 + blk * PW + col_first;       // = 0 (mod 4)
                const int p_first = P_top4 - (blk * PW + col_first);
                const bool pcheck = P_top4 - (blk * PW + NT + 3) < 0;       // block-uniform: some projector columns are off-image
                if (blk == 0) { bv = -INFINITY; bs = 0; }
                int bi = -1;
                float *srow = &S.stage[L][NQ * cq];
                // 44 columns in four TMEM loads; the next load is in flight while a chunk is processed
                uint32_t ra[16], rb[16];
                auto cells = [&](const uint32_t *r, int i0, int n) {
#pragma unroll
                    for (int g = 0; g < n / 4; ++g) {
                        const bool ok = (unsigned)(s_first + i0 + 4 * g) < (unsigned)D;
                        const float4 e4 = ey[i0 / 4 + g];
                        const float ee[4] = {e4.x, e4.y, e4.z, e4.w};
                        float v[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int i = i0 + 4 * g + k;
                            float val = (__uint_as_float(r[4 * g + k]) + kEps) * rsqrtf(fmaf(e2, ee[k], kEps));
                            if (pcheck) {
                                const bool on = p_first - i >= 0;
                                val = on ? val : -2.f;
                                if (ok && on && val >= bv) { bv = val; bi = i; }
                            } else if (ok && val >= bv) { bv = val; bi = i; }
                            v[k] = val;
                        }
                        *reinterpret_cast<float4 *>(srow + i0 + 4 * g) = make_float4(v[0], v[1], v[2], v[3]);
                    }
                };
                tmem_ld16(ra, taddr);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tmem_ld16(rb, taddr + 16);
                if (pcheck) {
                    cells(ra, 0, 16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    tmem_ld8(ra, taddr + 32);
                    cells(rb, 16, 16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    tmem_ld4(rb, taddr + 40);
                    cells(ra, 32, 8);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    cells(rb, 40, 4);
                } else {
                    cells(ra, 0, 16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    tmem_ld8(ra, taddr + 32);
                    cells(rb, 16, 16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    tmem_ld4(rb, taddr + 40);
                    cells(ra, 32, 8);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    cells(rb, 40, 4);
                }
                if (bi >= 0) bs = s_first + bi;
                if (blk == NBLK - 1) { S.wv[cq][L] = bv; S.ws[cq][L] = bs; }
                TOCK(t_math);
                TICK();
                asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory");
                TOCK(t_bar);
                TICK();
                // ---- write-out: 128 rows x 44 float4 slots, consecutive threads -> consecutive 16 B of one camera column ----
                float *obase = out + ((size_t)b * H + y) * W * D;
                if (tid < 11 * (PW / 4)) {
                    const int g4 = tid % (PW / 4), row0 = tid / (PW / 4);
                    const int s_off = blk * PW - 131 + 4 * g4;
#pragma unroll
                    for (int k = 0; k < 12; ++k) {
                        const int Lr = row0 + 11 * k;
                        const int qL = Lr >> 5, mL = 4 * (Lr & 31) + qL, xr = x0 + mL;
                        const int s = mL + s_off + ((3 - qL) & 3);
                        if (Lr < MT && xr < W && (unsigned)s < (unsigned)D)
                            __stcs(reinterpret_cast<float4 *>(obase + (size_t)xr * D + s), *reinterpret_cast<const float4 *>(&S.stage[Lr][4 * g4]));
                    }
                }
                if (blk == NBLK - 1 && cq == 0 && x < W) {   // row complete: merge the four column quarters of camera column x
                    float mv = S.wv[0][L];
                    int ms = S.ws[0][L];
#pragma unroll
                    for (int c = 1; c < 4; ++c) {
                        const float ov = S.wv[c][L];
                        const int os = S.ws[c][L];
                        if (ov > mv || (ov == mv && os > ms)) { mv = ov; ms = os; }
                    }
                    best[((size_t)b * H + y) * W + x] = mv;
                    index[((size_t)b * H + y) * W + x] = ms;
                }
                TOCK(t_out);
                TICK();
                asm volatile("bar.sync 2, %0;" ::"n"(NWORK) : "memory");
                TOCK(t_bar);
            }
        }
        if (dbg && blockIdx.x == 3 && blockIdx.y == 3 && blockIdx.z == 0 && (tid == 0 || tid == 300)) {
            long long *d = dbg + (tid ? 8 : 0);
            d[0] = t_wait; d[1] = t_build; d[2] = t_math; d[3] = t_bar; d[4] = t_out; d[6] = njobs;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
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

int main(int argc, char **argv) {
    const int H = 375, W = 1242, B = 8;
    const int RB = argc > 1 ? atoi(argv[1]) : 25;
    const size_t npx = (size_t)H * W, ncell = npx * D;
    std::vector<float> hc(npx * B), hp(npx * B);
    float *cam, *proj, *out, *best;
    int *index;
    double *ref;
    const int check_rows = 24;
    CK(cudaMalloc(&cam, npx * 4 * B)); CK(cudaMalloc(&proj, npx * 4 * B)); CK(cudaMalloc(&out, ncell * 4 * B));
    CK(cudaMalloc(&best, npx * 4 * B)); CK(cudaMalloc(&index, npx * 4 * B));
    CK(cudaMalloc(&ref, (size_t)check_rows * W * D * 8));
    const size_t smem = sizeof(Smem);
    printf("shared memory per CTA: %zu bytes, %d projector blocks per row\n", smem, NBLK);
    std::vector<float> hb(npx);
    std::vector<int> hi(npx);
    CK(cudaFuncSetAttribute(tc_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> ho((size_t)check_rows * W * D);
    std::vector<double> hr((size_t)check_rows * W * D);
    dim3 grid((W + MT - 1) / MT, (H + RB - 1) / RB, 1);
    for (int family = 0; family < 3; ++family) {
        srand(1234 + family);
        for (size_t i = 0; i < npx * B; ++i) {
            const float u = rand() / (float)RAND_MAX, v = rand() / (float)RAND_MAX;
            const int x = i % W, y = (i / W) % H;
            if (family == 0) { hc[i] = u; hp[i] = v; }
            else if (family == 1) { hc[i] = 0.2f + 0.6f * x / W + 0.05f * (u - 0.5f); hp[i] = 0.2f + 0.6f * x / W + 0.05f * (v - 0.5f); }
            else { hc[i] = 0.5f + 0.4f * sinf(x * 0.01f) * cosf(y * 0.02f) + 0.01f * (u - 0.5f); hp[i] = 0.5f + 0.4f * sinf((x + 40) * 0.01f) * cosf(y * 0.02f) + 0.01f * (v - 0.5f); }
        }
        CK(cudaMemcpy(cam, hc.data(), npx * 4 * B, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(proj, hp.data(), npx * 4 * B, cudaMemcpyHostToDevice));
        CK(cudaMemset(out, 0xff, ncell * 4));
        tc_forward<<<grid, NTHREADS, smem>>>(cam, proj, out, best, index, H, W, RB, nullptr);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        double worst = 0; size_t bad = 0, wta_bad = 0, unwritten = 0;
        CK(cudaMemcpy(hb.data(), best, npx * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hi.data(), index, npx * 4, cudaMemcpyDeviceToHost));
        for (int y0 : {0, 180, H - check_rows}) {
            ref_forward<<<(unsigned)(((size_t)check_rows * W * D + 255) / 256), 256>>>(cam, proj, ref, H, W, y0, check_rows);
            CK(cudaMemcpy(hr.data(), ref, hr.size() * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(ho.data(), out + (size_t)y0 * W * D, hr.size() * 4, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < hr.size(); ++i) {
                const double e = fabs((double)ho[i] - hr[i]);
                if (ho[i] != ho[i]) ++unwritten;
                if (!(e <= 1e-5)) ++bad;
                if (e > worst) worst = e;
            }
            for (size_t px = 0; px < (size_t)check_rows * W; ++px) {   // WTA = arg-max of the kernel's own volume, ties -> largest s
                float m = -INFINITY; int ms = 0;
                for (int s = D - 1; s >= 0; --s) if (ho[px * D + s] > m && ho[px * D + s] != -2.f) { m = ho[px * D + s]; ms = s; }
                const size_t gp = (size_t)y0 * W + px;
                if (hb[gp] != m || hi[gp] != ms) ++wta_bad;
            }
        }
        printf("family %d: max |cost - fp64| = %.3e over %zu cells, %zu above 1e-5 (%zu never written), %zu WTA mismatches\n", family,
               worst, 3 * hr.size(), bad, unwritten, wta_bad);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    long long *dbg; CK(cudaMalloc(&dbg, 16 * 8)); CK(cudaMemset(dbg, 0, 16 * 8));
    grid.z = B;
    tc_forward<<<grid, NTHREADS, smem>>>(cam, proj, out, best, index, H, W, RB, dbg);
    cudaEventRecord(e0);
    for (int it = 0; it < 5; ++it) tc_forward<<<grid, NTHREADS, smem>>>(cam, proj, out, best, index, H, W, RB, dbg);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long hd[16]; CK(cudaMemcpy(hd, dbg, sizeof(hd), cudaMemcpyDeviceToHost));
    for (int t = 0; t < 2; ++t)
        printf("thread %3d clocks per job: wait %lld  build %lld  math %lld  barriers %lld  write-out %lld  (jobs %lld)\n", t ? 300 : 0,
               hd[8 * t] / hd[8 * t + 6], hd[8 * t + 1] / hd[8 * t + 6], hd[8 * t + 2] / hd[8 * t + 6], hd[8 * t + 3] / hd[8 * t + 6],
               hd[8 * t + 4] / hd[8 * t + 6], hd[8 * t + 6]);
    printf("prototype v3, %d pairs, RB=%d, grid %d CTAs: %.3f ms = %.1f Gcell/s (%.1f%% of the 4 B/cell HBM roofline at 6550 GB/s)\n", B, RB,
           grid.x * grid.y * grid.z, ms / 5, B * ncell / (ms / 5) * 1e-6, B * ncell * 4.0 / (ms / 5) * 1e-6 / 6550 * 100);
    return 0;
}
