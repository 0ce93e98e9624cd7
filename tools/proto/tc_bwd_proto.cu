// Feasibility prototype (not part of the product): the ZNCC backward (camera gradient) on the tensor cores, the
// counterpart of custereomatching_b200/csrc/tc_forward.cu.  Per job (row y, projector block):
//   MMA1   D1[128 x 192]  = CC * PC^T                         (centred patches, 3xTF32, as in the forward)
//   epi1   a = g / den, b = g * ey2 * (exy + eps) / den^3     (reference stereo_matching_kernel.cu:135,145-148)
//          a -> tf32 hi / lo back into TMEM (hi over D1 in place), Bs += b per camera column
//   MMA2   G1[128 x 32] += A_hi * PC_hi + A_hi * PC_lo + A_lo * PC_hi      (A operand from TMEM, PC tile MN-major)
// and per row:  patch_grad[x][tap] = G1[x][tap] - Bs[x] * cc[x][tap], scattered into a per-tile gradient image in
// shared memory (5 column phases, no atomics), written to a scratch tile; a finalize kernel adds the overlapping
// halos of neighbouring tiles in a fixed order.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int KW = 5, R = 2, NTAP = 25, TAPS = 32, CH = 8, CHW = 7, MT = 128;
constexpr int PW = 112, NT = 128, NQ = 28, SLD = 116;
constexpr int RING = 6, CAMW = 136, PRW = 728, RBMAX = 16, GW = MT + 4;
constexpr int NWORK = 512, NTHREADS = NWORK + 32;
constexpr int COL_AH = 0, COL_AL = 128, COL_G1 = 256;
constexpr float kEps = 1e-8f;

struct Smem {
    float Ahi[CH][MT][4], Alo[CH][MT][4];
    float Bhi[CH][NT][4], Blo[CH][NT][4];     // PC, K-major (MMA1)
    float B2hi[NT][TAPS], B2lo[NT][TAPS];        // PC again, rows of 32 taps with 32-byte units XOR (row % 4): MN-major (MMA2)
    float stage[MT][SLD];
    float camring[RING][CAMW], prjring[RING][PRW];
    float gring[RBMAX + 4][GW];
    float ex2[MT];
    float ey2[4][NT + 4];
    float bs[4][MT];
    unsigned long long ops1_bar, ops2_bar, mma1_bar, mma2_bar;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// round to nearest tf32 (the tensor core itself truncates the low 13 mantissa bits): the residual v - hi then has at most
// 12 significant bits and loses at most one of them when it is used as the second tf32 operand
__device__ __forceinline__ float tf32_hi(float v) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); return __uint_as_float(r); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 0) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | (uint64_t)(lbo_bytes >> 4) << 16 | (uint64_t)(sbo_bytes >> 4) << 32 | 1ull << 46 | (uint64_t)layout << 61;
}
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
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
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                   "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void load_ring_row(Smem &S, const float *cam, const float *proj, int H, int W, int yy, int cam_x0,
                                              int prj_x0, int prj_w, int tid) {
    constexpr int CW = MT + 2 * R;
    const int slot = (yy + RING) % RING;
    const bool row_ok = yy >= 0 && yy < H;
    for (int i = tid; i < CW + prj_w; i += NWORK) {
        if (i < CW) {
            const int xc = cam_x0 + i;
            S.camring[slot][i] = (row_ok && xc >= 0 && xc < W) ? __ldg(cam + (int64_t)yy * W + xc) : 0.f;
        } else {
            const int pc = prj_x0 + (i - CW);
            S.prjring[slot][i - CW] = (row_ok && pc >= 0 && pc < W) ? __ldg(proj + (int64_t)yy * W + pc) : 0.f;
        }
    }
}

template <int ROWS>
__device__ __forceinline__ float build_patch(const float *ring, int pitch, int y, int col0, float (*hi)[ROWS][4], float (*lo)[ROWS][4], int row,
                                             float (*hi2)[TAPS] = nullptr, float (*lo2)[TAPS] = nullptr) {
    float v[4 * CHW];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < KW; ++i) {
        const float *src = ring + ((y + i + RING - R) % RING) * pitch + col0;
#pragma unroll
        for (int j = 0; j < KW; ++j) { v[i * KW + j] = src[j]; sum += v[i * KW + j]; }
    }
    const float mean = sum / (float)NTAP;
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < 4 * CHW; ++t) {
        v[t] = t < NTAP ? v[t] - mean : 0.f;
        q = fmaf(v[t], v[t], q);
    }
#pragma unroll
    for (int c = 0; c < CHW; ++c) {
        float4 h, l;
        h.x = tf32_hi(v[4 * c]); h.y = tf32_hi(v[4 * c + 1]); h.z = tf32_hi(v[4 * c + 2]); h.w = tf32_hi(v[4 * c + 3]);
        l.x = v[4 * c] - h.x; l.y = v[4 * c + 1] - h.y; l.z = v[4 * c + 2] - h.z; l.w = v[4 * c + 3] - h.w;
        *reinterpret_cast<float4 *>(hi[c][row]) = h;
        *reinterpret_cast<float4 *>(lo[c][row]) = l;
        if (hi2) {
            const int off = ((((c >> 1) ^ (row & 3)) << 1) + (c & 1)) * 4;
            *reinterpret_cast<float4 *>(&hi2[row][off]) = h;
            *reinterpret_cast<float4 *>(&lo2[row][off]) = l;
        }
    }
    return q;
}

__global__ void __launch_bounds__(NTHREADS, 1)
    tc_backward(const float *__restrict__ cam_all, const float *__restrict__ proj_all, const float *__restrict__ grad,
                float *__restrict__ scratch, int B, int H, int W, int D, int RB, int n_bands, int sbo2, int lbo2, float *dbg) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nblk = (D + 131 + PW - 1) / PW, p_span = (nblk - 1) * PW + NT - 1, prj_w = p_span + 1 + 2 * R;
    const int n_xt = (W + MT - 1) / MT;
    const int64_t n_tiles = (int64_t)B * n_bands * n_xt;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&S.ops1_bar)), "r"(NWORK));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&S.ops2_bar)), "r"(NWORK));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.mma1_bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.mma2_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = tid; i < (int)((sizeof(S.Ahi) + sizeof(S.Alo) + sizeof(S.Bhi) + sizeof(S.Blo) + sizeof(S.B2hi) + sizeof(S.B2lo)) / 16); i += NTHREADS)
        reinterpret_cast<float4 *>(smem_raw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = S.tmem_base;

    uint32_t J = 0;
    if (warp == NWORK / 32) {
        // ================= MMA warp =================
        const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
        // MMA2: N = 32 taps, B operand MN-major (bit 16)
        const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(TAPS >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
        const uint32_t a_hi = smem_u32(S.Ahi), a_lo = smem_u32(S.Alo), b_hi = smem_u32(S.Bhi), b_lo = smem_u32(S.Blo);
        const uint32_t b2_hi = smem_u32(S.B2hi), b2_lo = smem_u32(S.B2lo);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int nb = (int)((tile / n_xt) % n_bands);
            const int rows = min(RB, H - nb * RB);
            for (int r = 0; r < rows; ++r)
                for (int blk = 0; blk < nblk; ++blk, ++J) {
                    bar_wait(smem_u32(&S.ops1_bar), J & 1);
                    fence_after();
                    if (lane == 0) {
                        uint32_t acc = 0;
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {
                            const uint32_t a = pass == 0 ? a_lo : a_hi, bb = pass == 1 ? b_lo : b_hi;
#pragma unroll
                            for (int kk = 0; kk < TAPS / 8; ++kk) {
                                mma_ss(tmem_base + COL_AH, make_desc(a + kk * 2 * (MT * 16), MT * 16, 128),
                                       make_desc(bb + kk * 2 * (NT * 16), NT * 16, 128), idesc1, acc);
                                acc = 1;
                            }
                        }
                        commit(smem_u32(&S.mma1_bar));
                    }
                    __syncwarp();
                    bar_wait(smem_u32(&S.ops2_bar), J & 1);
                    fence_after();
                    if (lane == 0) {
                        uint32_t acc = blk == 0 ? 0u : 1u;
#pragma unroll 1
                        for (int pass = 0; pass < 3; ++pass) {   // A_lo*PC_hi, A_hi*PC_lo, A_hi*PC_hi
                            const uint32_t ta = tmem_base + (pass == 0 ? COL_AL : COL_AH), bb = pass == 1 ? b2_lo : b2_hi;
#pragma unroll 4
                            for (int kk = 0; kk < NT / 8; ++kk) {
                                mma_ts(tmem_base + COL_G1, ta + kk * 8, make_desc(bb + kk * 1024, lbo2, sbo2, 1), idesc2, acc);
                                acc = 1;
                            }
                        }
                        commit(smem_u32(&S.mma2_bar));
                    }
                    __syncwarp();
                }
        }
    } else {
        // ================= worker warps =================
        const int q = warp & 3, cq = warp >> 2, L = 32 * q + lane, mp = 4 * lane + q;
        const int phi = (3 - q) & 3, col_first = phi + NQ * cq;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int xt = (int)(tile % n_xt), nb = (int)((tile / n_xt) % n_bands), b = (int)(tile / ((int64_t)n_xt * n_bands));
            const int x0 = xt * MT, h0 = nb * RB;
            const int rows = min(RB, H - h0);
            const float *cam = cam_all + (int64_t)b * H * W, *proj = proj_all + (int64_t)b * H * W;
            const int P_top4 = x0 + MT - 1 + 4;
            const int cam_x0 = x0 - R, prj_x0 = P_top4 - p_span - R;
            worker_sync();
            for (int yy = h0 - R; yy <= h0 + R; ++yy) load_ring_row(S, cam, proj, H, W, yy, cam_x0, prj_x0, prj_w, tid);
            for (int i = tid; i < (RBMAX + 4) * GW; i += NWORK) (&S.gring[0][0])[i] = 0.f;
            worker_sync();
            for (int r = 0; r < rows; ++r) {
                const int y = h0 + r;
                float bsum = 0.f;
                for (int blk = 0; blk < nblk; ++blk, ++J) {
                    if (J > 0) bar_wait(smem_u32(&S.mma2_bar), (J - 1) & 1);   // MMA2 of the previous job no longer reads PC / A / TMEM
                    const int nA = blk == 0 ? MT : 0;
                    if (tid < nA) {
                        const int col0 = 4 * (tid & 31) + (tid >> 5);
                        S.ex2[tid] = build_patch<MT>(&S.camring[0][0], CAMW, y, col0, S.Ahi, S.Alo, tid);
                    } else if (tid < nA + NT) {
                        const int n = tid - nA;
                        const float e = build_patch<NT>(&S.prjring[0][0], PRW, y, p_span - blk * PW - n, S.Bhi, S.Blo, n, S.B2hi, S.B2lo);
#pragma unroll
                        for (int f = 0; f < 4; ++f)
                            if (n - f >= 0) S.ey2[f][n - f] = e;
                    }
                    if (blk == nblk - 1) load_ring_row(S, cam, proj, H, W, y + R + 1, cam_x0, prj_x0, prj_w, tid);
                    // gradient tile of this job -> stage (same slots as the forward's write-out), zeros where no cell exists
                    if (tid < 18 * NQ) {
                        const int g4 = tid % NQ, row0 = tid / NQ;
                        const int s_off = blk * PW - (MT + 3) + 4 * g4;
                        const float *gbase = grad + ((int64_t)b * H + y) * W * D;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int Lr = row0 + 18 * k;
                            const int qL = Lr >> 5, mL = 4 * (Lr & 31) + qL, xr = x0 + mL;
                            const int s = mL + s_off + ((3 - qL) & 3);
                            if (Lr < MT) {
                                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (xr < W && (unsigned)s < (unsigned)D) g = __ldcs(reinterpret_cast<const float4 *>(gbase + (int64_t)xr * D + s));
                                *reinterpret_cast<float4 *>(&S.stage[Lr][4 * g4]) = g;
                            }
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    fence_before();
                    bar_arrive(smem_u32(&S.ops1_bar));
                    worker_sync();                                   // stage and ey2 / ex2 are complete
                    bar_wait(smem_u32(&S.mma1_bar), J & 1);
                    fence_after();
                    // ---- epilogue 1 ----
                    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16);
                    const float e2 = S.ex2[L];
                    const float4 *ey = reinterpret_cast<const float4 *>(&S.ey2[phi][NQ * cq]);
                    const float4 *gs = reinterpret_cast<const float4 *>(&S.stage[L][NQ * cq]);
                    const int p_first = P_top4 - (blk * PW + col_first);
                    const uint32_t t_ah = lane_addr + COL_AH + col_first, t_al = lane_addr + COL_AL + col_first;
                    uint32_t d[16], dl[4], ah[16], al[16];
                    auto proc = [&](const uint32_t *dd, int i0, int n) {
#pragma unroll
                        for (int g = 0; g < n / 4; ++g) {
                            const float4 e4 = ey[i0 / 4 + g], g4v = gs[i0 / 4 + g];
                            const float ee[4] = {e4.x, e4.y, e4.z, e4.w}, gg[4] = {g4v.x, g4v.y, g4v.z, g4v.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int i = i0 + 4 * g + k;
                                const float rs = rsqrtf(fmaf(e2, ee[k], kEps));
                                float a = gg[k] * rs;                               // g / den; g is 0 where the cell does not exist
                                if (p_first - i < 0) a = 0.f;                       // off-image projector column: constant cost
                                const float cost = (__uint_as_float(dd[4 * g + k]) + kEps) * rs;
                                bsum = fmaf(a * cost, ee[k] * rs, bsum);       // g * ey2 * (exy + eps) / den^3
                                const float h = tf32_hi(a);
                                ah[4 * g + k] = __float_as_uint(h);
                                al[4 * g + k] = __float_as_uint(a - h);
                            }
                        }
                        if (n == 16) { tmem_st16(t_ah + i0, ah); tmem_st16(t_al + i0, al); }
                        else if (n == 8) { tmem_st8(t_ah + i0, ah); tmem_st8(t_al + i0, al); }
                        else { tmem_st4(t_ah + i0, ah); tmem_st4(t_al + i0, al); }
                    };
                    tmem_ld4(dl, t_ah + 24);
                    tmem_ld16(d, t_ah);
                    tmem_wait_ld();
                    // columns outside [phi, phi + 176) carry no cells of this job: zero them in both A halves (the D1
                    // columns this overwrites are already in this thread's registers)
                    {
                        const uint32_t zeros[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                        if (cq == 0) { tmem_st4(lane_addr + COL_AH, zeros); tmem_st4(lane_addr + COL_AL, zeros); }
                        if (cq == 3) { tmem_st16(lane_addr + COL_AH + PW, zeros); tmem_st16(lane_addr + COL_AL + PW, zeros); }
                        tmem_wait_st();
                    }
                    proc(d, 0, 16);
                    tmem_ld8(d, t_ah + 16); tmem_wait_ld();
                    proc(d, 16, 8);
                    proc(dl, 24, 4);
                    tmem_wait_st();
                    fence_before();
                    bar_arrive(smem_u32(&S.ops2_bar));
                }
                // ---- row complete: patch gradients out of G1 ----
                bar_wait(smem_u32(&S.mma2_bar), (J - 1) & 1);
                fence_after();
                S.bs[cq][L] = bsum;
                worker_sync();
                if (cq == 0) {
                    uint32_t g1[TAPS];
                    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16);
                    tmem_ld16(g1, lane_addr + COL_G1);
                    tmem_ld16(g1 + 16, lane_addr + COL_G1 + 16);
                    tmem_wait_ld();
                    const float Bs = (S.bs[0][L] + S.bs[1][L]) + (S.bs[2][L] + S.bs[3][L]);
                    if (dbg && tile == 1 && r == 5) {
                        for (int t = 0; t < TAPS; ++t) dbg[mp * 33 + t] = __uint_as_float(g1[t]);
                        dbg[mp * 33 + 32] = Bs;
                    }
                    float pg[NTAP];
#pragma unroll
                    for (int t = 0; t < NTAP; ++t) {
                        const float cc = S.Ahi[t / 4][L][t % 4] + S.Alo[t / 4][L][t % 4];
                        pg[t] = fmaf(-Bs, cc, __uint_as_float(g1[t]));
                    }
                    // scatter: tap (i, j) of camera column mp belongs to gradient pixel (r + i, mp + j); one column offset
                    // per phase, so no two threads touch the same element
#pragma unroll
                    for (int j = 0; j < KW; ++j) {
#pragma unroll
                        for (int i = 0; i < KW; ++i) S.gring[r + i][mp + j] += pg[i * KW + j];
                        asm volatile("bar.sync 2, 128;" ::: "memory");
                    }
                }
                fence_before();
                worker_sync();
            }
            // the tile's gradient image (rows h0-2 .. h0+RB+1, columns x0-2 .. x0+129) -> scratch
            float *dst = scratch + tile * (int64_t)((RBMAX + 4) * GW);
            for (int i = tid; i < (RBMAX + 4) * GW; i += NWORK) dst[i] = (&S.gring[0][0])[i];
        }
    }
    fence_before();
    __syncthreads();
    fence_after();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// camera_grad[b, y, x] = sum of the (up to 3 x 2) tiles whose gradient image covers the pixel, in a fixed order
__global__ void tc_backward_finalize(const float *__restrict__ scratch, float *__restrict__ out, int B, int H, int W, int RB, int n_bands) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (int64_t)B * H * W) return;
    const int x = id % W, y = (id / W) % H, b = id / ((int64_t)W * H);
    const int n_xt = (W + MT - 1) / MT;
    float acc = 0.f;
    for (int nb = max(0, y / RB - 1); nb <= min(n_bands - 1, y / RB + 1); ++nb) {
        const int row = y - (nb * RB - R);
        if (row < 0 || row >= RB + 4) continue;
        for (int xt = max(0, x / MT - 1); xt <= min(n_xt - 1, x / MT + 1); ++xt) {
            const int col = x - (xt * MT - R);
            if (col < 0 || col >= GW) continue;
            acc += scratch[(((int64_t)b * n_bands + nb) * n_xt + xt) * ((RBMAX + 4) * GW) + row * GW + col];
        }
    }
    out[id] = acc;
}

// fp64 reference: per cell the reference's patch gradient (kernel.cu:135-148), scattered with double atomics
__global__ void ref_backward(const float *cam, const float *proj, const float *grad, double *out, int H, int W, int D) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)H * W * D) return;
    const int s = id % D, x = (id / D) % W, y = id / ((long long)D * W);
    const int p = x - s;
    if (p < 0) return;
    double c[NTAP], q[NTAP], cm = 0, pm = 0;
    for (int i = 0; i < KW; ++i)
        for (int j = 0; j < KW; ++j) {
            const int yy = y + i - R, xc = x + j - R, xp = p + j - R;
            const bool oky = yy >= 0 && yy < H;
            c[i * KW + j] = (oky && xc >= 0 && xc < W) ? cam[(size_t)yy * W + xc] : 0.f;
            q[i * KW + j] = (oky && xp >= 0 && xp < W) ? proj[(size_t)yy * W + xp] : 0.f;
            cm += c[i * KW + j]; pm += q[i * KW + j];
        }
    cm /= NTAP; pm /= NTAP;
    double exy = 0, ex2 = 0, ey2 = 0;
    for (int t = 0; t < NTAP; ++t) { c[t] -= cm; q[t] -= pm; exy += c[t] * q[t]; ex2 += c[t] * c[t]; ey2 += q[t] * q[t]; }
    const double den = sqrt(ex2 * ey2 + 1e-8), g = grad[id];
    const double a = g / den, bb = g * ey2 * (exy + 1e-8) / (den * den * den);
    for (int i = 0; i < KW; ++i)
        for (int j = 0; j < KW; ++j) {
            const int yy = y + i - R, xc = x + j - R;
            if (yy >= 0 && yy < H && xc >= 0 && xc < W) atomicAdd(out + (size_t)yy * W + xc, a * q[i * KW + j] - bb * c[i * KW + j]);
        }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

static int run(int B, int H, int W, int D, bool check, int sbo2, int lbo2) {
    const size_t npx = (size_t)H * W, ncell = npx * D;
    std::vector<float> hc(npx * B), hp(npx * B), hg(ncell * B);
    srand(99);
    for (size_t i = 0; i < npx * B; ++i) {
        const float u = rand() / (float)RAND_MAX, v = rand() / (float)RAND_MAX;
        const int x = i % W, y = (i / W) % H;
        hc[i] = 0.5f + 0.4f * sinf(x * 0.01f) * cosf(y * 0.02f) + 0.02f * (u - 0.5f);
        hp[i] = 0.5f + 0.4f * sinf((x + 40) * 0.01f) * cosf(y * 0.02f) + 0.02f * (v - 0.5f);
    }
    for (size_t i = 0; i < ncell * B; ++i) hg[i] = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    float *cam, *proj, *grad, *out, *scratch;
    CK(cudaMalloc(&cam, npx * 4 * B)); CK(cudaMalloc(&proj, npx * 4 * B)); CK(cudaMalloc(&grad, ncell * 4 * B)); CK(cudaMalloc(&out, npx * 4 * B));
    CK(cudaMemcpy(cam, hc.data(), npx * 4 * B, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(proj, hp.data(), npx * 4 * B, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(grad, hg.data(), ncell * 4 * B, cudaMemcpyHostToDevice));
    const int RB = RBMAX, n_bands = (H + RB - 1) / RB, n_xt = (W + MT - 1) / MT;
    const int64_t n_tiles = (int64_t)B * n_bands * n_xt;
    CK(cudaMalloc(&scratch, n_tiles * (RBMAX + 4) * GW * 4));
    float *dbg; CK(cudaMalloc(&dbg, 128 * 33 * 4)); CK(cudaMemset(dbg, 0, 128 * 33 * 4));
    const size_t smem = sizeof(Smem);
    CK(cudaFuncSetAttribute(tc_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>(n_tiles, 148);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 2; ++it) {
        cudaEventRecord(e0);
        tc_backward<<<grid, NTHREADS, smem>>>(cam, proj, grad, scratch, B, H, W, D, RB, n_bands, sbo2, lbo2, dbg);
        tc_backward_finalize<<<(unsigned)((npx * B + 255) / 256), 256>>>(scratch, out, B, H, W, RB, n_bands);
        cudaEventRecord(e1);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("B=%d H=%d W=%d D=%d smem=%zu: %.3f ms = %.1f Gcell/s", B, H, W, D, smem, ms, B * ncell / ms * 1e-6);
    if (check) {
        double *ref; CK(cudaMalloc(&ref, npx * 8)); CK(cudaMemset(ref, 0, npx * 8));
        ref_backward<<<(unsigned)((ncell + 255) / 256), 256>>>(cam, proj, grad, ref, H, W, D);
        std::vector<double> hr(npx); std::vector<float> ho(npx);
        CK(cudaMemcpy(hr.data(), ref, npx * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(ho.data(), out, npx * 4, cudaMemcpyDeviceToHost));
        double worst = 0, scale = 0;
        for (size_t i = 0; i < npx; ++i) { scale = fmax(scale, fabs(hr[i])); worst = fmax(worst, fabs(ho[i] - hr[i])); }
        printf("   max |grad - fp64| = %.3e, max |grad| = %.3e, relative %.2e", worst, scale, worst / scale);
        cudaFree(ref);
        // G1 / Bs of tile 1 (camera columns 128..255), row 5, against a host evaluation in double
        std::vector<float> hd(128 * 33);
        CK(cudaMemcpy(hd.data(), dbg, hd.size() * 4, cudaMemcpyDeviceToHost));
        double wg = 0, sg = 0, wb = 0, sb = 0;
        const int y = 5;
        for (int m = 0; m < 128 && 128 + m < W; ++m) {
            const int x = 128 + m;
            double G[NTAP] = {0}, Bsum = 0;
            for (int s = 0; s < D; ++s) {
                const int p = x - s;
                if (p < 0) continue;
                double c[NTAP], q[NTAP], cm = 0, pm = 0;
                for (int i = 0; i < KW; ++i)
                    for (int j = 0; j < KW; ++j) {
                        const int yy = y + i - R, xc = x + j - R, xp = p + j - R;
                        const bool oky = yy >= 0 && yy < H;
                        c[i * KW + j] = (oky && xc >= 0 && xc < W) ? hc[(size_t)yy * W + xc] : 0.f;
                        q[i * KW + j] = (oky && xp >= 0 && xp < W) ? hp[(size_t)yy * W + xp] : 0.f;
                        cm += c[i * KW + j]; pm += q[i * KW + j];
                    }
                cm /= NTAP; pm /= NTAP;
                double exy = 0, ex2 = 0, ey2 = 0;
                for (int t = 0; t < NTAP; ++t) { c[t] -= cm; q[t] -= pm; exy += c[t] * q[t]; ex2 += c[t] * c[t]; ey2 += q[t] * q[t]; }
                const double den = sqrt(ex2 * ey2 + 1e-8), g = hg[((size_t)y * W + x) * D + s];
                for (int t = 0; t < NTAP; ++t) G[t] += g / den * q[t];
                Bsum += g * ey2 * (exy + 1e-8) / (den * den * den);
            }
            for (int t = 0; t < NTAP; ++t) { wg = fmax(wg, fabs(hd[m * 33 + t] - G[t])); sg = fmax(sg, fabs(G[t])); }
            wb = fmax(wb, fabs(hd[m * 33 + 32] - Bsum)); sb = fmax(sb, fabs(Bsum));
            if (m == 3) { printf("\n   x=131: G1 dev %g %g %g %g | host %g %g %g %g ; Bs dev %g host %g", hd[m*33], hd[m*33+1], hd[m*33+2], hd[m*33+24], G[0], G[1], G[2], G[24], hd[m*33+32], Bsum); }
        }
        printf("\n   G1: max err %.3e (scale %.3e)   Bs: max err %.3e (scale %.3e)", wg, sg, wb, sb);
    }
    printf("\n");
    cudaFree(dbg); cudaFree(cam); cudaFree(proj); cudaFree(grad); cudaFree(out); cudaFree(scratch);
    return 0;
}

int main(int argc, char **argv) {
    const int sbo2 = argc > 1 ? atoi(argv[1]) : 512, lbo2 = argc > 2 ? atoi(argv[2]) : 0;
    printf("MMA2 B descriptor: SBO %d LBO %d\n", sbo2, lbo2);
    if (run(1, 40, 300, 192, true, sbo2, lbo2)) return 1;
    if (run(1, 48, 600, 192, true, sbo2, lbo2)) return 1;
    if (argc > 3) return 0;
    if (run(1, 375, 1242, 192, false, sbo2, lbo2)) return 1;
    if (run(8, 375, 1242, 192, false, sbo2, lbo2)) return 1;
    return 0;
}
