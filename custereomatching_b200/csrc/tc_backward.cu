// Tensor-core backward: the camera gradient of ill-conditioned inputs on tcgen05 (sm_100a), the counterpart of
// tc_forward.cu.
//
// What it restates: get_patches_grad_kernel + patches_grad_to_image_kernel (reference
// custma/src/stereo_matching_kernel.cu:75-152, :155-179).  Per cell the reference forms  A = g / den,
// B = g * ey2 * (exy + eps) / den^3  (:135,:145-148) and adds the patch gradient A * pc[p][tap] - B * cc[x][tap] to the
// per-pixel patch buffer with atomicAdd (:149), which a second kernel scatters to the image (:178).  Summed over the
// projector columns of one camera pixel that is again a dense contraction:
//     MMA1   D1[128 x 128]  = CC * PC^T                                 exy, as in the forward (3xTF32, fp32 in TMEM)
//     epi1   a = g * rsqrt(ex2 * ey2 + eps)  -> two tf32 halves written back into TMEM (tcgen05.st),
//            csum += a * (exy - ex2 * ey2) / den^2                       per camera column, fp32 (see "row")
//     MMA2   G1[128 x 32]  += A_hi * PC_hi + A_hi * PC_lo + A_lo * PC_hi          A operand from TMEM, PC MN-major
//     row    patch_grad[x] = G1[x] - (<G1[x], cc[x]> / ex2) * cc[x] + (eps * csum / ex2) * cc[x]
//            (sum_p (a pc - b cc) with the component along cc taken analytically instead of as a difference of two
//            large sums), scattered into the tile's gradient image in shared memory, one window column per phase:
//            no two threads touch one element, no atomics
// Each tile (128 camera columns x up to 16 rows) writes its gradient image with a window-radius halo to a scratch
// tile; tc_backward_finalize_kernel adds the overlapping halos in a fixed order: deterministic, like the sliding path.
//
// Pipeline of a job j (workers = 16 warps, one more warp issues the MMAs): barrier | cp.async of the upstream-gradient
// tile into the stage | build CC (first job of a row) and PC | arrive -> MMA1(j) | wait for the copies, barrier, wait
// for MMA1(j) (which, the tensor core running in order, also means MMA2(j-1) is done) | epi1 | arrive -> MMA2(j), which
// runs under the build of job j+1 (D1, A_hi, A_lo and G1 have their own TMEM columns; the MN-major PC copy alternates
// between two buffers).
//
// PC lives in shared memory twice: K-major without swizzle for MMA1 (as in tc_forward.cu) and as rows of 32 taps with
// 32-byte units XOR (row % 4) for MMA2 - the only MN-major layout the tensor core accepts for 32-bit operands
// (UMMA layout type 1, "128B swizzle with 32-byte base"; tools/tc_mma_probe.cu: every other MN-major layout made the
// MMA a no-op).  Thread / TMEM-lane / column mapping and the image ring are those of tc_forward.cu, with 112 projector
// columns per job so that everything fits in 226 KB of shared memory.
#include <algorithm>

#include "sliding_common.cuh"

namespace custma {
namespace tcb {

constexpr int MT = 128;                       // camera columns per tile = MMA M = TMEM lanes
constexpr int PW = 112, NT = 128;             // projector columns processed / computed per job
constexpr int NQ = PW / 4, SLD = PW + 4;      // columns per worker warp, stage pitch
constexpr int TAPS2 = 32;                     // MMA2 N: window taps padded to one 128-byte row
constexpr int RING = 6, CAMW = 136, PRW = 696;
constexpr int RBMAX = 16, GW = MT + 4;        // most rows per tile, width of the tile's gradient image
constexpr int kMaxBlocks = 6, kMaxD = kMaxBlocks * PW - 131;
constexpr int NWORK = 512, NTHREADS = NWORK + 32;
constexpr int COL_D1 = 0, COL_AH = NT, COL_AL = 2 * NT, COL_G1 = 3 * NT;   // TMEM columns: D1, A_hi, A_lo, G1 (416 of 512)

template <int KW>
struct Geom {
    static constexpr int R = KW / 2, NTAP = KW * KW, TAPS = (NTAP + 7) / 8 * 8, CH = TAPS / 4, CHW = (NTAP + 3) / 4;
};

template <int KW>
struct Smem {
    using G = Geom<KW>;
    float Ahi[G::CH][MT][4], Alo[G::CH][MT][4];
    float Bhi[G::CH][NT][4], Blo[G::CH][NT][4];   // PC, K-major (MMA1)
    float B2hi[2][NT][TAPS2], B2lo[2][NT][TAPS2];  // PC, MN-major with the 32-byte-base swizzle (MMA2); by job parity,
                                                   // so that MMA2 of job j runs while job j+1 is being built
    float stage[MT][SLD];                          // cost_volume_grad tile of the job
    float camring[RING][CAMW], prjring[RING][PRW];
    float gring[RBMAX + 4][GW];                       // the tile's camera-gradient image incl. halo
    float ex2[MT];
    float ey2[4][NT + 4];
    float bs[4][MT];
    unsigned long long ops1_bar, ops2_bar, mma1_bar, mma2_bar;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Split for 3xTF32.  The tensor core reads only the upper 19 bits of an fp32 operand (it truncates), so both halves are
// rounded to nearest here, with two integer instructions each: hi = rn_tf32(v), lo = rn_tf32(v - hi).  Left to the
// hardware, the truncation of lo alone would cost 2^-22 per operand, several times the fp32 rounding of the reference.
__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float tf32_hi(float v) { return tf32_rn(v); }
__device__ __forceinline__ float tf32_lo(float v, float hi) { return tf32_rn(v - hi); }
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
// the MMA-issuing warp waits for a whole operand build: poll slowly and leave its scheduler's issue slots to the four
// worker warps that share it (ncu r01: 15 % of all executed instructions were these spin loops)
__device__ __forceinline__ void bar_wait_backoff(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(64);
    }
}
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

template <int KW>
__device__ __forceinline__ void load_ring_row(Smem<KW> &S, const float *cam, const float *proj, int H, int W, int yy, int cam_x0,
                                              int prj_x0, int prj_w, int tid) {
    constexpr int CW = MT + 2 * Geom<KW>::R;
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

template <int KW, int ROWS>
__device__ __forceinline__ float build_patch(const float *ring, int pitch, int y, int col0, float (*hi)[ROWS][4], float (*lo)[ROWS][4], int row,
                                             float (*hi2)[TAPS2] = nullptr, float (*lo2)[TAPS2] = nullptr) {
    using G = Geom<KW>;
    constexpr int R = G::R, NTAP = G::NTAP, CHW = G::CHW;
    float v[4 * CHW];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < KW; ++i) {
        const float *src = ring + ((y + i + RING - R) % RING) * pitch + col0;
#pragma unroll
        for (int j = 0; j < KW; ++j) { v[i * KW + j] = src[j]; sum += v[i * KW + j]; }
    }
    const float mean = sum * (1.f / (float)NTAP);
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
        l.x = tf32_lo(v[4 * c], h.x); l.y = tf32_lo(v[4 * c + 1], h.y); l.z = tf32_lo(v[4 * c + 2], h.z); l.w = tf32_lo(v[4 * c + 3], h.w);
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

template <int KW>
__global__ void __launch_bounds__(NTHREADS, 1)
    tc_backward_kernel(const Problem p, const int RB, const int n_bands, const uint32_t *__restrict__ fb_count, const uint32_t threshold,
                       const float *__restrict__ cam_all, const float *__restrict__ proj_all, const float *__restrict__ grad,
                       float *__restrict__ scratch) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    using G = Geom<KW>;
    constexpr int R = G::R, NTAP = G::NTAP, TAPS = G::TAPS;
    // adaptive use: only when the sliding path's verdict flagged more work items than the threshold
    if (fb_count != nullptr && *fb_count <= threshold) return;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem<KW> &S = *reinterpret_cast<Smem<KW> *>(smem_raw);
    const int B = p.B, H = p.H, W = p.W, D = p.D;
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
    if (warp < NWORK / 32) {
        // A_hi / A_lo columns outside a warp quadrant's window [phi, phi + PW) are never written by the epilogue: zero once
        const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
        const uint32_t zeros[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if ((warp >> 2) == 0) { tmem_st4(lane_addr + COL_AH, zeros); tmem_st4(lane_addr + COL_AL, zeros); }
        if ((warp >> 2) == 3) { tmem_st16(lane_addr + COL_AH + PW, zeros); tmem_st16(lane_addr + COL_AL + PW, zeros); }
        tmem_wait_st();
        fence_before();
    }
    if (warp == NWORK / 32) {
        // ================= MMA warp =================
        const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
        // MMA2: N = 32 taps, B operand MN-major (bit 16)
        const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(TAPS2 >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
        const uint32_t a_hi = smem_u32(S.Ahi), a_lo = smem_u32(S.Alo), b_hi = smem_u32(S.Bhi), b_lo = smem_u32(S.Blo);
        const uint32_t b2_hi[2] = {smem_u32(S.B2hi[0]), smem_u32(S.B2hi[1])}, b2_lo[2] = {smem_u32(S.B2lo[0]), smem_u32(S.B2lo[1])};
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int nb = (int)((tile / n_xt) % n_bands);
            const int rows = min(RB, H - nb * RB);
            for (int r = 0; r < rows; ++r)
                for (int blk = 0; blk < nblk; ++blk, ++J) {
                    bar_wait_backoff(smem_u32(&S.ops1_bar), J & 1);
                    fence_after();
                    if (lane == 0) {
                        uint32_t acc = 0;
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {
                            const uint32_t a = pass == 0 ? a_lo : a_hi, bb = pass == 1 ? b_lo : b_hi;
#pragma unroll
                            for (int kk = 0; kk < TAPS / 8; ++kk) {
                                mma_ss(tmem_base + COL_D1, make_desc(a + kk * 2 * (MT * 16), MT * 16, 128),
                                       make_desc(bb + kk * 2 * (NT * 16), NT * 16, 128), idesc1, acc);
                                acc = 1;
                            }
                        }
                        commit(smem_u32(&S.mma1_bar));
                    }
                    __syncwarp();
                    bar_wait_backoff(smem_u32(&S.ops2_bar), J & 1);
                    fence_after();
                    if (lane == 0) {
                        uint32_t acc = blk == 0 ? 0u : 1u;
#pragma unroll 1
                        for (int pass = 0; pass < 3; ++pass) {   // A_lo*PC_hi, A_hi*PC_lo, A_hi*PC_hi
                            const uint32_t ta = tmem_base + (pass == 0 ? COL_AL : COL_AH), bb = pass == 1 ? b2_lo[J & 1] : b2_hi[J & 1];
#pragma unroll 4
                            for (int kk = 0; kk < NT / 8; ++kk) {
                                mma_ts(tmem_base + COL_G1, ta + kk * 8, make_desc(bb + kk * 1024, 0, 512, 1), idesc2, acc);
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
            for (int yy = h0 - R; yy <= h0 + R; ++yy) load_ring_row<KW>(S, cam, proj, H, W, yy, cam_x0, prj_x0, prj_w, tid);
            for (int i = tid; i < (RBMAX + 4) * GW; i += NWORK) (&S.gring[0][0])[i] = 0.f;
            worker_sync();
            for (int r = 0; r < rows; ++r) {
                const int y = h0 + r;
                float bsum = 0.f;
                for (int blk = 0; blk < nblk; ++blk, ++J) {
                    worker_sync();   // every worker is done with the previous job's stage / ey2; its GEMM 1 has been waited for
                    // gradient tile of this job -> stage (same slots as the forward's write-out; zeros where no cell exists),
                    // as 16-byte cp.async so that it is in flight during the operand build
                    if (tid < 18 * NQ) {
                        const int g4 = tid % NQ, row0 = tid / NQ;
                        const int s_off = blk * PW - (MT + 3) + 4 * g4;
                        const bool grow = y >= p.g0 && y < p.g1;        // rows outside [g0, g1) carry no gradient
                        const float *gbase = grad + ((int64_t)b * p.grows() + (grow ? y - p.g0 : 0)) * W * D;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int Lr = row0 + 18 * k;
                            const int qL = Lr >> 5, mL = 4 * (Lr & 31) + qL, xr = x0 + mL;
                            const int s = mL + s_off + ((3 - qL) & 3);
                            if (Lr < MT) {
                                const bool live = grow && xr < W && (unsigned)s < (unsigned)D;
                                const float *src = live ? gbase + (int64_t)xr * D + s : gbase;
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(&S.stage[Lr][4 * g4])), "l"(src),
                                             "r"(live ? 16 : 0) : "memory");
                            }
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    const int nA = blk == 0 ? MT : 0;
                    if (tid < nA) {
                        const int col0 = 4 * (tid & 31) + (tid >> 5);
                        S.ex2[tid] = build_patch<KW, MT>(&S.camring[0][0], CAMW, y, col0, S.Ahi, S.Alo, tid);
                    } else if (tid < nA + NT) {
                        const int n = tid - nA;
                        const float e = build_patch<KW, NT>(&S.prjring[0][0], PRW, y, p_span - blk * PW - n, S.Bhi, S.Blo, n, S.B2hi[J & 1], S.B2lo[J & 1]);
#pragma unroll
                        for (int f = 0; f < 4; ++f)
                            if (n - f >= 0) S.ey2[f][n - f] = e;
                    }
                    if (blk == nblk - 1) load_ring_row<KW>(S, cam, proj, H, W, y + R + 1, cam_x0, prj_x0, prj_w, tid);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    fence_before();
                    bar_arrive(smem_u32(&S.ops1_bar));
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    worker_sync();                                   // stage and ey2 / ex2 are complete
                    bar_wait(smem_u32(&S.mma1_bar), J & 1);   // GEMM 1 of this job - and, the tensor core running in order, GEMM 2 of the
                    fence_after();                           // previous one: A_hi / A_lo may be overwritten
                    // ---- epilogue 1 ----
                    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16);
                    const float e2 = S.ex2[L];
                    const float4 *ey = reinterpret_cast<const float4 *>(&S.ey2[phi][NQ * cq]);
                    const float4 *gs = reinterpret_cast<const float4 *>(&S.stage[L][NQ * cq]);
                    const int p_first = P_top4 - (blk * PW + col_first);
                    const uint32_t t_d1 = lane_addr + COL_D1 + col_first, t_ah = lane_addr + COL_AH + col_first, t_al = lane_addr + COL_AL + col_first;
                    uint32_t d[16], dl[4], ah[16], al[16];
                    auto proc = [&](const uint32_t *dd, int i0, int n) {
#pragma unroll
                        for (int g = 0; g < n / 4; ++g) {
                            const float4 e4 = ey[i0 / 4 + g], g4v = gs[i0 / 4 + g];
                            const float ee[4] = {e4.x, e4.y, e4.z, e4.w}, gg[4] = {g4v.x, g4v.y, g4v.z, g4v.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int i = i0 + 4 * g + k;
                                const float rs = rsqrt_fast(fmaf(e2, ee[k], kEps));
                                float a = gg[k] * rs;                               // g / den; g is 0 where the cell does not exist
                                if (p_first - i < 0) a = 0.f;                       // off-image projector column: constant cost
                                // what is left of a*mu - b (mu = exy / ex2, see the row epilogue): a * (exy - ex2*ey2) / den^2
                                bsum = fmaf(a * rs * rs, fmaf(-e2, ee[k], __uint_as_float(dd[4 * g + k])), bsum);
                                const float h = tf32_hi(a);
                                ah[4 * g + k] = __float_as_uint(h);
                                al[4 * g + k] = __float_as_uint(tf32_lo(a, h));
                            }
                        }
                        if (n == 16) { tmem_st16(t_ah + i0, ah); tmem_st16(t_al + i0, al); }
                        else if (n == 8) { tmem_st8(t_ah + i0, ah); tmem_st8(t_al + i0, al); }
                        else { tmem_st4(t_ah + i0, ah); tmem_st4(t_al + i0, al); }
                    };
                    tmem_ld16(d, t_d1);
                    tmem_ld4(dl, t_d1 + 24);
                    tmem_wait_ld();
                    proc(d, 0, 16);
                    tmem_ld8(d, t_d1 + 16); tmem_wait_ld();
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
                    uint32_t g1[TAPS2];
                    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16);
                    tmem_ld16(g1, lane_addr + COL_G1);
                    tmem_ld16(g1 + 16, lane_addr + COL_G1 + 16);
                    tmem_wait_ld();
                    // patch_grad = sum_p (a_p * pc_p - b_p * cc)   (reference kernel.cu:145-158).  With pc_p = mu_p * cc + pc_p^perp,
                    // mu_p = <pc_p, cc> / <cc, cc> = exy_p / ex2, this is  sum_p a_p * pc_p^perp + (sum_p (a_p mu_p - b_p)) * cc,
                    // i.e. G1 with its component along cc projected out, plus a term that is O(eps):
                    //   a mu - b = a * eps * (exy - ex2*ey2) / (ex2 * den^2).
                    // Subtracting Bs * cc from G1 directly would cancel two large sums on highly correlated images (and
                    // expose the tensor core's truncating accumulation, which is mostly along G1 ~ cc); the projection
                    // removes that component whatever its rounding was.
                    const float csum = (S.bs[0][L] + S.bs[1][L]) + (S.bs[2][L] + S.bs[3][L]);
                    const float e2x = S.ex2[L], inv_e2 = e2x > 0.f ? 1.f / e2x : 0.f;
                    float cc[NTAP], dot = 0.f;
#pragma unroll
                    for (int t = 0; t < NTAP; ++t) {
                        cc[t] = S.Ahi[t / 4][L][t % 4] + S.Alo[t / 4][L][t % 4];
                        dot = fmaf(__uint_as_float(g1[t]), cc[t], dot);
                    }
                    const float coef = fmaf(kEps, csum, -dot) * inv_e2;
                    float pg[NTAP];
#pragma unroll
                    for (int t = 0; t < NTAP; ++t) pg[t] = fmaf(coef, cc[t], __uint_as_float(g1[t]));
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
            float *dst = scratch + tile * (int64_t)((RB + 4) * GW);
            for (int i = tid; i < (RB + 4) * GW; i += NWORK) dst[i] = (&S.gring[0][0])[i];
        }
    }
    fence_before();
    __syncthreads();
    fence_after();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// camera_grad[b, y, x] = sum of the (up to 3 x 2) tiles whose gradient image covers the pixel, in a fixed order
__global__ void __launch_bounds__(256)
    tc_backward_finalize_kernel(const Problem p, const int R, const int RB, const int n_bands, const uint32_t *__restrict__ fb_count,
                                const uint32_t threshold, const float *__restrict__ scratch, float *__restrict__ out) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    if (fb_count != nullptr && *fb_count <= threshold) return;
    const int H = p.H, W = p.W;
    const int n_xt = (W + MT - 1) / MT;
    // grid-stride: a fixed small grid, so the launch costs nothing when the kernel is not needed
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < p.pixels(); id += (int64_t)gridDim.x * blockDim.x) {
    const int x = id % W, y = (id / W) % H, b = id / ((int64_t)W * H);
    float acc = 0.f;
    for (int nb = max(0, y / RB - 1); nb <= min(n_bands - 1, y / RB + 1); ++nb) {
        const int row = y - (nb * RB - R);
        if (row < 0 || row >= RB + 2 * R) continue;
        for (int xt = max(0, x / MT - 1); xt <= min(n_xt - 1, x / MT + 1); ++xt) {
            const int col = x - (xt * MT - R);
            if (col < 0 || col >= MT + 2 * R) continue;
            acc += scratch[(((int64_t)b * n_bands + nb) * n_xt + xt) * ((RB + 4) * GW) + row * GW + col];
        }
    }
    out[id] = acc;
    }
}


// Rows per tile: the fewest rounds of tiles over the persistent CTAs (one per SM of the current device; the workspace
// query and the launch see the same device), little per-tile start-up (ring fill, halo rows).
static int pick_rows(const Problem &p) {
    const int64_t per_row = (int64_t)p.B * ((p.W + MT - 1) / MT);
    int best_rb = RBMAX;
    double best_cost = 1e30;
    for (int rb = RBMAX; rb >= 4; --rb) {
        const int64_t tiles = per_row * ((p.H + rb - 1) / rb);
        const int64_t sms = device_sm_count();
        const double c = (double)((tiles + sms - 1) / sms) * (rb + 2.0);
        if (c < best_cost) { best_cost = c; best_rb = rb; }
    }
    return best_rb;
}

template <int KW>
static int launch_k(const Problem &p, const float *grad, const float *cam, const float *proj, float *camera_grad,
                    float *scratch, const uint32_t *fb_count, uint32_t threshold, cudaStream_t stream) {
    int dev = 0, n_sm = 148;
    CUSTMA_CUDA_CHECK(cudaGetDevice(&dev));
    CUSTMA_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int RB = pick_rows(p), n_bands = (p.H + RB - 1) / RB;
    const int64_t n_tiles = (int64_t)p.B * n_bands * ((p.W + MT - 1) / MT);
    auto kern = tc_backward_kernel<KW>;
    const size_t smem = sizeof(Smem<KW>);
    CUSTMA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUSTMA_CUDA_CHECK(launch_chained(kern, dim3((unsigned)std::min<int64_t>(n_tiles, n_sm)), dim3(NTHREADS), smem, stream, p, RB, n_bands,
                                     fb_count, threshold, cam, proj, grad, scratch));
    CUSTMA_LAUNCH_CHECK("tc_backward_kernel");
    CUSTMA_CUDA_CHECK(launch_chained(tc_backward_finalize_kernel, dim3((unsigned)std::min<int64_t>((p.pixels() + 255) / 256, 8 * n_sm)), dim3(256), 0,
                                     stream, p, KW / 2, RB, n_bands, fb_count, threshold, (const float *)scratch, camera_grad));
    CUSTMA_LAUNCH_CHECK("tc_backward_finalize_kernel");
    return CUSTMA_OK;
}

}  // namespace tcb

bool tc_backward_supported(const Problem &p) {
    return p.banded && (p.k == 3 || p.k == 5) && p.D >= 4 && (p.D & 3) == 0 && p.D <= tcb::kMaxD;
}

size_t tc_backward_scratch_bytes(const Problem &p) {
    if (!tc_backward_supported(p)) return 0;
    const int rb = tcb::pick_rows(p);
    const int64_t n_tiles = (int64_t)p.B * ((p.H + rb - 1) / rb) * ((p.W + tcb::MT - 1) / tcb::MT);
    return align256((size_t)n_tiles * (rb + 4) * tcb::GW * sizeof(float));
}

int launch_tc_backward(const Problem &p, const float *grad, const float *cam, const float *proj, float *camera_grad,
                       float *scratch, const uint32_t *fb_count, uint32_t threshold, cudaStream_t stream) {
    if (!tc_backward_supported(p)) return set_error(CUSTMA_ERR_UNSUPPORTED, "no tensor-core backward for k=%d D=%d", p.k, p.D);
    return p.k == 3 ? tcb::launch_k<3>(p, grad, cam, proj, camera_grad, scratch, fb_count, threshold, stream)
                    : tcb::launch_k<5>(p, grad, cam, proj, camera_grad, scratch, fb_count, threshold, stream);
}

}  // namespace custma
