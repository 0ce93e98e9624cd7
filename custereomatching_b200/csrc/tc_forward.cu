// Tensor-core forward: the ZNCC cost volume of ill-conditioned inputs as a dense contraction on tcgen05 (sm_100a).
//
// What it restates: forward_cost_volume_kernel (reference custma/src/stereo_matching_kernel.cu:17-72) in the
// reference's own, well-conditioned form: both patches are centred by their window means first (:39-57), then
//     exy[x, p] = sum_tap cc[x][tap] * pc[p][tap]                                                  (:60-69)
// For one image row this IS a dense contraction, CC[128 x k*k] * PC[columns x k*k]^T (examples/verify.py:116 writes it
// as a bmm).  The sliding-window kernels (sliding_forward.cu) avoid the k*k factor with raw-moment box filters, which
// is 25x less arithmetic but cancels on low-texture images; their verdict (sliding_prep.cu) hands such inputs to this
// kernel instead of the slow per-cell fallback.  Here the k*k factor goes to the tensor cores: three
// tcgen05.mma kind::tf32 passes (lo*hi + hi*lo + hi*hi, "3xTF32": operands split into two tf32 halves, fp32
// accumulation in TMEM) reproduce the fp32 result to ~1e-6 (measured, tools/tc_proto3.cu), with no pivots and no
// conditioning limits.
//
//   CTA            persistent, one per SM: 16 worker warps + 1 MMA-issuing warp, tiles = 128 camera columns x RB rows
//   job            (row y, projector block blk):  D[128 x 192] (TMEM, fp32) = CC[128 x K] * PC[192 x K]^T
//   iteration j    workers: wait MMA j-1 | build the operand tiles of job j from a 6-row image ring in shared memory
//                  | arrive on ops_bar | epilogue of job j-1 from TMEM buffer (j-1)&1 | staged 128-bit stores
//                  MMA warp: wait ops_bar | 3 x K/8 tcgen05.mma into TMEM buffer j&1 | tcgen05.commit -> mma_bar[j&1]
//   TMEM lane L    camera column x0 + 4*(L%32) + L/32: the 32 lanes of a warp share x mod 4, so one warp-uniform
//                  column shift makes every thread's accumulator registers start at a disparity s = 0 (mod 4); 128-bit
//                  shared and global accesses then line up although s = x - p shears the tile
//   WTA            an epilogue thread owns one camera column: a thread-local running maximum (ties -> larger s =
//                  lower projector column, as the atomicMax keys of the sliding kernels), merged over the four column
//                  quarters through shared memory; written as the same packed keys, decoded by wta_decode_kernel
//
// Operand tiles are K-major without swizzle: element (row, k) lives at chunk (k/4) * rows*16 B + row*16 B + (k%4)*4 B, so
// a core matrix (8 rows x 16 B) is 128 contiguous bytes, SBO (next 8 rows) = 128 B, LBO (next 16 B of K) = rows*16 B.
#include <algorithm>

#include "sliding_common.cuh"

namespace custma {
namespace tc {

constexpr int MT = 128;                    // camera columns per tile = MMA M = TMEM lanes
constexpr int PW = 176, NT = 192;          // projector columns processed / computed per job (NT - PW >= 3: the shift)
constexpr int NQ = PW / 4;                 // columns per worker warp (four warps share a TMEM lane quadrant)
constexpr int SLD = PW + 4;                // stage pitch (floats)
constexpr int RING = 6, CAMW = 136, PRW = 728;
constexpr int kMaxBlocks = 4, kMaxD = kMaxBlocks * PW - 131;   // 573 -> D <= 572
constexpr int NWORK = 512, NTHREADS = NWORK + 32;

template <int KW>
struct Geom {
    static constexpr int R = KW / 2, NTAP = KW * KW, TAPS = (NTAP + 7) / 8 * 8, CH = TAPS / 4, CHW = (NTAP + 3) / 4;
};

template <int KW>
struct Smem {
    using G = Geom<KW>;
    float Ahi[G::CH][MT][4], Alo[G::CH][MT][4];
    float Bhi[G::CH][NT][4], Blo[G::CH][NT][4];
    float stage[MT][SLD];
    float camring[RING][CAMW], prjring[RING][PRW];
    float ex2[2][MT];
    float ey2[2][4][NT + 4];   // [job parity][shift f][column c] = second moment of computed column c + f
    float wv[4][MT];
    int ws[4][MT];
    unsigned long long ops_bar, mma_bar[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Split for 3xTF32.  The tensor core reads only the upper 19 bits of an fp32 operand (it truncates), so both halves are
// rounded to nearest here, with two integer instructions each: hi = rn_tf32(v), lo = rn_tf32(v - hi).  Left to the
// hardware, the truncation of lo alone would cost 2^-22 per operand, several times the fp32 rounding of the reference.
__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float tf32_hi(float v) { return tf32_rn(v); }
__device__ __forceinline__ float tf32_lo(float v, float hi) { return tf32_rn(v - hi); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // start address, leading / stride byte offsets (all >> 4), descriptor version 1, no swizzle
    return (uint64_t)((addr & 0x3ffffu) >> 4) | (uint64_t)(lbo_bytes >> 4) << 16 | (uint64_t)(sbo_bytes >> 4) << 32 | 1ull << 46;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory"); }

// image row yy of the tile's camera / projector column ranges into its ring slot; zeros outside the image
// (the reference's query_ij, kernel.cu:6-12)
template <int KW>
__device__ __forceinline__ void load_ring_row(Smem<KW> &S, const float *cam, const float *proj, int H, int W, int yy,
                                              int cam_x0, int prj_x0, int prj_w, int tid) {
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

// One patch row out of the ring: k*k taps centred by their mean (reference kernel.cu:39-57), split into tf32 hi / lo
// halves (the K padding of the operand tiles stays zero).  Returns the centred second moment (ex2 / ey2, :63-68).
template <int KW, int ROWS>
__device__ __forceinline__ float build_patch(const float *ring, int pitch, int y, int col0, float (*hi)[ROWS][4],
                                             float (*lo)[ROWS][4], int row) {
    using G = Geom<KW>;
    float v[4 * G::CHW];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < KW; ++i) {
        const float *src = ring + ((y + i + RING - G::R) % RING) * pitch + col0;
#pragma unroll
        for (int j = 0; j < KW; ++j) { v[i * KW + j] = src[j]; sum += v[i * KW + j]; }
    }
    const float mean = sum * (1.f / (float)G::NTAP);
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < 4 * G::CHW; ++t) {
        v[t] = t < G::NTAP ? v[t] - mean : 0.f;
        q = fmaf(v[t], v[t], q);
    }
#pragma unroll
    for (int c = 0; c < G::CHW; ++c) {
        float4 h, l;
        h.x = tf32_hi(v[4 * c]); h.y = tf32_hi(v[4 * c + 1]); h.z = tf32_hi(v[4 * c + 2]); h.w = tf32_hi(v[4 * c + 3]);
        l.x = tf32_lo(v[4 * c], h.x); l.y = tf32_lo(v[4 * c + 1], h.y); l.z = tf32_lo(v[4 * c + 2], h.z); l.w = tf32_lo(v[4 * c + 3], h.w);
        *reinterpret_cast<float4 *>(hi[c][row]) = h;
        *reinterpret_cast<float4 *>(lo[c][row]) = l;
    }
    return q;
}

template <int KW, bool COST, bool WTA>
__global__ void __launch_bounds__(NTHREADS, 1)
    tc_forward_kernel(const Problem p, const int RB, const int n_bands, const uint32_t *__restrict__ fb_count,
                      const uint32_t threshold, const float *__restrict__ cam_all, const float *__restrict__ proj_all,
                      float *__restrict__ cost, unsigned long long *__restrict__ keys) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    using G = Geom<KW>;
    // adaptive use: only when the sliding path's verdict flagged more work items than the threshold
    if (fb_count != nullptr && *fb_count <= threshold) return;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem<KW> &S = *reinterpret_cast<Smem<KW> *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.H, W = p.W, D = p.D;
    const int nblk = (D + 131 + PW - 1) / PW;                 // projector blocks per row
    const int p_span = (nblk - 1) * PW + NT - 1;              // computed projector columns: P_top4 - p_span .. P_top4
    const int prj_w = p_span + 1 + 2 * G::R;
    const int n_xt = (W + MT - 1) / MT;
    const int64_t n_tiles = (int64_t)p.B * n_bands * n_xt;

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
    // zero the operand tiles once: the K padding is never written again
    for (int i = tid; i < (int)((sizeof(S.Ahi) + sizeof(S.Alo) + sizeof(S.Bhi) + sizeof(S.Blo)) / 16); i += NTHREADS)
        reinterpret_cast<float4 *>(smem_raw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = S.tmem_base;

    uint32_t J = 0;   // jobs issued so far by this CTA (barrier phases and TMEM buffers alternate with it)
    if (warp == NWORK / 32) {
        // ================= MMA warp =================
        // instruction descriptor: D = f32, A = B = tf32, both K-major, N = NT, M = MT
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
        const uint32_t a_hi = smem_u32(S.Ahi), a_lo = smem_u32(S.Alo), b_hi = smem_u32(S.Bhi), b_lo = smem_u32(S.Blo);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int nb = (int)((tile / n_xt) % n_bands);
            const int njobs = min(RB, H - nb * RB) * nblk;
            for (int j = 0; j < njobs; ++j, ++J) {
                bar_wait_backoff(smem_u32(&S.ops_bar), J & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    uint32_t acc = 0;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {   // small terms first: lo*hi, hi*lo, hi*hi
                        const uint32_t a = pass == 0 ? a_lo : a_hi, bb = pass == 1 ? b_lo : b_hi;
#pragma unroll
                        for (int kk = 0; kk < G::TAPS / 8; ++kk) {
                            mma_tf32(tmem_base + (J & 1) * 256, make_desc(a + kk * 2 * (MT * 16), MT * 16, 128),
                                     make_desc(bb + kk * 2 * (NT * 16), NT * 16, 128), idesc, acc);
                            acc = 1;
                        }
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&S.mma_bar[J & 1])) : "memory");
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
            const int x0 = xt * MT, h0 = nb * RB, x = x0 + mp;
            const int rows = min(RB, H - h0), njobs = rows * nblk;
            const float *cam = cam_all + (int64_t)b * H * W, *proj = proj_all + (int64_t)b * H * W;
            const int P_top4 = x0 + MT - 1 + 4;                // projector column of computed column 0 of block 0
            const int cam_x0 = x0 - G::R, prj_x0 = P_top4 - p_span - G::R;
            worker_sync();                                      // the previous tile is completely done with the ring
            for (int yy = h0 - G::R; yy <= h0 + G::R; ++yy) load_ring_row<KW>(S, cam, proj, H, W, yy, cam_x0, prj_x0, prj_w, tid);
            worker_sync();
            float bv = -INFINITY;
            int bs = 0;
            for (int j = 0; j <= njobs; ++j) {
                if (j > 0) bar_wait(smem_u32(&S.mma_bar[(J - 1) & 1]), ((J - 1) >> 1) & 1);   // MMA of the previous job is done
                if (j < njobs) {
                    const int y = h0 + j / nblk, blk = j % nblk;
                    const int nA = blk == 0 ? MT : 0;
                    if (tid < nA) {
                        const int col0 = 4 * (tid & 31) + (tid >> 5);      // TMEM lane tid <-> camera column x0 + col0
                        S.ex2[(j / nblk) & 1][tid] = build_patch<KW, MT>(&S.camring[0][0], CAMW, y, col0, S.Ahi, S.Alo, tid);
                    } else if (tid < nA + NT) {
                        const int n = tid - nA;                             // computed column n <-> projector column P_top4 - (blk*PW + n)
                        const float e = build_patch<KW, NT>(&S.prjring[0][0], PRW, y, p_span - blk * PW - n, S.Bhi, S.Blo, n);
#pragma unroll
                        for (int f = 0; f < 4; ++f)
                            if (n - f >= 0) S.ey2[J & 1][f][n - f] = e;
                    }
                    if (blk == nblk - 1)   // the image row the next row's windows need (its slot holds row y - 3)
                        load_ring_row<KW>(S, cam, proj, H, W, y + G::R + 1, cam_x0, prj_x0, prj_w, tid);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> tensor core
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&S.ops_bar)) : "memory");
                }
                if (j > 0) {
                    // ---- epilogue of the previous job: reference kernel.cu:71 on the accumulators ----
                    const int je = j - 1, y = h0 + je / nblk, blk = je % nblk;
                    const uint32_t Je = J - 1;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (Je & 1) * 256 + col_first;
                    const float e2 = S.ex2[(je / nblk) & 1][L];
                    const float4 *ey = reinterpret_cast<const float4 *>(&S.ey2[Je & 1][phi][NQ * cq]);
                    const int s_first = mp - (MT + 3) + blk * PW + col_first;   // = 0 (mod 4)
                    const int p_first = P_top4 - (blk * PW + col_first);
                    const bool pcheck = P_top4 - (blk * PW + NT + 3) < 0;        // block-uniform: some projector columns are off-image
                    if (blk == 0) { bv = -INFINITY; bs = 0; }
                    int bi = -1;
                    float *srow = &S.stage[L][NQ * cq];
                    uint32_t ra[16], rb[16];
                    auto cells = [&](const uint32_t *r, int i0, int n, bool check) {
#pragma unroll
                        for (int g = 0; g < n / 4; ++g) {
                            const bool ok = (unsigned)(s_first + i0 + 4 * g) < (unsigned)D;   // D % 4 == 0: whole groups
                            const float4 e4 = ey[i0 / 4 + g];
                            const float ee[4] = {e4.x, e4.y, e4.z, e4.w};
                            float v[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int i = i0 + 4 * g + k;
                                float val = (__uint_as_float(r[4 * g + k]) + kEps) * rsqrt_fast(fmaf(e2, ee[k], kEps));
                                if (check) {
                                    const bool on = p_first - i >= 0;
                                    val = on ? val : kInvalid;
                                    if (WTA && ok && on && val >= bv) { bv = val; bi = i; }
                                } else if (WTA && ok && val >= bv) { bv = val; bi = i; }
                                v[k] = val;
                            }
                            if (COST) *reinterpret_cast<float4 *>(srow + i0 + 4 * g) = make_float4(v[0], v[1], v[2], v[3]);
                        }
                    };
                    // 44 columns in four TMEM loads; the next load is in flight while a chunk is processed
                    tmem_ld16(ra, taddr);
                    tmem_wait_ld();
                    tmem_ld16(rb, taddr + 16);
                    if (pcheck) {
                        cells(ra, 0, 16, true);  tmem_wait_ld(); tmem_ld8(ra, taddr + 32);
                        cells(rb, 16, 16, true); tmem_wait_ld(); tmem_ld4(rb, taddr + 40);
                        cells(ra, 32, 8, true);  tmem_wait_ld();
                        cells(rb, 40, 4, true);
                    } else {
                        cells(ra, 0, 16, false);  tmem_wait_ld(); tmem_ld8(ra, taddr + 32);
                        cells(rb, 16, 16, false); tmem_wait_ld(); tmem_ld4(rb, taddr + 40);
                        cells(ra, 32, 8, false);  tmem_wait_ld();
                        cells(rb, 40, 4, false);
                    }
                    if (bi >= 0) bs = s_first + bi;
                    if (WTA && blk == nblk - 1) { S.wv[cq][L] = bv; S.ws[cq][L] = bs; }
                    worker_sync();
                    if (COST && tid < 11 * NQ) {
                        // write-out: thread = one 16-byte column of the stage, rows stepping by 11; consecutive threads
                        // write consecutive 16 bytes of one camera column's disparity run
                        const int g4 = tid % NQ, row0 = tid / NQ;
                        const int s_off = blk * PW - (MT + 3) + 4 * g4;
                        float *obase = cost + ((int64_t)b * H + y) * W * D;
#pragma unroll
                        for (int k = 0; k < 12; ++k) {
                            const int Lr = row0 + 11 * k;
                            const int qL = Lr >> 5, mL = 4 * (Lr & 31) + qL, xr = x0 + mL;
                            const int s = mL + s_off + ((3 - qL) & 3);
                            if (Lr < MT && xr < W && (unsigned)s < (unsigned)D)
                                __stcs(reinterpret_cast<float4 *>(obase + (int64_t)xr * D + s),
                                       *reinterpret_cast<const float4 *>(&S.stage[Lr][4 * g4]));
                        }
                    }
                    if (WTA && blk == nblk - 1 && cq == 0 && x < W) {   // row complete: merge the four column quarters
                        float mv = S.wv[0][L];
                        int ms = S.ws[0][L];
#pragma unroll
                        for (int c = 1; c < 4; ++c) {
                            const float ov = S.wv[c][L];
                            const int os = S.ws[c][L];
                            if (ov > mv || (ov == mv && os > ms)) { mv = ov; ms = os; }
                        }
                        keys[((int64_t)b * H + y) * W + x] = ((unsigned long long)float_to_ordered(mv) << 32) | (uint32_t)(ms + W);
                    }
                }
                worker_sync();   // stage, ring row and WTA hand-over are free again
                if (j < njobs) ++J;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// rows per tile: few enough tiles-per-CTA rounding loss, little per-tile start-up (ring fill: 2R+1 rows)
static int pick_rows(const Problem &p, int n_ctas) {
    const int64_t per_row = (int64_t)p.B * ((p.W + MT - 1) / MT);
    int best_rb = 8;
    double best_cost = 1e30;
    for (int rb = 4; rb <= 32; ++rb) {
        const int64_t tiles = per_row * ((p.H + rb - 1) / rb);
        const double c = (double)((tiles + n_ctas - 1) / n_ctas) * (rb + 1.5);
        if (c < best_cost) { best_cost = c; best_rb = rb; }
    }
    return std::min(best_rb, std::max(1, (int)p.H));
}

template <int KW, bool COST, bool WTA>
static int launch_one(const Problem &p, const float *cam, const float *proj, float *cost, unsigned long long *keys,
                      const uint32_t *fb_count, uint32_t threshold, cudaStream_t stream) {
    int dev = 0, n_sm = 148;
    CUSTMA_CUDA_CHECK(cudaGetDevice(&dev));
    CUSTMA_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int RB = pick_rows(p, n_sm), n_bands = (p.H + RB - 1) / RB;
    const int64_t n_tiles = (int64_t)p.B * n_bands * ((p.W + MT - 1) / MT);
    auto kern = tc_forward_kernel<KW, COST, WTA>;
    const size_t smem = sizeof(Smem<KW>);
    CUSTMA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUSTMA_CUDA_CHECK(launch_chained(kern, dim3((unsigned)std::min<int64_t>(n_tiles, n_sm)), dim3(NTHREADS), smem, stream, p, RB, n_bands,
                                     fb_count, threshold, cam, proj, cost, keys));
    CUSTMA_LAUNCH_CHECK("tc_forward_kernel");
    return CUSTMA_OK;
}

template <int KW>
static int launch_k(const Problem &p, const float *cam, const float *proj, float *cost, unsigned long long *keys,
                    const uint32_t *fb_count, uint32_t threshold, cudaStream_t stream) {
    if (cost && keys) return launch_one<KW, true, true>(p, cam, proj, cost, keys, fb_count, threshold, stream);
    if (cost) return launch_one<KW, true, false>(p, cam, proj, cost, keys, fb_count, threshold, stream);
    return launch_one<KW, false, true>(p, cam, proj, cost, keys, fb_count, threshold, stream);
}

}  // namespace tc

bool tc_forward_supported(const Problem &p) {
    return p.banded && (p.k == 3 || p.k == 5) && p.D >= 4 && (p.D & 3) == 0 && p.D <= tc::kMaxD;
}

int launch_tc_forward(const Problem &p, const float *cam, const float *proj, float *cost, unsigned long long *keys,
                      const uint32_t *fb_count, uint32_t threshold, cudaStream_t stream) {
    if (!tc_forward_supported(p)) return set_error(CUSTMA_ERR_UNSUPPORTED, "no tensor-core forward for k=%d D=%d", p.k, p.D);
    return p.k == 3 ? tc::launch_k<3>(p, cam, proj, cost, keys, fb_count, threshold, stream)
                    : tc::launch_k<5>(p, cam, proj, cost, keys, fb_count, threshold, stream);
}

}  // namespace custma
