// Shared pieces of the sliding-window kernels (sliding_prep.cu, sliding_forward.cu, sliding_backward.cu):
// the tiling, the workspace layout the prep kernels fill and the main kernels stream into shared memory with
// cp.async + mbarrier, and the PTX wrappers for those (sm_100a).
//
// Tiling of the cost volume (reference: custma/src/stereo_matching_kernel.cu:17-72 computes one cell per thread):
//   thread        4 camera columns x 4 disparities, marching down the rows of a row band; the k x k window sum of
//                 the product image cam[y,x] * proj[y,x-s] is kept as a register ring of k-1 row sums (BoxRing)
//   unit          16 lanes = 64 consecutive disparities of the same 4 columns (lane l owns s0 = s_base+64*su+4*l)
//   CTA           NU units x WG column groups = a tile of WTC = 4*WG columns, SC = 64*NU disparities, RB rows
// With the last axis of the volume on the lanes every store is a run of >= 256 contiguous bytes.
// s is the disparity (projector column d = w - s).  In banded mode s runs over [0, D); in full (reference-shaped)
// mode the same kernel covers s in [w - (W-1), w], i.e. every projector column, and writes cell [h, w, d = w - s].
#pragma once
#include "common.cuh"

namespace custma {

constexpr int kFallbackRows = 4;    // rows of a flagged tile per fallback work item (few tiles -> cut finely)
constexpr int kSlidingStages = 8;   // row slots of the shared-memory ring
constexpr int kLookahead = kSlidingStages - 2;  // a slot is refilled two steps after every warp released it

struct SlidingConfig {
    int K, NU, WG;
    __host__ __device__ int r() const { return K / 2; }
    __host__ __device__ int WTC() const { return 4 * WG; }
    __host__ __device__ int SC() const { return 64 * NU; }
    __host__ __device__ int CL() const { return (K + 3 + 3) & ~3; }  // camera values a thread reads per row (padded to 4)
    __host__ __device__ int PL() const { return (K + 6 + 3) & ~3; }  // projector values a thread reads per row
    __host__ __device__ int seg_cam() const { return 4 * (WG - 1) + CL(); }
    __host__ __device__ int seg_proj() const { return 4 * (WG + 16 * NU - 2) + PL(); }
    __host__ __device__ int seg_cs() const { return 4 * WG; }                      // camera statistics per row
    __host__ __device__ int seg_ps() const { return 4 * (WG + 16 * NU - 2) + 8; }  // projector statistics per row
    __host__ __device__ int consumer_threads() const { return 16 * NU * WG; }
};

// Everything the kernels need to find their data; computed on the host by make_sliding_layout().
struct SlidingLayout {
    int32_t K, r, NU, WG, WTC, SC;
    int32_t RB, RBH, NB;          // rows per band, rows of a band copy (RB + K - 1), number of bands
    int32_t n_wtiles, n_chunks;   // column tiles, disparity chunks per column tile
    int32_t banded, smin_full;    // full mode: chunk c of tile wt starts at rounddown4(w_base - (W-1)) + c*SC (chunk_s_base)
    int32_t cam_lc, cam_pitch;    // pivoted camera copy: [B][NB][RBH][cam_pitch], tile-major: tile wt holds image
                                  // columns wt*WTC - r + j at index wt*seg_cam + j (j < seg_cam), shifted by the
                                  // tile's own pivot; cam_lc is unused (kept 0)
    int32_t proj_lp, proj_pitch;  // pivoted projector copy: [B][NB][RBH][proj_pitch], column X at index X + proj_lp
    int32_t cs_pitch;             // camera statistics  [B][NB*RB][cs_pitch]   (A = window sum / k^2, ex2)
    int32_t ps_ld, ps_pitch;      // projector statistics [B][NB*RB][ps_pitch], column d at index d + ps_ld (Sp, ey2)
    int32_t seg_cam, seg_proj, seg_cs, seg_ps, slot_floats;
    // byte offsets into the workspace
    size_t off_minmax, off_camP, off_projP, off_A, off_ex2, off_Sp, off_ey2, off_wta, off_extra, total;
    // conditioning summaries per (pair, band, block of 16 columns) and the per-tile verdicts derived from them
    int32_t nblk_cs, nblk_ps;
    size_t off_rho_c, off_rho_p, off_bandany, off_flags, off_tileany, zero_end;
    size_t off_campiv;              // camera pivot per (pair, band, column tile)
    size_t off_fb_pm, off_fb_ey2;   // per-pixel projector window mean / second moment for the fallback kernels
    size_t off_fb_count, off_fb_list;   // work list of the fallback kernels: (tile, group of kFallbackRows rows) items
    int32_t fb_groups;              // row groups per band
};

// flags[tile] != 0: the tile is ill-conditioned for the O(1) window sums and is computed by the direct kernels
__host__ __device__ inline int64_t tile_index(const SlidingLayout &L, int b, int nb, int wt, int ch) {
    return (((int64_t)b * L.NB + nb) * L.n_wtiles + wt) * L.n_chunks + ch;
}

__host__ __device__ inline int round_down4(int v) { return v & ~3; }  // two's complement: also right for negatives

// first disparity of chunk `c` of the column tile starting at w_base
__host__ __device__ inline int chunk_s_base(const SlidingLayout &L, int W, int w_base, int c) {
    return (L.banded ? 0 : round_down4(w_base - (W - 1))) + c * L.SC;
}

// ---- mbarrier / bulk-copy wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
// producer side: the slot it waits for is usually many steps away, so poll slowly and leave the issue slots alone
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(128);
}
// global -> shared bulk copy (TMA, 1-D); size and both addresses are multiples of 16 bytes
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only) and its completion hook on an mbarrier
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
// the barrier receives one arrival from this thread once all of its earlier cp.async have landed (.noinc: the
// arrival is part of the barrier's expected count)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float rsqrt_fast(float x) {  // MUFU.RSQ, 2 ulp; x >= 1e-8 here, never denormal
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Packed fp32 pairs.  sm_100 issues two IEEE fp32 operations per FFMA2 / FADD2 / FMUL2 instruction; every half rounds
// exactly like the scalar instruction, so a packed formulation is bit-identical to the scalar one it replaces.  An
// operand is an aligned 64-bit register pair in either order (the instruction can swap its halves) or one 32-bit
// register broadcast to both halves: pk2(x, x) and pk2(hi, lo) of two neighbouring registers cost nothing, a pair that
// straddles an alignment boundary costs a move.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo2(f32x2 v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo;
}
__device__ __forceinline__ float hi2(f32x2 v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return hi;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// order-preserving map float -> uint32 (so that integer max is float max); never returns 0 for a finite float
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// compile-time geometry of one kernel configuration (mirrors SlidingConfig)
template <int K, int NU, int WG>
struct SlideGeom {
    static constexpr int r = K / 2, WTC = 4 * WG, SC = 64 * NU, CL = (K + 6) & ~3, PL = (K + 9) & ~3;
    static constexpr int SEG_CAM = 4 * (WG - 1) + CL, SEG_PROJ = 4 * (WG + 16 * NU - 2) + PL, SEG_CS = 4 * WG,
                         SEG_PS = 4 * (WG + 16 * NU - 2) + 8;
    static constexpr int OFF_PROJ = SEG_CAM, OFF_A = OFF_PROJ + SEG_PROJ, OFF_EX2 = OFF_A + SEG_CS,
                         OFF_SP = OFF_EX2 + SEG_CS, OFF_EY2 = OFF_SP + SEG_PS, SLOT = OFF_EY2 + SEG_PS;
    static constexpr int NCONS = 16 * NU * WG, NCW = (NCONS + 31) / 32, NS = kSlidingStages;
    static constexpr int PERIOD = K - 2;  // length of the pair-sum ring == unroll factor of the row loop
    // NCONS need not be a multiple of 32: the backward (no warp shuffles) runs 15 units = 240 threads for 192
    // disparities; the forward's WTA reduction shuffles within half-warps and uses whole warps only
    static_assert(K % 2 == 1 && K >= 3, "pair-sum ring is written for odd k");
};

// Horizontal window sums of one product row for a thread's 4 columns x 4 disparities, then the vertical k-row sum
// through the pair-sum ring:  box_t = c_t + P_{t-1} + P_{t-3} + ...,  P_t = c_t + c_{t-1}   ((k+1)/2 adds per cell).
// `seed` is added once to every horizontal sum (eps / k, so that the box holds raw + eps).
template <int K>
struct BoxRing {
    float cprev[4][4];
    float P[K - 2][4][4];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cprev[i][j] = 0.f;
#pragma unroll
                for (int m = 0; m < K - 2; ++m) P[m][i][j] = 0.f;
            }
    }
    // Q = t mod (K-2); a compile-time constant once the caller's row loop is unrolled by K-2.
    // DIR 1 / 2: the other three columns reuse the neighbouring column's sum (add the entering product, subtract the
    // leaving one), sliding left-to-right (1) or right-to-left (2); DIR 0: every column gets its own k-term chain.
    // Sliding carries the rounding of the leaving products into the next column.  That is harmless inside the image
    // but not next to the zero padding, where a product with the padding value -pivot can be orders of magnitude
    // larger than everything in the neighbouring windows - so the chain must always slide TOWARDS the padding: tiles
    // at the left border (camera columns < 0, projector columns w - s < 0) slide right-to-left, the others
    // left-to-right, and tiles that could see padding on both sides do not slide.
    template <int DIR>
    __device__ __forceinline__ void step(const int Q, const float *c, const float *pj, float seed, float (&bx)[4][4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const int i = DIR == 2 ? 3 - n : n;
                if (n == 0 || DIR == 0) {
                    s = fmaf(c[i], pj[i + 3 - j], seed);
#pragma unroll
                    for (int x = 1; x < K; ++x) s = fmaf(c[i + x], pj[i + x + 3 - j], s);
                } else if (DIR == 1) {
                    s = fmaf(c[i + K - 1], pj[i + K - 1 + 3 - j], s);
                    s = fmaf(-c[i - 1], pj[i - 1 + 3 - j], s);
                } else {
                    s = fmaf(c[i], pj[i + 3 - j], s);
                    s = fmaf(-c[i + K], pj[i + K + 3 - j], s);
                }
                float box = s;
#pragma unroll
                for (int m = 1; m <= K - 2; m += 2) box += P[(Q - m + 2 * (K - 2)) % (K - 2)][i][j];
                P[Q][i][j] = s + cprev[i][j];
                cprev[i][j] = s;
                bx[i][j] = box;
            }
        }
    }
};

// BoxRing on packed pairs: pair jp of column i holds the disparities (2 jp, 2 jp + 1); the same operations in the same
// order per cell, two cells per instruction
template <int K>
struct BoxRing2 {
    f32x2 cprev[4][2];
    f32x2 P[K - 2][4][2];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
                cprev[i][jp] = 0ull;
#pragma unroll
                for (int m = 0; m < K - 2; ++m) P[m][i][jp] = 0ull;
            }
    }
    template <int DIR>
    __device__ __forceinline__ void step(const int Q, const float *c, const float *pj, float seed, f32x2 (&bx)[4][2]) {
#pragma unroll
        for (int jp = 0; jp < 2; ++jp) {
            const int j = 2 * jp;
            // the projector values of disparities (j, j + 1) at camera column n: pj[n + 3 - j], pj[n + 2 - j]
            auto pp = [&](int n) { return pk2(pj[n + 3 - j], pj[n + 2 - j]); };
            f32x2 s = 0ull;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const int i = DIR == 2 ? 3 - n : n;
                if (n == 0 || DIR == 0) {
                    s = fma2(pk2(c[i], c[i]), pp(i), pk2(seed, seed));
#pragma unroll
                    for (int x = 1; x < K; ++x) s = fma2(pk2(c[i + x], c[i + x]), pp(i + x), s);
                } else if (DIR == 1) {
                    s = fma2(pk2(c[i + K - 1], c[i + K - 1]), pp(i + K - 1), s);
                    s = fma2(pk2(-c[i - 1], -c[i - 1]), pp(i - 1), s);
                } else {
                    s = fma2(pk2(c[i], c[i]), pp(i), s);
                    s = fma2(pk2(-c[i + K], -c[i + K]), pp(i + K), s);
                }
                f32x2 box = s;
#pragma unroll
                for (int m = 1; m <= K - 2; m += 2) box = add2(box, P[(Q - m + 2 * (K - 2)) % (K - 2)][i][jp]);
                P[Q][i][jp] = add2(s, cprev[i][jp]);
                cprev[i][jp] = s;
                bx[i][jp] = box;
            }
        }
    }
};

// Streams one row step (pivoted image rows + per-pixel statistics of the row that becomes complete) into its ring
// slot.  Every consumer thread copies at most NCH 16-byte chunks per step with cp.async and hooks its completion on
// the slot's "full" barrier (count = all consumer threads); a slot is refilled once its "empty" barrier (count = one
// arrival per warp) has completed, kLookahead steps ahead of its use.
// (A single thread issuing six cp.async.bulk copies per step was measured slower: the copies are 64-850 bytes and
// the consumers ended up waiting on them; per-thread 16-byte copies have no per-copy overhead to amortise.)
template <int K, int NU, int WG>
struct RowLoader {
    using G = SlideGeom<K, NU, WG>;
    static constexpr int CHUNKS = G::SLOT / 4, NCH = (CHUNKS + G::NCONS - 1) / G::NCONS;
    const float *src[NCH];
    int stride[NCH], dst[NCH], first[NCH];
    __device__ __forceinline__ void init(const SlidingLayout &L, const char *ws, int b, int nb, int h0, int w_base,
                                         int s_base) {
        const int xlo = w_base - G::r - s_base - G::SC + 1, dlo = w_base - s_base - G::SC + 1;
        const int64_t band = (int64_t)b * L.NB + nb;
        const int64_t srow = ((int64_t)b * L.NB * L.RB + h0) - (K - 1);  // statistics row of step t is srow + t
#pragma unroll
        for (int n = 0; n < NCH; ++n) {
            const int off = 4 * ((int)threadIdx.x + n * G::NCONS);
            dst[n] = off < G::SLOT ? off : -1;
            first[n] = off < G::OFF_A ? 0 : K - 1;
            if (off < G::OFF_PROJ) {
                stride[n] = L.cam_pitch;
                src[n] = (const float *)(ws + L.off_camP) + band * L.RBH * L.cam_pitch + (w_base / G::WTC) * G::SEG_CAM + off;
            } else if (off < G::OFF_A) {
                stride[n] = L.proj_pitch;
                src[n] = (const float *)(ws + L.off_projP) + band * L.RBH * L.proj_pitch + (xlo + L.proj_lp) + (off - G::OFF_PROJ);
            } else if (off < G::OFF_EX2) {
                stride[n] = L.cs_pitch;
                src[n] = (const float *)(ws + L.off_A) + srow * L.cs_pitch + w_base + (off - G::OFF_A);
            } else if (off < G::OFF_SP) {
                stride[n] = L.cs_pitch;
                src[n] = (const float *)(ws + L.off_ex2) + srow * L.cs_pitch + w_base + (off - G::OFF_EX2);
            } else if (off < G::OFF_EY2) {
                stride[n] = L.ps_pitch;
                src[n] = (const float *)(ws + L.off_Sp) + srow * L.ps_pitch + (dlo + L.ps_ld) + (off - G::OFF_SP);
            } else {
                stride[n] = L.ps_pitch;
                src[n] = (const float *)(ws + L.off_ey2) + srow * L.ps_pitch + (dlo + L.ps_ld) + (off - G::OFF_EY2);
            }
        }
    }
    // must be called for t = 0, 1, 2, ... in order (the source pointers advance by one row per call)
    __device__ __forceinline__ void issue(int t, float *smem, uint64_t *full_bar, uint64_t *empty_bar) {
        const int slot = t & (G::NS - 1);
        if (t >= G::NS) mbar_wait(&empty_bar[slot], ((t / G::NS) - 1) & 1);
        float *S = smem + slot * G::SLOT;
#pragma unroll
        for (int n = 0; n < NCH; ++n) {
            if (dst[n] >= 0 && t >= first[n]) cp_async16(S + dst[n], src[n]);
            src[n] += stride[n];
        }
        cp_async_mbar_arrive(&full_bar[slot]);
    }
};

// host side (sliding_prep.cu)
bool sliding_pick_config(const Problem &p, bool backward, SlidingConfig *cfg);
void make_sliding_layout(const Problem &p, const SlidingConfig &cfg, bool backward, SlidingLayout *L);
int validate_sliding_layout(const Problem &p, bool backward);
int launch_sliding_prep(const Problem &p, const SlidingLayout &L, const float *cam, const float *proj, char *ws,
                        cudaStream_t stream);
// sliding_fallback.cu: the flagged tiles, cell by cell in the reference's arithmetic
int launch_fallback_forward(const Problem &p, const SlidingLayout &L, const float *cam, const float *proj,
                            const char *ws, float *cost, unsigned long long *keys, const HeadOut &head,
                            uint32_t tc_threshold, cudaStream_t stream);
int launch_fallback_patch_grad(const Problem &p, const SlidingLayout &L, const float *grad, const HeadGrad &hg,
                               const float *cam, const float *proj, const char *ws, float *patch_grad,
                               uint32_t tc_threshold, cudaStream_t stream);

}  // namespace custma
