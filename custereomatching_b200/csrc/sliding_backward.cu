// Sliding-window backward: gradient of the ZNCC volume with respect to the camera image, one pass over the upstream
// gradient (4 bytes per cell read once), no global atomics, deterministic.
//
// What it restates: get_patches_grad_kernel + patches_grad_to_image_kernel (reference
// custma/src/stereo_matching_kernel.cu:75-152, :155-179).  The reference recomputes the window moments of every cell
// (6k^2 loads) and then issues k^2 same-address global atomics per cell.  Here (checked against
// autograd by the oracle tests), with a = g/den and bc = g*ey2*(exy+eps)/den^3 per cell (reference :135,:145-148):
//
//   camera_grad[y,x] = T1[y,x] - sum_{(h,w) covering (y,x)} ( Am[h,w] + Bs[h,w] * (cam[y,x] - cmean[h,w]) )
//   T1[y,x] = sum_s proj'[y,x-s] * (k x k box sum over (h,w) of a[h,w,s])     <- this kernel, per tile
//   Am[h,w] = sum_s a[h,w,s] * pmean'[h,w-s],   Bs[h,w] = sum_s bc[h,w,s]        <- this kernel, per pixel
//
// The main kernel marches down a row band like the forward does (same thread tile: 4 columns x 4 disparities, same
// pair-sum ring to recompute exy), keeps a second ring for the vertical k-row sum of a, convolves it horizontally in
// registers (prefix / suffix sums), multiplies by the projector row of the TARGET image row and sums the 16 per-thread
// partials over the 16 lanes of a unit through a shared-memory transpose (fixed summation tree).  Per-unit partial
// sums are added in a fixed order four steps later; every tile stores its T1 rows directly into one of four dense
// images chosen by (band parity, column-tile parity) - tiles of equal parities never overlap - and per-pixel Am / Bs go
// to per-chunk images; sliding_backward_finalize_kernel gathers them.
// The upstream gradient is prefetched with per-thread cp.async (each thread reads back only what it copied, so no
// barrier is involved), up to three row steps ahead.
#include "sliding_common.cuh"

namespace custma {

constexpr int kGradStages = 4;     // ring of upstream-gradient rows in shared memory
constexpr int kGradLookahead = 3;  // steps between issuing a gradient row and using it
constexpr int kBwdLookahead = 4;   // row slots are refilled 4 steps ahead, from the slot every warp released 4 steps ago
                                   // (8 slots): warps may drift up to 4 steps apart before anyone waits

template <int K, int NU, int WG>
struct BwdGeom {
    using F = SlideGeom<K, NU, WG>;
    static constexpr int OFF_PROJ = F::SEG_CAM, OFF_PROJT = OFF_PROJ + F::SEG_PROJ, OFF_A = OFF_PROJT + F::SEG_PROJ,
                         OFF_EX2 = OFF_A + F::SEG_CS, OFF_SP = OFF_EX2 + F::SEG_CS, OFF_EY2 = OFF_SP + F::SEG_PS,
                         OFF_STG = OFF_EY2 + F::SEG_PS;
    static constexpr int UNITS = NU * WG;
    static constexpr int SLOT = OFF_STG + 16 * UNITS;            // + one reduced value per lane of every unit
    static constexpr int CHUNKS = OFF_STG / 4, NCH = (CHUNKS + F::NCONS - 1) / F::NCONS;
    static constexpr int TW = F::WTC + K - 1;                    // columns of a T1 tile row
    static_assert(TW <= 96 && 96 + 2 * F::WTC <= F::NCONS, "reduce_step thread ranges");
    static constexpr int XPOSE_STRIDE = 20;                     // floats per lane row of the transpose scratch
    // per-unit scratch: 16 lane rows + 16 floats of padding - without it the two units of a warp (320 floats apart, a
    // multiple of 32 banks) collide on every scalar read of the transpose (ncu: 44 % of those wavefronts were excessive)
    static constexpr int XPOSE_UNIT = 16 * XPOSE_STRIDE + 16;
    // row ring | gradient ring (+ one stage that stays zero) | transpose scratch (16 lanes x 20 floats per unit)
    static constexpr int GRAD_FLOATS = (kGradStages + 1) * 16 * F::NCONS, XPOSE_FLOATS = UNITS * XPOSE_UNIT;
    static constexpr int SMEM_FLOATS = F::NS * SLOT + GRAD_FLOATS + XPOSE_FLOATS;
    static constexpr size_t SMEM_BYTES = (size_t)SMEM_FLOATS * sizeof(float);
};

// extra workspace of the backward, after the forward-style arrays (SlidingLayout::off_extra)
struct BwdLayout {
    int32_t TW, TH;              // T1 tile: TH = RBH rows x TW columns
    int32_t Hp, Wp;              // T1 images: [B][n_chunks][band parity][tile parity][Hp][Wp], row y + r, column x + r;
                                 // tiles of equal parities never overlap, so every tile stores its rows directly
    size_t off_T1, off_Am, off_Bs, off_patch, total;
};

static void make_bwd_layout(const Problem &p, const SlidingLayout &L, BwdLayout *BL) {
    BL->TW = L.WTC + L.K - 1;
    BL->TH = L.RBH;
    size_t off = L.off_extra;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    BL->Hp = L.NB * L.RB + L.K - 1;
    BL->Wp = L.n_wtiles * L.WTC + L.K - 1;
    BL->off_T1 = take((size_t)p.B * L.n_chunks * 4 * BL->Hp * BL->Wp * sizeof(float));
    const size_t rows = (size_t)p.B * L.n_chunks * L.NB * L.RB;
    BL->off_Am = take(rows * L.cs_pitch * sizeof(float));
    BL->off_Bs = take(rows * L.cs_pitch * sizeof(float));
    BL->off_patch = take((size_t)p.pixels() * p.k * p.k * sizeof(float));   // patch gradients of the flagged tiles
    BL->total = off;
}

// loader of the backward row steps: like RowLoader, plus the projector row of the target image row (k-1 steps old)
template <int K, int NU, int WG>
struct BwdRowLoader {
    using F = SlideGeom<K, NU, WG>;
    using G = BwdGeom<K, NU, WG>;
    const float *src[G::NCH];
    int stride[G::NCH], dst[G::NCH], first[G::NCH], last[G::NCH];
    __device__ __forceinline__ void init(const SlidingLayout &L, const char *ws, int b, int nb, int h0, int w_base,
                                         int s_base, int thread = -1) {
        if (thread < 0) thread = (int)threadIdx.x;   // the thread's index among the copying threads
        const int xlo = w_base - F::r - s_base - F::SC + 1, dlo = w_base - s_base - F::SC + 1;
        const int64_t band = (int64_t)b * L.NB + nb;
        const int64_t srow = ((int64_t)b * L.NB * L.RB + h0) - (K - 1);
#pragma unroll
        for (int n = 0; n < G::NCH; ++n) {
            const int off = 4 * (thread + n * F::NCONS);
            dst[n] = off < G::OFF_STG ? off : -1;
            if (off < G::OFF_PROJ) {
                first[n] = 0; last[n] = L.RBH; stride[n] = L.cam_pitch;
                src[n] = (const float *)(ws + L.off_camP) + band * L.RBH * L.cam_pitch + (w_base / F::WTC) * F::SEG_CAM + off;
            } else if (off < G::OFF_PROJT) {
                first[n] = 0; last[n] = L.RBH; stride[n] = L.proj_pitch;
                src[n] = (const float *)(ws + L.off_projP) + band * L.RBH * L.proj_pitch + (xlo + L.proj_lp) + (off - G::OFF_PROJ);
            } else if (off < G::OFF_A) {
                first[n] = K - 1; last[n] = L.RBH + K - 1; stride[n] = L.proj_pitch;
                src[n] = (const float *)(ws + L.off_projP) + (band * L.RBH - (K - 1)) * L.proj_pitch + (xlo + L.proj_lp) + (off - G::OFF_PROJT);
            } else if (off < G::OFF_EX2) {
                first[n] = K - 1; last[n] = L.RB + K - 1; stride[n] = L.cs_pitch;
                src[n] = (const float *)(ws + L.off_A) + srow * L.cs_pitch + w_base + (off - G::OFF_A);
            } else if (off < G::OFF_SP) {
                first[n] = K - 1; last[n] = L.RB + K - 1; stride[n] = L.cs_pitch;
                src[n] = (const float *)(ws + L.off_ex2) + srow * L.cs_pitch + w_base + (off - G::OFF_EX2);
            } else if (off < G::OFF_EY2) {
                first[n] = K - 1; last[n] = L.RB + K - 1; stride[n] = L.ps_pitch;
                src[n] = (const float *)(ws + L.off_Sp) + srow * L.ps_pitch + (dlo + L.ps_ld) + (off - G::OFF_SP);
            } else {
                first[n] = K - 1; last[n] = L.RB + K - 1; stride[n] = L.ps_pitch;
                src[n] = (const float *)(ws + L.off_ey2) + srow * L.ps_pitch + (dlo + L.ps_ld) + (off - G::OFF_EY2);
            }
        }
    }
    // in order t = 0, 1, 2, ...; the caller has already waited for the slot to be empty
    __device__ __forceinline__ void issue(int t, float *smem, uint64_t *full_bar) {
        const int slot = t & (F::NS - 1);
        float *S = smem + slot * G::SLOT;
#pragma unroll
        for (int n = 0; n < G::NCH; ++n) {
            if (dst[n] >= 0 && t >= first[n] && t < last[n]) cp_async16(S + dst[n], src[n]);
            src[n] += stride[n];
        }
        cp_async_mbar_arrive(&full_bar[slot]);
    }
};

// vertical k-row sum of a value stream through the pair-sum ring (same recurrence as BoxRing, no products)
template <int K>
struct SumRing {
    float prev[4][4];
    float P[K - 2][4][4];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                prev[i][j] = 0.f;
#pragma unroll
                for (int m = 0; m < K - 2; ++m) P[m][i][j] = 0.f;
            }
    }
    __device__ __forceinline__ void step(const int Q, const float (&a)[4][4], float (&sum)[4][4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float s = a[i][j];
#pragma unroll
                for (int m = 1; m <= K - 2; m += 2) s += P[(Q - m + 2 * (K - 2)) % (K - 2)][i][j];
                P[Q][i][j] = a[i][j] + prev[i][j];
                prev[i][j] = a[i][j];
                sum[i][j] = s;
            }
    }
};

// HG: the upstream gradient of a cell is rebuilt from the fused head's per-pixel state (HeadGrad, common.cuh) instead of
// being read from a [.., C] tensor: the "gradient row" the ring prefetches is then one float4 per camera column.
template <int K, int NU, int WG, int MODE, int DIR, bool HG>
__device__ __forceinline__ void backward_consumer(const Problem &p, const SlidingLayout &L, float *smem,
                                                  uint64_t *full_bar, uint64_t *empty_bar,
                                                  BwdRowLoader<K, NU, WG> &loader, int b, int h0, int rows, int w_base,
                                                  int s_base, int steps, const float *__restrict__ grad,
                                                  const HeadGrad &hg, float *__restrict__ T1tile, int T1pitch,
                                                  float *__restrict__ AmRow, float *__restrict__ BsRow) {
    using F = SlideGeom<K, NU, WG>;
    using G = BwdGeom<K, NU, WG>;
    constexpr int CL = F::CL, PL = F::PL, NS = F::NS, PERIOD = F::PERIOD, NT = F::NCONS;
    // gradient row hr is consumed at step hr + K - 1, so it cannot be requested more than K - 1 steps ahead
    constexpr int GLA = kGradLookahead < K - 1 ? kGradLookahead : K - 1;
    const int tid = threadIdx.x;
    const int l16 = tid & 15, u = tid >> 4, su = u % NU, wg = u / NU;
    const int w0 = w_base + 4 * wg, s0 = s_base + 64 * su + 4 * l16;
    const int pidx = 4 * (wg - 16 * su - l16 + 16 * NU - 1);
    const int C = p.C;
    const float seed = kEps / (float)K, inv_n = 1.f / (float)(K * K);
    float *gsm = smem + NS * G::SLOT;   // [kGradStages][4][NT] float4
    // upstream gradient of (row h0, column w0, disparity s0); row hr, column i: + (hr*W + i) * C
    // (the gradient buffer holds volume rows [p.g0, p.g1) only: custma_backward_rows)
    const float *gsrc = HG ? reinterpret_cast<const float *>(hg.state) + (((int64_t)b * p.H + h0) * p.W + w0) * 4
                           : grad + (((int64_t)b * p.grows() + (h0 - p.g0)) * p.W + w0) * C + (MODE == 2 ? 0 : s0);
    const int g_lo = HG ? 0 : max(0, p.g0 - h0), g_hi = HG ? rows : min(rows, p.g1 - h0);   // band rows that carry a gradient
    uint32_t cmask = 0;  // MODE 1: bit 4i+j = cell (w0+i, s0+j) exists and is valid; constant over the band's rows
    if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (w0 + i < p.W && w0 + i - (s0 + j) >= 0) cmask |= 1u << (4 * i + j);
    }
    const int64_t g_row = HG ? (int64_t)p.W * 4 : (int64_t)p.W * C, g_col = HG ? 4 : C;

    float *xps = gsm + G::GRAD_FLOATS + u * G::XPOSE_UNIT;   // this unit's transpose scratch
    const float *gzero = gsm + (kGradStages * 4) * (4 * NT) + 4 * tid;  // a stage that is never written: stays zero

    BoxRing<K> ring;
    SumRing<K> vring;
    ring.clear();
    vring.clear();

    // fixed-order sum of the per-unit partials of step ts (every warp has arrived on its empty barrier) -> workspace
    auto reduce_step = [&](int ts) {
        if (ts < K - 1) return;
        const int ty = ts - (K - 1);
        const float *stg = smem + (ts & (NS - 1)) * G::SLOT + G::OFF_STG;
        if (tid < G::TW) {
            if (ty < L.RBH) {
                const int x = tid;
                float acc = 0.f;
#pragma unroll
                for (int g = 0; g < WG; ++g) {
                    const int xi = x - 4 * g;
                    if (xi >= 0 && xi < K + 3) {
#pragma unroll
                        for (int q = 0; q < NU; ++q) acc += stg[(g * NU + q) * 16 + xi];
                    }
                }
                T1tile[(int64_t)ty * T1pitch + x] = acc;
            }
        } else if (tid >= 96 && tid < 96 + 2 * F::WTC) {
            if (ty < rows) {
                const int idx = tid - 96, w = idx % F::WTC, which = idx / F::WTC;  // 0: Bs, 1: Am
                float acc = 0.f;
#pragma unroll
                for (int q = 0; q < NU; ++q) acc += stg[((w >> 2) * NU + q) * 16 + 8 + 4 * which + (w & 3)];
                (which ? AmRow : BsRow)[(int64_t)ty * L.cs_pitch + w] = which ? acc * inv_n : acc;
            }
        }
    };

    // The row loop is straight-line: every step runs the window ring, the cell epilogue, the vertical ring and the
    // target-row product, also during the k-1 warm-up and k-1 flush steps.  Shared memory was zeroed at kernel start,
    // so steps without a cell row see finite stale statistics and a zero gradient (a = bc = 0), and steps without a
    // target row multiply zeros; no select or branch is needed in the loop body.
#pragma unroll 1
    for (int t0 = 0; t0 < steps; t0 += PERIOD) {
#pragma unroll
        for (int q = 0; q < PERIOD; ++q) {
            const int t = t0 + q;
            const int slot = t & (NS - 1);
            // ---- refill the slot of step t + lookahead (used last by step t + lookahead - 8), after finishing that step's sums
            if (t >= NS - kBwdLookahead) {
                const int ts = t - (NS - kBwdLookahead);
                mbar_wait(&empty_bar[ts & (NS - 1)], (ts / NS) & 1);
                reduce_step(ts);
            }
            if (t + kBwdLookahead < steps) loader.issue(t + kBwdLookahead, smem, full_bar);
            // ---- prefetch the upstream gradient row that step t + GLA consumes
            if (MODE != 2) {
                const int hp = t + GLA - (K - 1);
                if (hp >= g_lo && hp < g_hi) {
                    float *gdst = gsm + ((hp & (kGradStages - 1)) * 4) * (4 * NT) + 4 * tid;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (MODE == 0 || w0 + i < p.W)   // columns past the image stay zero from the initial fill
                            cp_async16(gdst + i * (4 * NT), gsrc + hp * g_row + (int64_t)i * g_col);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            mbar_wait(&full_bar[slot], (t / NS) & 1);
            float *S = smem + slot * G::SLOT;
            const int hr = t - (K - 1);
            const bool has_cells = hr >= g_lo && hr < g_hi;

            float gg[4][4];
            float c[CL], pj[PL];
#pragma unroll
            for (int v = 0; v < CL / 4; ++v)
                *reinterpret_cast<float4 *>(&c[4 * v]) = *reinterpret_cast<const float4 *>(S + 4 * wg + 4 * v);
#pragma unroll
            for (int v = 0; v < PL / 4; ++v)
                *reinterpret_cast<float4 *>(&pj[4 * v]) = *reinterpret_cast<const float4 *>(S + G::OFF_PROJ + pidx + 4 * v);
            float bx[4][4];
            ring.template step<DIR>(q, c, pj, seed, bx);

            float a4[4], e4[4], sp[8], ey[8];
            *reinterpret_cast<float4 *>(a4) = *reinterpret_cast<const float4 *>(S + G::OFF_A + 4 * wg);
            *reinterpret_cast<float4 *>(e4) = *reinterpret_cast<const float4 *>(S + G::OFF_EX2 + 4 * wg);
            *reinterpret_cast<float4 *>(&sp[0]) = *reinterpret_cast<const float4 *>(S + G::OFF_SP + pidx);
            *reinterpret_cast<float4 *>(&sp[4]) = *reinterpret_cast<const float4 *>(S + G::OFF_SP + pidx + 4);
            *reinterpret_cast<float4 *>(&ey[0]) = *reinterpret_cast<const float4 *>(S + G::OFF_EY2 + pidx);
            *reinterpret_cast<float4 *>(&ey[4]) = *reinterpret_cast<const float4 *>(S + G::OFF_EY2 + pidx + 4);
            if (MODE != 2) {
                asm volatile("cp.async.wait_group %0;" ::"n"(GLA) : "memory");
                const float *gs = has_cells ? gsm + ((hr & (kGradStages - 1)) * 4) * (4 * NT) + 4 * tid : gzero;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4 *>(gg[i]) = *reinterpret_cast<const float4 *>(gs + i * (4 * NT));
                if (MODE == 1 && !HG) {   // whatever the caller left in the invalid cells of the gradient must not leak
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (!((cmask >> (4 * i + j)) & 1u)) gg[i][j] = 0.f;
                }
            } else if (HG) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 st = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (has_cells && w0 + i < p.W) st = __ldg(reinterpret_cast<const float4 *>(gsrc + hr * g_row + (int64_t)i * 4));
                    gg[i][0] = st.x; gg[i][1] = st.y; gg[i][2] = st.z; gg[i][3] = st.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int s = s0 + j, d = w0 + i - s;
                        const bool valid = has_cells && w0 + i < p.W && d >= 0 && d < p.W && (!p.banded || s < p.D);
                        gg[i][j] = valid ? __ldg(gsrc + hr * g_row + (int64_t)i * C + (p.banded ? s : d)) : 0.f;
                    }
            }
            float a[4][4];
            float red[16];  // T1[0..8), Bs[0..4), Am[0..4)
            if (has_cells) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float bs = 0.f, am = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int di = i - j + 3;
                        const float e = fmaf(-a4[i], sp[di], bx[i][j]);              // exy + eps
                        const float rs = rsqrt_fast(fmaf(e4[i], ey[di], kEps));      // 1 / den
                        float g = gg[i][j];
                        if (HG) {   // softmax weight of the cell times (s - soft disparity), scaled per pixel (HeadGrad)
                            g = exp2_fast(fmaf(e * rs, hg.beta_log2e, -gg[i][0])) * fmaf((float)(s0 + j), gg[i][1], -gg[i][2]);
                            bool valid = true;
                            if (MODE == 1) valid = (cmask >> (4 * i + j)) & 1u;
                            if (MODE == 2) {
                                const int d = w0 + i - (s0 + j);
                                valid = w0 + i < p.W && d >= 0 && d < p.W && (!p.banded || s0 + j < p.D);
                            }
                            g = valid ? g : 0.f;
                        }
                        const float av = g * rs;                                     // a  = g / den               (:135,:145)
                        a[i][j] = av;                                                // bc = g*ey2*(exy+eps)/den^3 (:147)
                        bs = j == 0 ? (av * e) * (ey[di] * (rs * rs)) : fmaf(av * e, ey[di] * (rs * rs), bs);
                        am = j == 0 ? av * sp[di] : fmaf(av, sp[di], am);
                    }
                    red[8 + i] = bs;
                    red[12 + i] = am;
                }
            } else {   // warm-up and flush steps: no cell row
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
                    red[8 + i] = 0.f;
                    red[12 + i] = 0.f;
                }
            }
            // ---- vertical k-row sum of a, then the target row y = hr - r:  T1[x] += proj'[y, x - s] * sum_w va[w][s]
            float va[4][4];
            vring.step(q, a, va);
            {
                float pt[PL];
#pragma unroll
                for (int v = 0; v < PL / 4; ++v)
                    *reinterpret_cast<float4 *>(&pt[4 * v]) = *reinterpret_cast<const float4 *>(S + G::OFF_PROJT + pidx + 4 * v);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // target column xi (relative to w0 - r) is covered by the cells w_i with xi - (K-1) <= i <= xi:
                    // a prefix sum of va[.][j] while the window is still entering the tile, a suffix sum afterwards
                    float pre[4], suf[4];
                    pre[0] = va[0][j]; pre[1] = pre[0] + va[1][j]; pre[2] = pre[1] + va[2][j]; pre[3] = pre[2] + va[3][j];
                    suf[3] = va[3][j]; suf[2] = suf[3] + va[2][j]; suf[1] = suf[2] + va[1][j]; suf[0] = pre[3];
#pragma unroll
                    for (int xi = 0; xi < K + 3; ++xi) {
                        const float h = xi - (K - 1) <= 0 ? pre[xi < 3 ? xi : 3] : suf[xi - (K - 1)];
                        red[xi] = j == 0 ? pt[xi + 3] * h : fmaf(pt[xi - j + 3], h, red[xi]);
                    }
                }
            }
            // ---- sum the 16 partials over the 16 lanes of the unit through a shared-memory transpose (fixed order):
            //      lane l ends with value #l
            {
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    *reinterpret_cast<float4 *>(xps + l16 * G::XPOSE_STRIDE + 4 * v) =
                        make_float4(red[4 * v], red[4 * v + 1], red[4 * v + 2], red[4 * v + 3]);
                __syncwarp();
                float part[4];   // fixed tree: four chains of four instead of one chain of sixteen
#pragma unroll
                for (int m = 0; m < 4; ++m)
                    part[m] = (xps[(4 * m) * G::XPOSE_STRIDE + l16] + xps[(4 * m + 1) * G::XPOSE_STRIDE + l16]) +
                              (xps[(4 * m + 2) * G::XPOSE_STRIDE + l16] + xps[(4 * m + 3) * G::XPOSE_STRIDE + l16]);
                S[G::OFF_STG + u * 16 + l16] = (part[0] + part[1]) + (part[2] + part[3]);
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty_bar[slot]);
        }
    }
    // the sums of the last steps
    for (int ts = steps - (NS - kBwdLookahead); ts < steps; ++ts) {
        if (ts < 0) continue;
        mbar_wait(&empty_bar[ts & (NS - 1)], (ts / NS) & 1);
        reduce_step(ts);
    }
}


template <int K, int NU, int WG, bool HG>
__global__ void __launch_bounds__(16 * NU * WG, 1)
    sliding_backward_kernel(const Problem p, const SlidingLayout L, const BwdLayout BL, char *__restrict__ ws,
                            const float *__restrict__ grad, const uint32_t tc_threshold, const HeadGrad hg) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    using F = SlideGeom<K, NU, WG>;
    constexpr int WTC = F::WTC, SC = F::SC, NS = F::NS, NCW = F::NCW;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t full_bar[NS], empty_bar[NS];

    // so many flagged tiles that the tensor-core kernel (tc_backward.cu) computes the whole call
    if (*reinterpret_cast<const uint32_t *>(ws + L.off_fb_count) > tc_threshold) return;
    const int tid = threadIdx.x;
    const int wt = blockIdx.x / L.n_chunks, ch = blockIdx.x % L.n_chunks, nb = blockIdx.y, b = blockIdx.z;
    const int w_base = wt * WTC, s_base = chunk_s_base(L, p.W, w_base, ch), h0 = nb * L.RB;
    const int rows = min(L.RB, p.H - h0);
    // row steps: RBH to run every image row through the window ring + k-1 to flush the vertical ring of a
    const int steps = (L.RBH + K - 1 + F::PERIOD - 1) / F::PERIOD * F::PERIOD;
    // this tile's rows of the T1 image of its (band, tile) parity, and its rows of the per-chunk Am / Bs images
    float *T1tile = (float *)(ws + BL.off_T1) +
                    (((((int64_t)b * L.n_chunks + ch) * 2 + (nb & 1)) * 2 + (wt & 1)) * BL.Hp + h0) * BL.Wp + w_base;
    const int64_t prow = (((int64_t)b * L.n_chunks + ch) * L.NB * L.RB + h0) * L.cs_pitch + w_base;
    float *AmRow = (float *)(ws + BL.off_Am) + prow, *BsRow = (float *)(ws + BL.off_Bs) + prow;
    // ill-conditioned tiles belong to the direct two-pass kernels (sliding_fallback.cu): contribute zeros here
    if (reinterpret_cast<const uint8_t *>(ws + L.off_flags)[tile_index(L, b, nb, wt, ch)]) {
        for (int e = tid; e < BL.TH * BL.TW; e += F::NCONS) T1tile[(int64_t)(e / BL.TW) * BL.Wp + e % BL.TW] = 0.f;
        for (int e = tid; e < rows * WTC; e += F::NCONS) {
            AmRow[(int64_t)(e / WTC) * L.cs_pitch + e % WTC] = 0.f;
            BsRow[(int64_t)(e / WTC) * L.cs_pitch + e % WTC] = 0.f;
        }
        return;
    }

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full_bar[i], F::NCONS);
            mbar_init(&empty_bar[i], NCW);
        }
        mbar_fence_init();
    }
    // the row loop relies on never-loaded regions holding finite values and on one gradient stage staying zero
    for (int i = tid; i < BwdGeom<K, NU, WG>::SMEM_FLOATS / 4; i += F::NCONS)
        reinterpret_cast<float4 *>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    BwdRowLoader<K, NU, WG> loader;
    loader.init(L, ws, b, nb, h0, w_base, s_base);
    for (int t = 0; t < kBwdLookahead && t < steps; ++t) loader.issue(t, smem, full_bar);


    // bodies as in the forward: MODE 0 = all cells valid, 1 = 16-byte gradient loads with a validity mask, 2 = scalar;
    // DIR = direction of the sliding horizontal sums (towards the zero padding the tile can see, 0 = no sliding)
    const bool vec = p.banded && (p.D & 3) == 0 && s_base + SC <= p.D;
    const bool clean_left = w_base > 0 && w_base - (s_base + SC - 1) - (K / 2 + 3) >= 0;
    const bool clean_right = w_base + WTC + K <= p.W;
#define CUSTMA_BWD_BODY(MODE, DIR)                                                                                      \
    backward_consumer<K, NU, WG, MODE, DIR, HG>(p, L, smem, full_bar, empty_bar, loader, b, h0, rows, w_base, s_base,  \
                                                steps, grad, hg, T1tile, BL.Wp, AmRow, BsRow)
    if (vec && clean_left && w_base + WTC <= p.W) CUSTMA_BWD_BODY(0, 1);
    else if (vec && clean_right) CUSTMA_BWD_BODY(1, 2);
    else if (vec) CUSTMA_BWD_BODY(1, 0);
    else CUSTMA_BWD_BODY(2, 0);
#undef CUSTMA_BWD_BODY
}

// camera_grad[y,x] = sum of the T1 tiles that cover (y,x) - sum over the k x k cells (h,w) whose window holds (y,x) of
// ( Am[h,w] + Bs[h,w] * (cam'[y,x] - A[h,w]) ), cam' and A relative to the pivot of the band of row h.  Written as
// sum_h ( sum_w q1[h,w] + cam'_h[y,x] * sum_w q2[h,w] ) with q1 = Am - Bs*A, q2 = Bs staged per 32x32 pixel block (four pixels per thread) in
// shared memory.  Fixed summation order; out-of-image cells do not exist, out-of-image targets are never computed
// (reference :177).  Cells of flagged chunks arrive as ready-made patch gradients (sliding_fallback.cu).
constexpr int kFinTX = 32, kFinTY = 8, kFinRows = 32;   // thread block, pixel rows per block (kFinRows / kFinTY per thread)

// x / d for a block-local x: d-aligned base known, a short loop instead of an integer division
__device__ __forceinline__ int div_near(int x, int d, int q0) {   // q0 * d <= x required
    int q = q0;
    while ((q + 1) * d <= x) ++q;
    return q;
}

// ONE: the volume has a single disparity chunk per column tile (banded, D <= 256 - every BASELINE config but the 8K pair):
// the sums over chunks disappear at compile time (the generic loops were unrolled 16 times by the compiler and paid
// their set-up on every one of the six places they appear in)
template <int K, bool ONE>
__global__ void __launch_bounds__(kFinTX * kFinTY)
    sliding_backward_finalize_kernel(const Problem p, const SlidingLayout L, const BwdLayout BL,
                                     const char *__restrict__ ws, float *__restrict__ camera_grad, const uint32_t tc_threshold) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    if (*reinterpret_cast<const uint32_t *>(ws + L.off_fb_count) > tc_threshold) return;
    const int n_chunks = ONE ? 1 : L.n_chunks;
    constexpr int r = K / 2, back = K - 1 - r, SW = kFinTX + K - 1, SH = kFinRows + K - 1;
    // q1 = Am - Bs*A and q2 = Bs of the cells around the block; R* = their horizontal k-sums per (cell row, pixel column)
    __shared__ float q1[SH][SW + 1], q2[SH][SW + 1];
    __shared__ float R1[SH][kFinTX], R2A[SH][kFinTX], R2B[SH][kFinTX];
    const int x0 = blockIdx.x * kFinTX, y0 = blockIdx.y * kFinRows, b = blockIdx.z;
    const int tid = threadIdx.y * kFinTX + threadIdx.x;
    // block-uniform bases (64-bit once); per-thread offsets stay 32-bit (every image of one pair is < 2^31 floats)
    const int chunk_stride = L.NB * L.RB * L.cs_pitch;                      // Am / Bs: [B][n_chunks][NB*RB][cs_pitch]
    const float *Am = (const float *)(ws + BL.off_Am) + (int64_t)b * n_chunks * chunk_stride;
    const float *Bs = (const float *)(ws + BL.off_Bs) + (int64_t)b * n_chunks * chunk_stride;
    const float *A = (const float *)(ws + L.off_A) + (int64_t)b * chunk_stride;
    const float *camP = (const float *)(ws + L.off_camP) + (int64_t)b * L.NB * L.RBH * L.cam_pitch;
    const int img = BL.Hp * BL.Wp;                                          // T1: [B][n_chunks][2][2][Hp][Wp]
    const float *T1 = (const float *)(ws + BL.off_T1) + (int64_t)b * n_chunks * 4 * img;
    const uint8_t *tileany = (const uint8_t *)(ws + L.off_tileany) + (int64_t)b * L.NB * L.n_wtiles;
    // first band / column tile that can hold a cell of this block (the only divisions of the kernel)
    const int nb_lo = max(y0 - back, 0) / L.RB, wt_lo = max(x0 - back, 0) / L.WTC;
    int flagged_near = 0;
    {
        const int nb_hi = min((y0 + kFinRows - 1 + r) / L.RB, L.NB - 1), wt_hi = min((x0 + kFinTX - 1 + r) / L.WTC, L.n_wtiles - 1);
        for (int nb = nb_lo; nb <= nb_hi; ++nb)
            for (int wt = wt_lo; wt <= wt_hi; ++wt) flagged_near |= tileany[nb * L.n_wtiles + wt];
    }
    // ---- stage q1 = Am - Bs * A and q2 = Bs (summed over the chunks; flagged tiles hold zeros) for the cells around
    for (int e = tid; e < SH * SW; e += kFinTX * kFinTY) {
        const int hh = e / SW, ww = e - hh * SW;
        const int h = y0 - back + hh, w = x0 - back + ww;
        float v1 = 0.f, v2 = 0.f;
        if (h >= 0 && h < p.H && w >= 0 && w < p.W) {
            const int at = h * L.cs_pitch + w;
            float am = 0.f, bs = 0.f;
            for (int ch = 0; ch < n_chunks; ++ch) {
                am += Am[at + ch * chunk_stride];
                bs += Bs[at + ch * chunk_stride];
            }
            v1 = fmaf(-bs, A[at], am);
            v2 = bs;
        }
        q1[hh][ww] = v1;
        q2[hh][ww] = v2;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    const int wtA = div_near(x + r, L.WTC, wt_lo), rx = x + r - wtA * L.WTC;
    // ---- horizontal k-sums of the staged rows, once per (cell row, pixel column): R1 = sum_j q1, R2A / R2B = the part
    // of sum_j q2 whose cells lie in column tile wtA / wtA - 1
    for (int hh = threadIdx.y; hh < SH; hh += kFinTY) {
        float s1 = 0.f, s2A = 0.f, s2B = 0.f;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            s1 += q1[hh][threadIdx.x + K - 1 - j];
            if (j <= rx) s2A += q2[hh][threadIdx.x + K - 1 - j];
            else s2B += q2[hh][threadIdx.x + K - 1 - j];
        }
        R1[hh][threadIdx.x] = s1;
        R2A[hh][threadIdx.x] = s2A;
        R2B[hh][threadIdx.x] = s2B;
    }
    __syncthreads();
    if (x >= p.W) return;
    // pivoted camera value of the target pixel in the copies of column tiles wtA and wtA - 1 (offsets within a copy row)
    const int cA = wtA * L.seg_cam + (x - (wtA * L.WTC - r)), cB = (wtA - 1) * L.seg_cam + (x - ((wtA - 1) * L.WTC - r));
    const bool hasA = wtA < L.n_wtiles, hasB = wtA > 0;
    const bool two_x = hasB && rx < K - 1;                                   // T1: the previous column tile's image too
#pragma unroll 1
    for (int m = 0; m < kFinRows / kFinTY; ++m) {
        const int ty = threadIdx.y + m * kFinTY, y = y0 + ty;
        if (y >= p.H) break;
        // ---- T1 images: cell row y + r belongs to band nbA, whose image (parity nbA & 1) holds this target row; so does
        // the previous band's while y + r is inside the k-1 rows the two bands share.  Same along x.  Unwritten parts
        // of the images are never read.
        const int nbA = div_near(y + r, L.RB, nb_lo), ry = y + r - nbA * L.RB;
        float acc = 0.f;
        {
            const int off = (y + r) * BL.Wp + (x + r);
#pragma unroll
            for (int db = 0; db < 2; ++db) {
                const int nb = nbA - db;
                if (nb < 0 || nb >= L.NB || (db && ry >= K - 1)) continue;
#pragma unroll
                for (int dw = 0; dw < 2; ++dw) {
                    const int wt = wtA - dw;
                    if (dw ? !two_x : !hasA) continue;
                    const float *src = T1 + ((nb & 1) * 2 + (wt & 1)) * img + off;
                    for (int ch = 0; ch < n_chunks; ++ch) acc += src[ch * 4 * img];
                }
            }
        }
        // ---- per-pixel terms: cell rows h = y + r - i (band nbA unless i > ry), cell columns w = x + r - j (column tile
        // wtA unless j > rx); cam' of the target pixel relative to the pivot of the cell's own tile: two bands x two tiles
        float cvA0 = 0.f, cvB0 = 0.f, cvA1 = 0.f, cvB1 = 0.f;
        if (nbA < L.NB) {
            const float *crow = camP + (nbA * L.RBH + ry) * L.cam_pitch;
            cvA0 = hasA ? crow[cA] : 0.f;
            cvB0 = hasB ? crow[cB] : 0.f;
        }
        if (nbA >= 1 && ry < K - 1) {
            const float *crow = camP + ((nbA - 1) * L.RBH + ry + L.RB) * L.cam_pitch;
            cvA1 = hasA ? crow[cA] : 0.f;
            cvB1 = hasB ? crow[cB] : 0.f;
        }
        float sub = 0.f;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int h = y - i + r;
            if (h < 0 || h >= p.H) continue;
            const bool own = i <= ry;
            const int hh = ty + K - 1 - i;
            sub += fmaf(own ? cvA0 : cvA1, R2A[hh][threadIdx.x], fmaf(own ? cvB0 : cvB1, R2B[hh][threadIdx.x], R1[hh][threadIdx.x]));
        }
        acc -= sub;
        if (flagged_near) {   // cells of flagged chunks arrive as ready-made patch gradients (reference :172-178 as a gather)
            const float *patch = (const float *)(ws + BL.off_patch) + (int64_t)b * p.H * p.W * (K * K);
            for (int i = 0; i < K; ++i) {
                const int h = y - i + r;
                if (h < 0 || h >= p.H) continue;
                const int nb = i <= ry ? nbA : nbA - 1;
                for (int j = 0; j < K; ++j) {
                    const int w = x - j + r;
                    if (w < 0 || w >= p.W) continue;
                    if (tileany[nb * L.n_wtiles + (j <= rx ? wtA : wtA - 1)])
                        acc += patch[((int64_t)h * p.W + w) * (K * K) + i * K + j];
                }
            }
        }
        camera_grad[((int64_t)b * p.H + y) * p.W + x] = acc;
    }
}

static void launch_finalize(const Problem &p, const SlidingLayout &L, const BwdLayout &BL, const char *ws, float *camera_grad,
                            uint32_t thr, dim3 fgrid, dim3 fblock, cudaStream_t stream) {
    const bool one = L.n_chunks == 1;
    if (p.k == 3) {
        if (one) launch_chained(sliding_backward_finalize_kernel<3, true>, fgrid, fblock, 0, stream, p, L, BL, ws, camera_grad, thr);
        else launch_chained(sliding_backward_finalize_kernel<3, false>, fgrid, fblock, 0, stream, p, L, BL, ws, camera_grad, thr);
    } else {
        if (one) launch_chained(sliding_backward_finalize_kernel<5, true>, fgrid, fblock, 0, stream, p, L, BL, ws, camera_grad, thr);
        else launch_chained(sliding_backward_finalize_kernel<5, false>, fgrid, fblock, 0, stream, p, L, BL, ws, camera_grad, thr);
    }
}

template <int K, int NU, int WG>
static int launch_bwd_cfg(const Problem &p, const SlidingLayout &L, const BwdLayout &BL, char *ws, const float *grad,
                          uint32_t thr, const HeadGrad &hg, cudaStream_t stream) {
    dim3 grid(L.n_wtiles * L.n_chunks, L.NB, p.B);
    const size_t smem = BwdGeom<K, NU, WG>::SMEM_BYTES;
    auto kern = hg.state ? sliding_backward_kernel<K, NU, WG, true> : sliding_backward_kernel<K, NU, WG, false>;
    CUSTMA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUSTMA_CUDA_CHECK(launch_chained(kern, grid, dim3(16 * NU * WG), smem, stream, p, L, BL, ws, grad, thr, hg));
    CUSTMA_LAUNCH_CHECK("sliding_backward_kernel");
    return CUSTMA_OK;
}

template <int K>
static int launch_bwd_k(const SlidingConfig &cfg, const Problem &p, const SlidingLayout &L, const BwdLayout &BL, char *ws,
                        const float *grad, uint32_t thr, const HeadGrad &hg, cudaStream_t stream) {
    switch (cfg.NU) {
        case 1: return launch_bwd_cfg<K, 1, 16>(p, L, BL, ws, grad, thr, hg, stream);
        case 2: return launch_bwd_cfg<K, 2, 8>(p, L, BL, ws, grad, thr, hg, stream);
        case 3: return launch_bwd_cfg<K, 3, 5>(p, L, BL, ws, grad, thr, hg, stream);
        default: return launch_bwd_cfg<K, 4, 4>(p, L, BL, ws, grad, thr, hg, stream);
    }
}

bool sliding_backward_supported(const Problem &p) {
    SlidingConfig cfg;
    return sliding_pick_config(p, true, &cfg);
}

size_t sliding_backward_workspace_bytes(const Problem &p) {
    SlidingConfig cfg;
    if (!sliding_pick_config(p, true, &cfg)) return 0;
    SlidingLayout L;
    make_sliding_layout(p, cfg, true, &L);
    BwdLayout BL;
    make_bwd_layout(p, L, &BL);
    return align256(BL.total) + tc_backward_scratch_bytes(p);
}

int launch_sliding_backward(const Problem &p, const float *grad, const float *cam, const float *proj,
                            float *camera_grad, void *workspace, size_t workspace_bytes, bool force_tensor,
                            cudaStream_t stream, CallPhase phase) {
    SlidingConfig cfg;
    if (!sliding_pick_config(p, true, &cfg)) return set_error(CUSTMA_ERR_UNSUPPORTED, "no sliding-window kernel for k=%d", p.k);
    SlidingLayout L;
    make_sliding_layout(p, cfg, true, &L);
    BwdLayout BL;
    make_bwd_layout(p, L, &BL);
    const size_t need = align256(BL.total) + tc_backward_scratch_bytes(p);
    if (workspace_bytes < need)
        return set_error(CUSTMA_ERR_WORKSPACE, "sliding backward needs %zu workspace bytes, %zu given", need, workspace_bytes);
    char *ws = (char *)workspace;
    float *tc_scratch = (float *)(ws + align256(BL.total));
    if (force_tensor) {
        if (!tc_backward_supported(p))
            return set_error(CUSTMA_ERR_UNSUPPORTED, "CUSTMA_FLAG_TENSOR needs a banded volume with D %% 4 == 0, D <= 540 and k = 3 or 5");
        if (phase == kCallPrepareOnly) return CUSTMA_OK;   // the forced tensor-core kernels prepare nothing
        return launch_tc_backward(p, grad, cam, proj, camera_grad, tc_scratch, nullptr, 0, stream);
    }
    int rc = phase == kCallPrepared ? CUSTMA_OK : launch_sliding_prep(p, L, cam, proj, ws, stream);
    if (rc || phase == kCallPrepareOnly) return rc;
    // same cross-over as the forward (sliding_forward.cu): above this share of flagged work the tensor-core kernel
    // computes the whole call and the sliding / fallback / finalize kernels return at once
    const double items = (double)p.B * L.NB * L.n_wtiles * L.fb_groups;
    const uint32_t thr = tc_backward_supported(p) ? (uint32_t)(0.04 * items) : 0xffffffffu;
    rc = p.k == 3 ? launch_bwd_k<3>(cfg, p, L, BL, ws, grad, thr, HeadGrad(), stream)
                  : launch_bwd_k<5>(cfg, p, L, BL, ws, grad, thr, HeadGrad(), stream);
    if (rc) return rc;
    if ((rc = launch_fallback_patch_grad(p, L, grad, HeadGrad(), cam, proj, ws, (float *)(ws + BL.off_patch), thr, stream))) return rc;
    const dim3 fgrid((p.W + kFinTX - 1) / kFinTX, (p.H + kFinRows - 1) / kFinRows, p.B), fblock(kFinTX, kFinTY);
    launch_finalize(p, L, BL, ws, camera_grad, thr, fgrid, fblock, stream);
    CUSTMA_LAUNCH_CHECK("sliding_backward_finalize_kernel");
    if (thr != 0xffffffffu)
        return launch_tc_backward(p, grad, cam, proj, camera_grad, tc_scratch, (const uint32_t *)(ws + L.off_fb_count), thr, stream);
    return CUSTMA_OK;
}

// ---- fused head: backward without a volume ---------------------------------------------------------------------
// per pixel: (x, y, z, -) = (beta_log2e * max, beta * gd * mask / Z, y * soft, -) from the forward's state and the upstream
// gradient gd of the masked soft disparity; every cell's gradient is then exp2(beta_log2e * c - x) * (s * y - z)
__global__ void __launch_bounds__(256)
    head_grad_prep_kernel(int64_t pixels, const float4 *__restrict__ state, const float *__restrict__ soft_grad, float beta,
                          float4 *__restrict__ out) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= pixels) return;
    const float4 st = state[pix];                      // (beta_log2e * M, 1 / Z, soft, mask)
    const float y = beta * soft_grad[pix] * st.w * st.y;
    out[pix] = make_float4(st.x, y, y * st.z, 0.f);
}

size_t sliding_head_backward_workspace_bytes(const Problem &p) {
    SlidingConfig cfg;
    if (!sliding_pick_config(p, true, &cfg)) return 0;
    SlidingLayout L;
    make_sliding_layout(p, cfg, true, &L);
    BwdLayout BL;
    make_bwd_layout(p, L, &BL);
    return align256(BL.total) + align256((size_t)p.pixels() * sizeof(float4));
}

int launch_sliding_backward_head(const Problem &p, const float *soft_grad, const float *cam, const float *proj,
                                 const float4 *head_state, float beta, float *camera_grad, void *workspace,
                                 size_t workspace_bytes, cudaStream_t stream) {
    SlidingConfig cfg;
    if (!sliding_pick_config(p, true, &cfg))
        return set_error(CUSTMA_ERR_UNSUPPORTED, "the fused head's backward needs kernel_size 3 or 5, got %d", p.k);
    SlidingLayout L;
    make_sliding_layout(p, cfg, true, &L);
    BwdLayout BL;
    make_bwd_layout(p, L, &BL);
    const size_t need = align256(BL.total) + align256((size_t)p.pixels() * sizeof(float4));
    if (workspace_bytes < need)
        return set_error(CUSTMA_ERR_WORKSPACE, "fused head backward needs %zu workspace bytes, %zu given", need, workspace_bytes);
    char *ws = (char *)workspace;
    float4 *hs = (float4 *)(ws + align256(BL.total));
    head_grad_prep_kernel<<<(unsigned)((p.pixels() + 255) / 256), 256, 0, stream>>>(p.pixels(), head_state, soft_grad, beta, hs);
    CUSTMA_LAUNCH_CHECK("head_grad_prep_kernel");
    HeadGrad hg;
    hg.state = hs;
    hg.beta_log2e = beta * 1.4426950408889634f;
    int rc = launch_sliding_prep(p, L, cam, proj, ws, stream);
    if (rc) return rc;
    const uint32_t thr = 0xffffffffu;   // flagged tiles go to the per-cell fallback (the tensor-core kernel reads a volume)
    rc = p.k == 3 ? launch_bwd_k<3>(cfg, p, L, BL, ws, nullptr, thr, hg, stream)
                  : launch_bwd_k<5>(cfg, p, L, BL, ws, nullptr, thr, hg, stream);
    if (rc) return rc;
    if ((rc = launch_fallback_patch_grad(p, L, nullptr, hg, cam, proj, ws, (float *)(ws + BL.off_patch), thr, stream))) return rc;
    const dim3 fgrid((p.W + kFinTX - 1) / kFinTX, (p.H + kFinRows - 1) / kFinRows, p.B), fblock(kFinTX, kFinTY);
    launch_finalize(p, L, BL, ws, camera_grad, thr, fgrid, fblock, stream);
    CUSTMA_LAUNCH_CHECK("sliding_backward_finalize_kernel");
    return CUSTMA_OK;
}

}  // namespace custma
