// Prep kernels of the sliding-window path: they turn the two images into the aligned, padded, pivoted arrays that
// the main kernels stream into shared memory with no bounds checks and no per-element address arithmetic, and they
// decide per tile whether the fast kernels may be used at all.
//
//   pivot_kernel              projector: min / max of the in-image pixels of every row band -> pivot = mid-range;
//                             camera: the same per (band, column tile), over the tile's columns + window halo.
//                             ZNCC is invariant to a constant added to an image (reference kernel.cu:39-70 subtracts
//                             the window mean); subtracting a local constant first keeps the raw products small,
//                             which is what makes the O(1)-per-cell window sums safe in fp32.
//   band_copy_kernel          pivoted copies: out-of-image pixels hold (0 - pivot), i.e. the reference's zero padding
//                             (query_ij, kernel.cu:6-12) in pivoted coordinates; projector rows carry aprons so that
//                             every tile's segment is a 16-byte aligned run, camera rows are stored tile-major
//   band_stats_kernel         per pixel: window sum / mean and centred second moment of the PIVOTED values (so that
//                             exy = sum(cam'*proj') - A*Sp cancels consistently), and the conditioning rho of the window
//   tile_flags_kernel         per tile: can the fp32 error of the raw window sums exceed the tolerance?  Flagged tiles
//                             are queued for the direct-arithmetic fallback kernels (sliding_fallback.cu).
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "sliding_common.cuh"

namespace custma {

bool sliding_pick_config(const Problem &p, bool backward, SlidingConfig *cfg) {
    // instances: forward k = 3, 5, 7; backward k = 3, 5 (its per-thread partial-sum layout holds k + 3 <= 8 columns)
    if (!(p.k == 3 || p.k == 5 || (p.k == 7 && !backward))) return false;
    cfg->K = p.k;
    const int span = p.banded ? p.D : p.W + 32;
    cfg->NU = span <= 64 ? 1 : span <= 128 ? 2 : span <= 192 ? 3 : 4;
    if (!backward) {
        // forward: 6 warps (168 registers) per CTA, two CTAs per SM so that one CTA's start-up hides behind the other
        const int wg[5] = {0, 12, 6, 4, 3};
        cfg->WG = wg[cfg->NU];
    } else {
        // backward: two register rings per cell (window sum + vertical sum of a) need ~250 registers: one CTA of
        // at most 256 threads per SM, so NU * WG <= 16
        const int wg[5] = {0, 16, 8, 5, 4};   // NU = 3: 15 units = 240 threads (the last warp is half full)
        cfg->WG = wg[cfg->NU];
    }
    return true;
}

static int roundup4(int v) { return (v + 3) & ~3; }
static int mod4(int v) { return ((v % 4) + 4) % 4; }

void make_sliding_layout(const Problem &p, const SlidingConfig &cfg, bool backward, SlidingLayout *L) {
    L->K = cfg.K; L->r = cfg.r(); L->NU = cfg.NU; L->WG = cfg.WG; L->WTC = cfg.WTC(); L->SC = cfg.SC();
    L->banded = p.banded; L->smin_full = 0;
    L->n_wtiles = (p.W + L->WTC - 1) / L->WTC;
    L->n_chunks = p.banded ? (p.D + L->SC - 1) / L->SC : (p.W + L->WTC + 3 + L->SC - 1) / L->SC;
    // Rows per band.  Every band recomputes k-1 warm-up rows (twice that in the backward), so use as few bands as
    // possible, of equal height, while one pair alone still gives two waves of resident CTAs (forward: 2 CTAs per SM,
    // backward: 1).  Independent of B, so that a batch and its pairs one by one give the same bits.
    // RB + K - 1 row steps must be a whole number of pair-sum-ring periods (K - 2).
    const int64_t tiles = (int64_t)L->n_wtiles * L->n_chunks;
    const int64_t sms = device_sm_count();
    const int64_t want = backward ? 2 * sms : 4 * sms;
    auto fit_up = [&](int rows) { while ((rows + cfg.K - 1) % (cfg.K - 2)) ++rows; return rows; };
    int RB = 16;
    for (int nbands = 1; nbands <= p.H; ++nbands) {
        const int rb = fit_up((p.H + nbands - 1) / nbands);
        if (rb > 128) continue;
        RB = rb;
        if (nbands * tiles >= want || rb <= 16) break;
    }
    if (backward) {
        // One CTA per SM: a pair takes ceil(CTAs / SMs) rounds of RB + 2 (k-1) row steps.  Among the band heights whose
        // estimate is within 12 % of the best take the tallest (fewest warm-up rows - what a batch, whose CTAs fill the
        // rounds anyway, wants).  Measured, 1242x375: 4 bands instead of 5 -> one pair 0.330 -> 0.299 ms, 8 pairs 1.620 ->
        // 1.586 ms (7 bands: 0.284 / 1.675 ms).
        auto cost = [&](int rb) {
            const int64_t ctas = (int64_t)((p.H + rb - 1) / rb) * tiles;
            return ((ctas + sms - 1) / sms) * (rb + 2 * (cfg.K - 1));
        };
        int64_t best = INT64_MAX;
        for (int nbands = 1; nbands <= p.H; ++nbands) {
            const int rb = fit_up((p.H + nbands - 1) / nbands);
            if (rb > 128) continue;
            if (rb < 8) break;
            best = std::min(best, cost(rb));
        }
        for (int nbands = 1; nbands <= p.H && best != INT64_MAX; ++nbands) {
            const int rb = fit_up((p.H + nbands - 1) / nbands);
            if (rb > 128) continue;
            if (rb < 8) break;
            if (cost(rb) * 100 <= best * 112) { RB = rb; break; }
        }
    }
    if (const char *e = getenv(backward ? "CUSTMA_BANDS_BWD" : "CUSTMA_BANDS")) {   // experiment: force the number of bands
        const int nb = atoi(e);
        if (nb > 0) RB = std::min(128, fit_up((p.H + nb - 1) / nb));
    }
    L->RB = RB; L->RBH = RB + cfg.K - 1; L->NB = (p.H + RB - 1) / RB;
    L->seg_cam = cfg.seg_cam(); L->seg_proj = cfg.seg_proj(); L->seg_cs = cfg.seg_cs(); L->seg_ps = cfg.seg_ps();
    L->slot_floats = L->seg_cam + L->seg_proj + 2 * L->seg_cs + 2 * L->seg_ps;

    L->cam_lc = 0;
    L->cam_pitch = L->n_wtiles * L->seg_cam;                          // tile-major, seg_cam is a multiple of 4
    int min_xlo = 0, max_xhi = 0, min_dlo = 0, max_dhi = 0;
    for (int wt = 0; wt < L->n_wtiles; ++wt)
        for (int c = 0; c < L->n_chunks; ++c) {
            const int w_base = wt * L->WTC, s_base = chunk_s_base(*L, p.W, w_base, c);
            const int xlo = w_base - L->r - s_base - L->SC + 1, dlo = w_base - s_base - L->SC + 1;
            min_xlo = std::min(min_xlo, xlo); max_xhi = std::max(max_xhi, xlo + L->seg_proj);
            min_dlo = std::min(min_dlo, dlo); max_dhi = std::max(max_dhi, dlo + L->seg_ps);
        }
    // (xlo + lp) % 4 == 0 with xlo == 1 - r (mod 4);  (dlo + ld) % 4 == 0 with dlo == 1 (mod 4)
    int lp = -min_xlo; lp += mod4(L->r - 1 - lp);
    int ld = -min_dlo; ld += mod4(3 - ld);
    L->proj_lp = lp; L->proj_pitch = roundup4(lp + max_xhi);
    L->ps_ld = ld; L->ps_pitch = roundup4(ld + max_dhi);
    L->cs_pitch = L->n_wtiles * L->WTC;

    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t bands = (size_t)p.B * L->NB, rows = (size_t)p.B * L->NB * L->RB;
    L->nblk_cs = (L->cs_pitch + 15) / 16; L->nblk_ps = (L->ps_pitch + 15) / 16;
    // zero-initialised region (one memset): band min/max accumulators, WTA keys, worst window conditioning per block
    L->off_minmax = take(2 * bands * 2 * sizeof(uint32_t));            // [img][pair*band][max(v), max(-v)] ordered
    L->off_wta = take((size_t)p.pixels() * sizeof(unsigned long long));  // packed (best, s) keys
    L->off_rho_c = take(bands * L->n_wtiles * sizeof(float));          // per column tile
    L->off_rho_p = take(bands * L->nblk_ps * sizeof(float));
    L->off_bandany = take(bands * sizeof(uint32_t));                    // does the band hold a flagged tile?
    L->off_fb_count = take(sizeof(uint32_t));                          // number of fallback work items
    L->zero_end = off;
    L->off_flags = take((size_t)p.B * L->NB * L->n_wtiles * L->n_chunks);
    L->off_tileany = take((size_t)p.B * L->NB * L->n_wtiles);
    L->off_campiv = take((size_t)p.B * L->NB * L->n_wtiles * sizeof(float));
    L->off_camP = take(bands * L->RBH * L->cam_pitch * sizeof(float));
    L->off_projP = take(bands * L->RBH * L->proj_pitch * sizeof(float));
    L->off_A = take(rows * L->cs_pitch * sizeof(float));
    L->off_ex2 = take(rows * L->cs_pitch * sizeof(float));
    L->off_Sp = take(rows * L->ps_pitch * sizeof(float));
    L->off_ey2 = take(rows * L->ps_pitch * sizeof(float));
    L->fb_groups = (L->RB + kFallbackRows - 1) / kFallbackRows;
    L->off_fb_list = take((size_t)p.B * L->NB * L->n_wtiles * L->fb_groups * sizeof(uint32_t));
    L->off_fb_pm = take((size_t)p.pixels() * sizeof(float));
    L->off_fb_ey2 = take((size_t)p.pixels() * sizeof(float));
    L->off_extra = off;
    L->total = off;
}

// Host-side self-check of the tiling: every segment a kernel copies out of the workspace arrays must lie inside its
// row and be 16-byte aligned, for every tile of the problem.  (compute-sanitizer is not available on the test pool;
// tests/test_layout_cpu.py sweeps shapes through this instead.)
int validate_sliding_layout(const Problem &p, bool backward) {
    SlidingConfig cfg;
    if (!sliding_pick_config(p, backward, &cfg)) return CUSTMA_OK;   // no fast-path instance: nothing to check
    SlidingLayout L;
    make_sliding_layout(p, cfg, backward, &L);
    auto bad = [&](const char *what, int wt, int c, int a, int b) {
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "layout check failed: %s (tile %d chunk %d: %d vs %d) for B=%d H=%d W=%d D=%d k=%d %s",
                         what, wt, c, a, b, p.B, p.H, p.W, p.D, p.k, backward ? "backward" : "forward");
    };
    if (L.RB < 1 || (L.RB + L.K - 1) % (L.K - 2) != 0) return bad("band height", 0, 0, L.RB, L.K);
    if (L.NB * L.RB < p.H || (L.NB - 1) * L.RB >= p.H) return bad("band count", 0, 0, L.NB * L.RB, p.H);
    if ((L.cam_pitch | L.proj_pitch | L.cs_pitch | L.ps_pitch) & 3) return bad("pitch alignment", 0, 0, L.cam_pitch, L.proj_pitch);
    if (L.n_wtiles * L.WTC < p.W) return bad("column tiles", 0, 0, L.n_wtiles * L.WTC, p.W);
    for (int wt = 0; wt < L.n_wtiles; ++wt) {
        const int w_base = wt * L.WTC;
        const int c0 = wt * L.seg_cam;
        if ((c0 & 3) || c0 + L.seg_cam > L.cam_pitch) return bad("camera segment", wt, 0, c0 + L.seg_cam, L.cam_pitch);
        if (L.seg_cam < L.WTC + L.K - 1) return bad("camera segment width", wt, 0, L.seg_cam, L.WTC + L.K - 1);
        if (w_base + L.seg_cs > L.cs_pitch) return bad("camera statistics segment", wt, 0, w_base + L.seg_cs, L.cs_pitch);
        int s_lo = 1 << 30, s_hi = -(1 << 30);
        for (int c = 0; c < L.n_chunks; ++c) {
            const int s_base = chunk_s_base(L, p.W, w_base, c);
            const int x0 = w_base - L.r - s_base - L.SC + 1 + L.proj_lp, d0 = w_base - s_base - L.SC + 1 + L.ps_ld;
            if (s_base & 3) return bad("chunk alignment", wt, c, s_base, 4);
            if (x0 < 0 || (x0 & 3) || x0 + L.seg_proj > L.proj_pitch) return bad("projector segment", wt, c, x0 + L.seg_proj, L.proj_pitch);
            if (d0 < 0 || (d0 & 3) || d0 + L.seg_ps > L.ps_pitch) return bad("projector statistics segment", wt, c, d0 + L.seg_ps, L.ps_pitch);
            s_lo = std::min(s_lo, s_base); s_hi = std::max(s_hi, s_base + L.SC - 1);
        }
        // the chunks of a column tile cover every disparity any of its columns needs
        const int need_lo = p.banded ? 0 : w_base - (p.W - 1), need_hi = p.banded ? p.D - 1 : std::min(w_base + L.WTC, p.W) - 1;
        if (s_lo > need_lo || s_hi < need_hi) return bad("disparity coverage", wt, 0, s_lo, s_hi);
    }
    // the column of the copies that holds image column X = -r .. W-1+(K-1-r) of every statistics window exists
    if (L.proj_lp < L.r) return bad("left apron", 0, 0, L.proj_lp, L.r);
    const size_t offs[] = {L.off_minmax, L.off_wta, L.off_rho_c, L.off_rho_p, L.off_bandany, L.off_fb_count, L.zero_end, L.off_flags, L.off_tileany, L.off_campiv, L.off_camP, L.off_projP, L.off_A, L.off_ex2, L.off_Sp,
                           L.off_ey2, L.off_fb_list, L.off_fb_pm, L.off_fb_ey2, L.off_extra, L.total};
    for (size_t i = 1; i < sizeof(offs) / sizeof(offs[0]); ++i)
        if (offs[i] < offs[i - 1] || (offs[i] & 255)) return bad("workspace offsets", (int)i, 0, (int)(offs[i] >> 8), (int)(offs[i - 1] >> 8));
    return CUSTMA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// camera pivot per (pair, band, column tile): mid-range of the in-image pixels of the tile's footprint (its columns plus
// the window halo, the band's rows plus halo).  A block takes a run of whole tiles: one thread per image column marches
// down the band's rows (coalesced, each pixel read once per band), the per-column extrema go to shared memory and one
// thread per tile combines its columns.  (One warp per tile, each reading its own 20-column footprint, was 25-36 us per
// call: latency-bound and uncoalesced.)
constexpr int kPivotThreads = 256;
__global__ void __launch_bounds__(kPivotThreads)
    pivot_kernel(Problem p, SlidingLayout L, int tiles_per_block, const float *__restrict__ cam,
                 const float *__restrict__ proj, float *__restrict__ campiv, uint32_t *__restrict__ minmax) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    __shared__ float cmax[kPivotThreads], cmin[kPivotThreads];
    const bool is_proj = (int)blockIdx.z >= p.B;        // z >= B: the projector's band extrema (one pivot per band)
    const int nb = blockIdx.y, b = blockIdx.z % p.B, wt0 = blockIdx.x * tiles_per_block;
    const float *plane = (is_proj ? proj : cam) + (int64_t)b * p.H * p.W;
    const int y0 = max(0, nb * L.RB - L.r), y1 = min(p.H, nb * L.RB + L.RBH - L.r);
    const int xb = wt0 * L.WTC - (is_proj ? 0 : L.r);                   // image column of thread 0
    const int ncols = min(tiles_per_block, L.n_wtiles - wt0) * L.WTC + (is_proj ? 0 : L.K - 1);
    const int x = xb + (int)threadIdx.x;
    float vmax = -INFINITY, vmin = INFINITY;
    if ((int)threadIdx.x < ncols && x >= 0 && x < p.W) {
        const float *col = plane + x;
#pragma unroll 8
        for (int y = y0; y < y1; ++y) {
            const float v = __ldg(col + (int64_t)y * p.W);
            vmax = fmaxf(vmax, v);
            vmin = fminf(vmin, v);
        }
    }
    if (is_proj) {
        // block-reduce, then one pair of atomics per block (non-negative keys: float_to_ordered is monotonic)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        }
        if ((threadIdx.x & 31) == 0) { cmax[threadIdx.x >> 5] = vmax; cmin[threadIdx.x >> 5] = vmin; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int wi = 1; wi < kPivotThreads / 32; ++wi) { vmax = fmaxf(vmax, cmax[wi]); vmin = fminf(vmin, cmin[wi]); }
            if (vmax >= vmin) {
                uint32_t *mm = minmax + ((size_t)(p.B + b) * L.NB + nb) * 2;
                atomicMax(mm, float_to_ordered(vmax));
                atomicMax(mm + 1, float_to_ordered(-vmin));
            }
        }
        return;
    }
    cmax[threadIdx.x] = vmax;
    cmin[threadIdx.x] = vmin;
    __syncthreads();
    const int t = threadIdx.x;
    if (t < tiles_per_block && wt0 + t < L.n_wtiles) {
        float hi = -INFINITY, lo = INFINITY;
        for (int c = t * L.WTC; c < t * L.WTC + L.WTC + L.K - 1; ++c) {
            hi = fmaxf(hi, cmax[c]);
            lo = fminf(lo, cmin[c]);
        }
        const float pv = 0.5f * (hi + lo);
        campiv[((int64_t)b * L.NB + nb) * L.n_wtiles + wt0 + t] = (hi >= lo && isfinite(pv)) ? pv : 0.f;
    }
}

__device__ __forceinline__ float band_pivot(const uint32_t *__restrict__ minmax, const Problem &p,
                                            const SlidingLayout &L, int img, int b, int nb) {
    const uint32_t *mm = minmax + ((size_t)(img * p.B + b) * L.NB + nb) * 2;
    const uint32_t hi = mm[0], lo = mm[1];
    if (hi == 0u || lo == 0u) return 0.f;
    const float pv = 0.5f * (ordered_to_float(hi) - ordered_to_float(lo));
    return isfinite(pv) ? pv : 0.f;
}

// four consecutive elements of the two band copies per thread
__global__ void __launch_bounds__(256)
    band_copy_kernel(Problem p, SlidingLayout L, const float *__restrict__ cam, const float *__restrict__ proj,
                     const uint32_t *__restrict__ minmax, const float *__restrict__ campiv,
                     float *__restrict__ camP, float *__restrict__ projP) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    const int img = blockIdx.z / p.B, b = blockIdx.z % p.B;
    const int pitch = img ? L.proj_pitch : L.cam_pitch, left = L.proj_lp;
    const int t = blockIdx.y % L.RBH, nb = blockIdx.y / L.RBH;
    const int y = nb * L.RB - L.r + t;
    const float pvb = img ? band_pivot(minmax, p, L, 1, b, nb) : 0.f;
    const float *row = (img ? proj : cam) + ((int64_t)b * p.H + y) * p.W;
    float *out = (img ? projP : camP) + (((int64_t)b * L.NB + nb) * L.RBH + t) * pitch;
    const bool yin = y >= 0 && y < p.H;
    const int ci = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (ci < pitch) {
        float v[4];
        // projector: column X sits at index X + proj_lp, one pivot per band; camera: tile-major (tile wt holds image
        // columns wt*WTC - r + j at index wt*seg_cam + j), one pivot per tile (4 consecutive indices share a tile)
        const int wt = img ? 0 : ci / L.seg_cam;
        const float pv = img ? pvb : campiv[((int64_t)b * L.NB + nb) * L.n_wtiles + wt];
        const int xbase = img ? ci - left : wt * L.WTC - L.r + (ci - wt * L.seg_cam);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int x = xbase + e;
            v[e] = ((yin && x >= 0 && x < p.W) ? __ldg(row + x) : 0.f) - pv;
        }
        *reinterpret_cast<float4 *>(out + ci) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// Window statistics of the pivoted band copies: one thread per column marches down the band keeping the last k-1
// horizontal k-sums of v and v*v in registers (O(k) loads per pixel instead of k*k).  Every thread sums its own
// windows, so it may shift the data by its own constant: it uses a pixel next to its first window, which keeps the
// one-pass second moment e2 = sum u^2 - (sum u)^2 / n (u = v - shift) accurate in fp32 however far the band pivot is
// from the local brightness; the window sum relative to the band pivot is recovered as sum u + n * shift.
// camera: A = window mean, ex2;   projector: Sp = window sum, ey2.
// Conditioning of a window: rho = n * max|v|^2 / e2 (how much larger the raw sums are than what survives the
// cancellation); the worst rho per block of 16 columns feeds the tile verdict.  rho = 0 for an exactly constant window
// of pivoted zeros, infinite for any other flat window.
// Output rows per thread of band_stats_kernel (+ k-1 warm-up rows): 24, or 12 / 8 when a call is too small to fill the
// device with 24 (one KITTI pair: 24 -> 11 us).  A thread's shift comes from the 24-row block its rows lie in, so the
// statistics do not depend on the choice by a bit.
constexpr int kStatRows = 24;

template <int K>
__global__ void __launch_bounds__(128)
    band_stats_kernel(Problem p, SlidingLayout L, const float *__restrict__ camP, const float *__restrict__ projP,
                      float *__restrict__ A, float *__restrict__ ex2, float *__restrict__ Sp,
                      float *__restrict__ ey2, float *__restrict__ rho_c, float *__restrict__ rho_p, const int stat_rows) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    // blockIdx.y = band * segments + segment: a thread marches stat_rows output rows (+ k-1 warm-up rows) of one column
    const int nseg = (L.RB + stat_rows - 1) / stat_rows;
    const int img = blockIdx.z / p.B, b = blockIdx.z % p.B, nb = blockIdx.y / nseg, seg = blockIdx.y % nseg;
    const int t_begin = seg * stat_rows, t_end = min(L.RBH, t_begin + stat_rows + K - 1);
    const int t_shift = t_begin / kStatRows * kStatRows;   // stat_rows divides kStatRows: segments nest in 24-row blocks
    const int pitch = img ? L.ps_pitch : L.cs_pitch, left = img ? L.ps_ld : 0;
    const int pitchP = img ? L.proj_pitch : L.cam_pitch;
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    const int x = ci - left;
    const bool col_ok = ci < pitch && x >= 0 && x < p.W;
    // column of the copy that holds image column x - r (camera copies are tile-major), and how far the copy extends
    // to either side of the window (the camera's ends with its tile)
    const int wt = x / L.WTC, xin = x - wt * L.WTC;
    const int c0 = !col_ok ? 0 : img ? x - L.r + L.proj_lp : wt * L.seg_cam + xin;
    const int room_l = img ? 3 : min(3, xin), room_r = img ? 3 : min(3, L.seg_cam - (xin + K));
    // in-image neighbours of the window inside the copy, to the left / right (0..3): constant per thread
    const int nl = max(0, min(room_l, x - L.r)), nr = max(0, min(room_r, p.W - (x - L.r + K)));
    const float *src = (img ? projP : camP) + ((int64_t)b * L.NB + nb) * L.RBH * pitchP + c0;
    float *o1 = (img ? Sp : A) + ((int64_t)b * L.NB + nb) * L.RB * pitch + ci;
    float *o2 = (img ? ey2 : ex2) + ((int64_t)b * L.NB + nb) * L.RB * pitch + ci;
    const float inv_n = 1.f / (float)(K * K);
    float r1[K - 1], r2[K - 1], rm[K - 1];
#pragma unroll
    for (int m = 0; m < K - 1; ++m) r1[m] = r2[m] = rm[m] = 0.f;
    const float shift = col_ok ? src[(int64_t)min(t_shift + K / 2, L.RBH - 1) * pitchP + K / 2] : 0.f;
    float rho = 0.f;
    for (int t = t_begin; t < t_end; ++t) {
        float h1 = 0.f, h2 = 0.f, hm = 0.f;
        if (col_ok) {
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const float v = src[(int64_t)t * pitchP + j], u = v - shift;
                h1 += u;
                h2 = fmaf(u, u, h2);
                hm = fmaxf(hm, fabsf(v));
            }
            // the kernels slide the horizontal sum over a thread's 4 columns (add the entering product, subtract the
            // leaving one), so a window also inherits the rounding of its up to 3 in-image neighbours on either side
            // (tiles that can see the zero padding do not slide): take the magnitude over that span
#pragma unroll
            for (int j = 1; j <= 3; ++j) {
                if (j <= nl) hm = fmaxf(hm, fabsf(src[(int64_t)t * pitchP - j]));
                if (j <= nr) hm = fmaxf(hm, fabsf(src[(int64_t)t * pitchP + K - 1 + j]));
            }
        }
        float s1 = h1, s2 = h2, sm = hm;
#pragma unroll
        for (int m = 0; m < K - 1; ++m) { s1 += r1[m]; s2 += r2[m]; sm = fmaxf(sm, rm[m]); }
#pragma unroll
        for (int m = 0; m < K - 2; ++m) { r1[m] = r1[m + 1]; r2[m] = r2[m + 1]; rm[m] = rm[m + 1]; }
        r1[K - 2] = h1; r2[K - 2] = h2; rm[K - 2] = hm;
        const int hr = t - (K - 1);
        if (hr >= t_begin && ci < pitch) {
            const bool ok = col_ok && nb * L.RB + hr < p.H;
            const float e2 = fmaxf(fmaf(-s1, s1 * inv_n, s2), 0.f);
            const float sum = fmaf((float)(K * K), shift, s1);   // window sum of v (relative to the band pivot)
            // neutral values elsewhere: they only ever meet masked cells
            o1[(int64_t)hr * pitch] = ok ? (img ? sum : sum * inv_n) : 0.f;
            o2[(int64_t)hr * pitch] = ok ? e2 : 1.f;
            if (ok && sm > 0.f) rho = fmaxf(rho, __fdividef((float)(K * K) * sm * sm, e2));   // e2 == 0 -> inf; a verdict, not a result
        }
    }
    // non-negative floats (and +inf) order like unsigned integers
    if (img) {   // projector: worst window per block of 16 statistics columns
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) rho = fmaxf(rho, __shfl_xor_sync(0xffffffffu, rho, o));
        if ((threadIdx.x & 15) == 0 && ci < pitch)
            atomicMax(reinterpret_cast<uint32_t *>(rho_p + ((int64_t)b * L.NB + nb) * L.nblk_ps + ci / 16), __float_as_uint(rho));
    } else if (col_ok && rho > 0.f) {   // camera: worst window per column tile
        atomicMax(reinterpret_cast<uint32_t *>(rho_c + ((int64_t)b * L.NB + nb) * L.n_wtiles + wt), __float_as_uint(rho));
    }
}

// Verdict per tile.  The raw window sum of a cell loses, relative to the normalisation sqrt(ex2 * ey2), about
//   eps_f32 * c * n * max|cam'| * max|proj'| / sqrt(ex2 * ey2) = eps_f32 * c * sqrt(rho_cam * rho_proj),
// c ~ 1 for the statistical worst case over millions of cells (measured: 5e-7 at sqrt(rho_cam * rho_proj) = 6).
// Tiles where the worst camera window (over the tile) and the worst projector window (over the tile's disparity
// footprint) keep this below ~3e-6 run the fast kernels; the others (low contrast against the band pivot, flat
// regions) are left to the direct two-pass kernels, which follow the reference's arithmetic order.
constexpr float kMaxConditioning = 40.f;
// 3 x 3 windows: nine terms average their rounding less (fuzzing found 8e-6 of cost scale just under 40), so they get
// half the budget
constexpr float kMaxConditioningK3 = 20.f;

__global__ void __launch_bounds__(128)
    tile_flags_kernel(Problem p, SlidingLayout L, const float *__restrict__ rho_c, const float *__restrict__ rho_p,
                      uint8_t *__restrict__ flags, uint8_t *__restrict__ tileany, uint32_t *__restrict__ bandany,
                      uint32_t *__restrict__ fb_count, uint32_t *__restrict__ fb_list) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t ntiles = (int64_t)p.B * L.NB * L.n_wtiles;
    if (id >= ntiles) return;
    const int wt = (int)(id % L.n_wtiles), nb = (int)((id / L.n_wtiles) % L.NB), b = (int)(id / ((int64_t)L.n_wtiles * L.NB));
    const int64_t band = (int64_t)b * L.NB + nb;
    const int w_base = wt * L.WTC;
    auto range_max = [](const float *a, int lo, int hi) { float m = 0.f; for (int i = lo / 16; i <= (hi - 1) / 16; ++i) m = fmaxf(m, a[i]); return m; };
    const float rc = rho_c[band * L.n_wtiles + wt];
    uint8_t any = 0;
    for (int ch = 0; ch < L.n_chunks; ++ch) {
        const int s_base = chunk_s_base(L, p.W, w_base, ch);
        const int dlo = w_base - s_base - L.SC + 1;
        const float rp = range_max(rho_p + band * L.nblk_ps, dlo + L.ps_ld, dlo + L.ps_ld + L.seg_ps);
        const bool fast = sqrtf(rc) * sqrtf(rp) <= (L.K == 3 ? kMaxConditioningK3 : kMaxConditioning);   // (not sqrtf(rc * rp): inf * 0 must flag)
        flags[id * L.n_chunks + ch] = fast ? 0 : 1;
        any |= fast ? 0 : 1;
    }
    tileany[id] = any;
    if (any) {   // queue the tile for the fallback kernels, kFallbackRows rows per work item
        atomicOr(bandany + band, 1u);
        const uint32_t first = atomicAdd(fb_count, (uint32_t)L.fb_groups);
        for (int g = 0; g < L.fb_groups; ++g) fb_list[first + g] = (uint32_t)(id * L.fb_groups + g);
    }
}

int launch_sliding_prep(const Problem &p, const SlidingLayout &L, const float *cam, const float *proj, char *ws,
                        cudaStream_t stream) {
    uint32_t *minmax = (uint32_t *)(ws + L.off_minmax);
    float *camP = (float *)(ws + L.off_camP), *projP = (float *)(ws + L.off_projP), *campiv = (float *)(ws + L.off_campiv);
    float *rho_c = (float *)(ws + L.off_rho_c), *rho_p = (float *)(ws + L.off_rho_p);
    CUSTMA_CUDA_CHECK(cudaMemsetAsync(ws + L.off_minmax, 0, L.zero_end - L.off_minmax, stream));
    {
        const int tiles_per_block = std::max(1, (kPivotThreads - (L.K - 1)) / L.WTC);
        dim3 pgrid((L.n_wtiles + tiles_per_block - 1) / tiles_per_block, L.NB, 2 * p.B);
        pivot_kernel<<<pgrid, kPivotThreads, 0, stream>>>(p, L, tiles_per_block, cam, proj, campiv, minmax);
        CUSTMA_LAUNCH_CHECK("pivot_kernel");
    }
    {
        dim3 grid((std::max(L.cam_pitch, L.proj_pitch) / 4 + 255) / 256, L.NB * L.RBH, 2 * p.B);
        CUSTMA_CUDA_CHECK(launch_chained(band_copy_kernel, grid, dim3(256), 0, stream, p, L, cam, proj, (const uint32_t *)minmax,
                                         (const float *)campiv, camP, projP));
        CUSTMA_LAUNCH_CHECK("band_copy_kernel");
    }
    {
        // one wave of 128-thread blocks at 16 per SM
        const int64_t xblocks = (std::max(L.cs_pitch, L.ps_pitch) + 127) / 128, wave = 16 * (int64_t)device_sm_count();
        int stat_rows = kStatRows;
        for (int r : {12, 8})
            if (xblocks * L.NB * ((L.RB + stat_rows - 1) / stat_rows) * 2 * p.B < wave) stat_rows = r;
        dim3 grid((unsigned)xblocks, L.NB * ((L.RB + stat_rows - 1) / stat_rows), 2 * p.B);
        auto kern = p.k == 3 ? band_stats_kernel<3> : p.k == 5 ? band_stats_kernel<5> : band_stats_kernel<7>;
        CUSTMA_CUDA_CHECK(launch_chained(kern, grid, dim3(128), 0, stream, p, L, (const float *)camP, (const float *)projP,
                                         (float *)(ws + L.off_A), (float *)(ws + L.off_ex2), (float *)(ws + L.off_Sp),
                                         (float *)(ws + L.off_ey2), rho_c, rho_p, stat_rows));
        CUSTMA_LAUNCH_CHECK("band_stats_kernel");
    }
    {
        const int64_t ntiles = (int64_t)p.B * L.NB * L.n_wtiles;
        CUSTMA_CUDA_CHECK(launch_chained(tile_flags_kernel, dim3((unsigned)((ntiles + 127) / 128)), dim3(128), 0, stream, p, L,
                                         (const float *)rho_c, (const float *)rho_p, (uint8_t *)(ws + L.off_flags),
                                         (uint8_t *)(ws + L.off_tileany), (uint32_t *)(ws + L.off_bandany),
                                         (uint32_t *)(ws + L.off_fb_count), (uint32_t *)(ws + L.off_fb_list)));
        CUSTMA_LAUNCH_CHECK("tile_flags_kernel");
    }
    return CUSTMA_OK;
}

}  // namespace custma
