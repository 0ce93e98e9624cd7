// Prep kernels of the sliding-window path: they turn the two images into the aligned, padded, pivoted arrays that
// the main kernels stream through TMA bulk copies with no bounds checks and no per-element address arithmetic.
//
//   band_minmax_kernel   per (image, pair, row band): min / max of the in-image pixels -> pivot = mid-range.
//                        ZNCC is invariant to a constant added to an image (reference kernel.cu:39-70 subtracts the
//                        window mean); subtracting a band-local constant first keeps the raw products small, which is
//                        what makes the O(1)-per-cell window sums safe in fp32.
//   band_copy_kernel     pivoted copies with aprons: out-of-image pixels hold (0 - pivot), i.e. the reference's zero
//                        padding (query_ij, kernel.cu:6-12) in pivoted coordinates.
//   band_stats_kernel    per pixel: window sum and centred second moment of the PIVOTED values, two passes in the
//                        reference's order (kernel.cu:40-70).  Statistics are taken from the shifted data so that
//                        exy = sum(cam'*proj') - A*Sp cancels consistently.
#include <algorithm>

#include "sliding_common.cuh"

namespace custma {

bool sliding_pick_config(const Problem &p, SlidingConfig *cfg) {
    if (p.k != 5) return false;
    cfg->K = p.k;
    const int span = p.banded ? p.D : p.W + 32;
    // 6 consumer warps + 1 producer warp per CTA, two CTAs per SM (one CTA's start-up hides behind the other)
    if (span <= 64) { cfg->NU = 1; cfg->WG = 12; }
    else if (span <= 128) { cfg->NU = 2; cfg->WG = 6; }
    else if (span <= 192) { cfg->NU = 3; cfg->WG = 4; }
    else { cfg->NU = 4; cfg->WG = 3; }
    return true;
}

static int roundup4(int v) { return (v + 3) & ~3; }
static int mod4(int v) { return ((v % 4) + 4) % 4; }

void make_sliding_layout(const Problem &p, const SlidingConfig &cfg, bool backward, SlidingLayout *L) {
    L->K = cfg.K; L->r = cfg.r(); L->NU = cfg.NU; L->WG = cfg.WG; L->WTC = cfg.WTC(); L->SC = cfg.SC();
    L->banded = p.banded; L->smin_full = 0;
    L->n_wtiles = (p.W + L->WTC - 1) / L->WTC;
    L->n_chunks = p.banded ? (p.D + L->SC - 1) / L->SC : (p.W + L->WTC + 3 + L->SC - 1) / L->SC;
    // rows per band: as large as possible (k-1 warm-up rows per band are recomputed) while keeping >= 4 waves of CTAs
    const int64_t tiles = (int64_t)L->n_wtiles * L->n_chunks * p.B;
    // RB + K - 1 row steps must be a whole number of pair-sum-ring periods (K - 2)
    auto fit = [&](int target) { return target - (target + cfg.K - 1) % (cfg.K - 2); };
    int RB = fit(64);
    for (int target = 48; target >= 16 && ((p.H + RB - 1) / RB) * tiles < 8 * 148; target -= 16) RB = fit(target);
    L->RB = RB; L->RBH = RB + cfg.K - 1; L->NB = (p.H + RB - 1) / RB;
    L->seg_cam = cfg.seg_cam(); L->seg_proj = cfg.seg_proj(); L->seg_cs = cfg.seg_cs(); L->seg_ps = cfg.seg_ps();
    L->slot_floats = L->seg_cam + L->seg_proj + 2 * L->seg_cs + 2 * L->seg_ps;

    L->cam_lc = L->r;                                                 // column w_base - r sits at index w_base: 16-byte aligned
    L->cam_pitch = roundup4((L->n_wtiles - 1) * L->WTC - L->r + L->cam_lc + L->seg_cam);
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
    L->off_minmax = take(2 * bands * 2 * sizeof(uint32_t));            // [img][pair*band][max(v), max(-v)] ordered
    L->off_wta = take((size_t)p.pixels() * sizeof(unsigned long long));  // packed (best, s) keys; memset with minmax
    L->off_camP = take(bands * L->RBH * L->cam_pitch * sizeof(float));
    L->off_projP = take(bands * L->RBH * L->proj_pitch * sizeof(float));
    L->off_A = take(rows * L->cs_pitch * sizeof(float));
    L->off_ex2 = take(rows * L->cs_pitch * sizeof(float));
    L->off_Sp = take(rows * L->ps_pitch * sizeof(float));
    L->off_ey2 = take(rows * L->ps_pitch * sizeof(float));
    L->off_extra = off;
    (void)backward;
    L->total = off;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    band_minmax_kernel(Problem p, SlidingLayout L, const float *__restrict__ cam, const float *__restrict__ proj,
                       uint32_t *__restrict__ minmax) {
    const int nb = blockIdx.y, img = blockIdx.z / p.B, b = blockIdx.z % p.B;
    const float *plane = (img ? proj : cam) + (int64_t)b * p.H * p.W;
    const int y0 = max(0, nb * L.RB - L.r), y1 = min(p.H, nb * L.RB + L.RB + L.K - 1 - L.r);
    const int64_t n = (int64_t)(y1 - y0) * p.W;
    float vmax = -INFINITY, vmin = INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __ldg(plane + (int64_t)y0 * p.W + i);
        vmax = fmaxf(vmax, v);
        vmin = fminf(vmin, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    }
    if ((threadIdx.x & 31) == 0 && vmax >= vmin) {
        uint32_t *mm = minmax + ((size_t)(img * p.B + b) * L.NB + nb) * 2;
        atomicMax(mm, float_to_ordered(vmax));
        atomicMax(mm + 1, float_to_ordered(-vmin));
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

// one thread per element of the two band copies
__global__ void __launch_bounds__(256)
    band_copy_kernel(Problem p, SlidingLayout L, const float *__restrict__ cam, const float *__restrict__ proj,
                     const uint32_t *__restrict__ minmax, float *__restrict__ camP, float *__restrict__ projP) {
    const int img = blockIdx.z / p.B, b = blockIdx.z % p.B;
    const int pitch = img ? L.proj_pitch : L.cam_pitch, left = img ? L.proj_lp : L.cam_lc;
    const int t = blockIdx.y % L.RBH, nb = blockIdx.y / L.RBH;
    const int y = nb * L.RB - L.r + t;
    const float pv = band_pivot(minmax, p, L, img, b, nb);
    const float *row = (img ? proj : cam) + ((int64_t)b * p.H + y) * p.W;
    float *out = (img ? projP : camP) + (((int64_t)b * L.NB + nb) * L.RBH + t) * pitch;
    const bool yin = y >= 0 && y < p.H;
    for (int ci = blockIdx.x * blockDim.x + threadIdx.x; ci < pitch; ci += gridDim.x * blockDim.x) {
        const int x = ci - left;
        const float v = (yin && x >= 0 && x < p.W) ? __ldg(row + x) : 0.f;
        out[ci] = v - pv;
    }
}

// one thread per element of the statistics rows (camera: A, ex2; projector: Sp, ey2)
__global__ void __launch_bounds__(256)
    band_stats_kernel(Problem p, SlidingLayout L, const float *__restrict__ cam, const float *__restrict__ proj,
                      const uint32_t *__restrict__ minmax, float *__restrict__ A, float *__restrict__ ex2,
                      float *__restrict__ Sp, float *__restrict__ ey2) {
    const int img = blockIdx.z / p.B, b = blockIdx.z % p.B;
    const int pitch = img ? L.ps_pitch : L.cs_pitch, left = img ? L.ps_ld : 0;
    const int h = blockIdx.y, nb = h / L.RB;   // h in [0, NB*RB): rows past H are filled with neutral values
    const float pv = band_pivot(minmax, p, L, img, b, nb);
    const float *plane = (img ? proj : cam) + (int64_t)b * p.H * p.W;
    const int64_t orow = ((int64_t)b * L.NB * L.RB + h) * pitch;
    const int k = p.k, r = p.r;
    const float inv_n = 1.f / (float)(k * k);
    for (int ci = blockIdx.x * blockDim.x + threadIdx.x; ci < pitch; ci += gridDim.x * blockDim.x) {
        const int x = ci - left;
        float sum = 0.f, e2 = 1.f, mean = 0.f;
        if (h < p.H && x >= 0 && x < p.W) {
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j) sum += query_ij(plane, p.H, p.W, h + i - r, x + j - r) - pv;
            mean = sum * inv_n;
            e2 = 0.f;
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j) {
                    const float c = (query_ij(plane, p.H, p.W, h + i - r, x + j - r) - pv) - mean;
                    e2 = fmaf(c, c, e2);
                }
        }
        if (img) { Sp[orow + ci] = sum; ey2[orow + ci] = e2; }
        else { A[orow + ci] = mean; ex2[orow + ci] = e2; }
    }
}

int launch_sliding_prep(const Problem &p, const SlidingLayout &L, const float *cam, const float *proj, char *ws,
                        cudaStream_t stream) {
    uint32_t *minmax = (uint32_t *)(ws + L.off_minmax);
    // min/max accumulators and WTA keys are adjacent: one memset
    CUSTMA_CUDA_CHECK(cudaMemsetAsync(ws + L.off_minmax, 0, L.off_camP - L.off_minmax, stream));
    {
        const int64_t n = (int64_t)L.RBH * p.W;
        dim3 grid((unsigned)std::min<int64_t>((n + 1023) / 1024, 64), L.NB, 2 * p.B);
        band_minmax_kernel<<<grid, 256, 0, stream>>>(p, L, cam, proj, minmax);
        CUSTMA_LAUNCH_CHECK("band_minmax_kernel");
    }
    {
        dim3 grid((std::max(L.cam_pitch, L.proj_pitch) + 255) / 256, L.NB * L.RBH, 2 * p.B);
        band_copy_kernel<<<grid, 256, 0, stream>>>(p, L, cam, proj, minmax, (float *)(ws + L.off_camP),
                                                  (float *)(ws + L.off_projP));
        CUSTMA_LAUNCH_CHECK("band_copy_kernel");
    }
    {
        dim3 grid((std::max(L.cs_pitch, L.ps_pitch) + 255) / 256, L.NB * L.RB, 2 * p.B);
        band_stats_kernel<<<grid, 256, 0, stream>>>(p, L, cam, proj, minmax, (float *)(ws + L.off_A),
                                                   (float *)(ws + L.off_ex2), (float *)(ws + L.off_Sp),
                                                   (float *)(ws + L.off_ey2));
        CUSTMA_LAUNCH_CHECK("band_stats_kernel");
    }
    return CUSTMA_OK;
}

}  // namespace custma
