// Fallback of the sliding-window path: the tiles that tile_flags_kernel (sliding_prep.cu) judged ill-conditioned for
// the O(1) window sums are computed here cell by cell with centred two-pass sums, i.e. the arithmetic of the
// reference kernels (custma/src/stereo_matching_kernel.cu:39-71 forward, :96-151 backward).  One warp per camera
// pixel; every cell recomputes its projector window mean and moments, as the reference does.  Slow by design: it
// only runs where the fast path would lose accuracy (flat or very low-contrast regions, borders of images with a
// large DC level), and it returns at once for tiles without a flagged chunk.
#include <algorithm>

#include "sliding_common.cuh"

namespace custma {

constexpr int kFbWarps = 8;
constexpr int kFbRows = kFallbackRows;
static int fb_grid() { return device_sm_count() * 8; }   // CTAs looping over the work list (they return at once when it is empty)

// chunk of the sliding tiling that owns cell (w, c) of column tile w_base
__device__ __forceinline__ int cell_chunk(const Problem &p, const SlidingLayout &L, int w_base, int w, int c) {
    const int s = p.banded ? c : w - c;
    return (s - chunk_s_base(L, p.W, w_base, 0)) / L.SC;
}

// centred camera patch of pixel (h,w) into shared memory, returns (via all lanes) its second moment
__device__ __forceinline__ float warp_camera_patch(const Problem &p, const float *cam_plane, int h, int w, float *camc) {
    const int lane = threadIdx.x & 31, kk = p.k * p.k;
    float sum = 0.f;
    for (int t = lane; t < kk; t += 32) {
        const float v = query_ij(cam_plane, p.H, p.W, h + t / p.k - p.r, w + t % p.k - p.r);
        camc[t] = v;
        sum += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float cm = sum / (float)kk;
    float e2 = 0.f;
    for (int t = lane; t < kk; t += 32) {
        const float c = camc[t] - cm;
        camc[t] = c;
        e2 = fmaf(c, c, e2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e2 += __shfl_xor_sync(0xffffffffu, e2, o);
    __syncwarp();
    return e2;
}

// Projector window statistics (mean over k*k including the zero padding, centred second moment: reference :40-70) of
// every pixel of the bands that hold a flagged tile.  They do not depend on the camera column, so computing them once
// per pixel instead of once per cell makes the fallback three times cheaper.
__global__ void __launch_bounds__(256)
    fallback_proj_stats_kernel(const Problem p, const SlidingLayout L, const float *__restrict__ proj,
                               const uint32_t *__restrict__ bandany, float *__restrict__ pm_out,
                               float *__restrict__ ey2_out) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    const int nb = blockIdx.y, b = blockIdx.z;
    if (!bandany[b * L.NB + nb]) return;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= p.W) return;
    const float *plane = proj + (int64_t)b * p.H * p.W;
    for (int h = nb * L.RB; h < min(p.H, (nb + 1) * L.RB); ++h) {
        float pm = 0.f;
        for (int i = 0; i < p.k; ++i)
            for (int j = 0; j < p.k; ++j) pm += query_ij(plane, p.H, p.W, h + i - p.r, d + j - p.r);
        pm /= (float)(p.k * p.k);
        float ey2 = 0.f;
        for (int i = 0; i < p.k; ++i)
            for (int j = 0; j < p.k; ++j) {
                const float q = query_ij(plane, p.H, p.W, h + i - p.r, d + j - p.r) - pm;
                ey2 = fmaf(q, q, ey2);
            }
        const int64_t o = ((int64_t)b * p.H + h) * p.W + d;
        pm_out[o] = pm;
        ey2_out[o] = ey2;
    }
}

// centred correlation of the camera patch with the projector window of column d (reference :56-70)
__device__ __forceinline__ float cell_exy(const Problem &p, const float *proj_plane, const float *camc, int h, int d, float pm) {
    float exy = 0.f;
    for (int i = 0; i < p.k; ++i)
        for (int j = 0; j < p.k; ++j)
            exy = fmaf(camc[i * p.k + j], query_ij(proj_plane, p.H, p.W, h + i - p.r, d + j - p.r) - pm, exy);
    return exy;
}

__global__ void __launch_bounds__(kFbWarps * 32)
    fallback_forward_kernel(const Problem p, const SlidingLayout L, const float *__restrict__ cam,
                            const float *__restrict__ proj, const uint8_t *__restrict__ flags,
                            const uint32_t *__restrict__ fb_count, const uint32_t *__restrict__ fb_list,
                            const float *__restrict__ fb_pm, const float *__restrict__ fb_ey2,
                            float *__restrict__ cost, unsigned long long *__restrict__ keys, const uint32_t tc_threshold,
                            const HeadOut head) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    extern __shared__ float smem[];
    const uint32_t n_items = *fb_count;
    if (n_items > tc_threshold) return;   // so much is flagged that the tensor-core kernel computes the whole call
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const uint32_t code = fb_list[item];
    const int64_t t3 = code / L.fb_groups;
    const int rg = (int)(code % L.fb_groups);
    const int wt = (int)(t3 % L.n_wtiles), nb = (int)((t3 / L.n_wtiles) % L.NB), b = (int)(t3 / ((int64_t)L.n_wtiles * L.NB));
    const uint8_t *fl = flags + t3 * L.n_chunks;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h0 = nb * L.RB + rg * kFbRows, w_base = wt * L.WTC;
    const int rows = min(min(kFbRows, L.RB - rg * kFbRows), p.H - h0), cols = min(L.WTC, p.W - w_base);
    const float *cam_plane = cam + (int64_t)b * p.H * p.W, *proj_plane = proj + (int64_t)b * p.H * p.W;
    float *camc = smem + warp * (p.k * p.k + (head.part ? p.C : 0));
    float *vst = camc + p.k * p.k;     // head: the costs of this pixel's flagged cells (-inf where there is none)
    for (int pi = warp; pi < rows * cols; pi += kFbWarps) {
        const int h = h0 + pi / cols, w = w_base + pi % cols;
        const int64_t pix = ((int64_t)b * p.H + h) * p.W + w;
        __syncwarp();
        const float ex2 = warp_camera_patch(p, cam_plane, h, w, camc);
        float bv = -INFINITY;
        int bs = 0;
        for (int c = lane; c < p.C; c += 32) {
            if (!fl[cell_chunk(p, L, w_base, w, c)]) {
                if (head.part) vst[c] = -INFINITY;
                continue;
            }
            const int d = p.banded ? w - c : c;
            float v = kInvalid;
            if (d >= 0) {
                const int64_t o = ((int64_t)b * p.H + h) * p.W + d;
                const float exy = cell_exy(p, proj_plane, camc, h, d, fb_pm[o]);
                v = (exy + kEps) / sqrtf(fmaf(ex2, fb_ey2[o], kEps));   // reference :71
                const int s = w - d;
                if (v > bv || (v == bv && s > bs)) { bv = v; bs = s; }
            }
            if (cost) cost[pix * p.C + c] = v;
            if (head.part) vst[c] = d >= 0 ? v : -INFINITY;
        }
        if (keys) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int os = __shfl_xor_sync(0xffffffffu, bs, o);
                if (ov > bv || (ov == bv && os > bs)) { bv = ov; bs = os; }
            }
            if (lane == 0 && bv > -INFINITY)
                atomicMax(keys + pix, ((unsigned long long)float_to_ordered(bv) << 32) | (uint32_t)(bs + p.W));
        }
        if (head.part) {
            // softmax partial of ALL flagged cells of this pixel relative to their maximum bv (every lane holds it after
            // the reduction above); it goes to the first slot of the first flagged chunk, the other slots of the flagged
            // chunks are marked empty - the sliding kernel writes the slots of the chunks it keeps
            __syncwarp();
            float z = 0.f, n = 0.f;
            for (int c = lane; c < p.C; c += 32) {
                const float v = vst[c];
                if (v > -INFINITY) {
                    const float ex = exp2_fast(head.beta_log2e * (v - bv));
                    z += ex;
                    n = fmaf(ex, (float)(p.banded ? c : w - c), n);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                z += __shfl_xor_sync(0xffffffffu, z, o);
                n += __shfl_xor_sync(0xffffffffu, n, o);
            }
            if (lane == 0) {
                bool first = true;
                for (int ch = 0; ch < L.n_chunks; ++ch) {
                    if (!fl[ch]) continue;
                    for (int un = 0; un < L.NU; ++un) {
                        head.part[(int64_t)(ch * L.NU + un) * p.pixels() + pix] =
                            first ? make_float4(bv, z, n, 0.f) : make_float4(-INFINITY, 0.f, 0.f, 0.f);
                        first = false;
                    }
                }
            }
        }
    }
    }   // work items
}

// patch gradient (k*k values per pixel) of the flagged cells; layout and formula of direct_patch_grad_kernel
__global__ void __launch_bounds__(kFbWarps * 32)
    fallback_patch_grad_kernel(const Problem p, const SlidingLayout L, const float *__restrict__ grad, const HeadGrad hg,
                               const float *__restrict__ cam, const float *__restrict__ proj,
                               const uint8_t *__restrict__ flags, const uint32_t *__restrict__ fb_count,
                               const uint32_t *__restrict__ fb_list, const float *__restrict__ fb_pm,
                               const float *__restrict__ fb_ey2, float *__restrict__ patch_grad, const uint32_t tc_threshold) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    extern __shared__ float smem[];
    const uint32_t n_items = *fb_count;
    if (n_items > tc_threshold) return;   // the tensor-core kernel computes the whole call
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const uint32_t code = fb_list[item];
    const int64_t t3 = code / L.fb_groups;
    const int rg = (int)(code % L.fb_groups);
    const int wt = (int)(t3 % L.n_wtiles), nb = (int)((t3 / L.n_wtiles) % L.NB), b = (int)(t3 / ((int64_t)L.n_wtiles * L.NB));
    const uint8_t *fl = flags + t3 * L.n_chunks;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, kk = p.k * p.k;
    const int h0 = nb * L.RB + rg * kFbRows, w_base = wt * L.WTC;
    const int rows = min(min(kFbRows, L.RB - rg * kFbRows), p.H - h0), cols = min(L.WTC, p.W - w_base);
    const float *cam_plane = cam + (int64_t)b * p.H * p.W, *proj_plane = proj + (int64_t)b * p.H * p.W;
    float *camc = smem + (size_t)warp * (kk + 2 * p.C);
    float *a_s = camc + kk, *pm_s = a_s + p.C;
    for (int pi = warp; pi < rows * cols; pi += kFbWarps) {
        const int h = h0 + pi / cols, w = w_base + pi % cols;
        const int64_t pix = ((int64_t)b * p.H + h) * p.W + w;
        __syncwarp();
        const float ex2 = warp_camera_patch(p, cam_plane, h, w, camc);
        float bsum = 0.f;
        for (int c = lane; c < p.C; c += 32) {
            const int d = p.banded ? w - c : c;
            float a = 0.f, pm = 0.f;
            if (d >= 0 && fl[cell_chunk(p, L, w_base, w, c)]) {
                const int64_t o = ((int64_t)b * p.H + h) * p.W + d;
                pm = fb_pm[o];
                const float ey2 = fb_ey2[o];
                const float exy = cell_exy(p, proj_plane, camc, h, d, pm);
                const float den = sqrtf(fmaf(ex2, ey2, kEps));
                float g;
                if (hg.state) {   // fused head: the cell's upstream gradient from the per-pixel softmax state
                    const float4 st = hg.state[pix];
                    const float cv = (exy + kEps) / den;
                    g = exp2_fast(fmaf(cv, hg.beta_log2e, -st.x)) * fmaf((float)(p.banded ? c : w - c), st.y, -st.z);
                } else {
                    g = (h >= p.g0 && h < p.g1) ? grad[(((int64_t)b * p.grows() + (h - p.g0)) * p.W + w) * p.C + c] : 0.f;
                }
                a = g / den;                                          // reference :135,:145
                bsum += g * ey2 * (exy + kEps) / (den * den * den);   // reference :147
            }
            a_s[c] = a;
            pm_s[c] = pm;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
        __syncwarp();
        for (int t = 0; t < kk; ++t) {
            const int y = h + t / p.k - p.r, xo = t % p.k - p.r;
            const bool yin = y >= 0 && y < p.H;
            const float *prow = proj_plane + (int64_t)y * p.W;
            float acc = 0.f;
            for (int c = lane; c < p.C; c += 32) {
                const float a = a_s[c];
                if (a != 0.f) {
                    const int x = (p.banded ? w - c : c) + xo;
                    acc = fmaf(a, ((yin && x >= 0 && x < p.W) ? __ldg(prow + x) : 0.f) - pm_s[c], acc);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) patch_grad[pix * kk + t] = acc - bsum * camc[t];
        }
    }
    }   // work items
}

static int launch_fallback_proj_stats(const Problem &p, const SlidingLayout &L, const float *proj, const char *ws,
                                      cudaStream_t stream) {
    dim3 grid((p.W + 255) / 256, L.NB, p.B);
    CUSTMA_CUDA_CHECK(launch_chained(fallback_proj_stats_kernel, grid, dim3(256), 0, stream, p, L, proj,
                                     (const uint32_t *)(ws + L.off_bandany), (float *)(const_cast<char *>(ws) + L.off_fb_pm),
                                     (float *)(const_cast<char *>(ws) + L.off_fb_ey2)));
    CUSTMA_LAUNCH_CHECK("fallback_proj_stats_kernel");
    return CUSTMA_OK;
}

int launch_fallback_forward(const Problem &p, const SlidingLayout &L, const float *cam, const float *proj,
                            const char *ws, float *cost, unsigned long long *keys, const HeadOut &head,
                            uint32_t tc_threshold, cudaStream_t stream) {
    int rc = launch_fallback_proj_stats(p, L, proj, ws, stream);
    if (rc) return rc;
    const size_t smem = (size_t)kFbWarps * (p.k * p.k + (head.part ? p.C : 0)) * sizeof(float);
    if (smem > 200 * 1024)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "fallback forward: last axis %d too long for shared memory", p.C);
    if (smem > 48 * 1024)
        CUSTMA_CUDA_CHECK(cudaFuncSetAttribute(fallback_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUSTMA_CUDA_CHECK(launch_chained(fallback_forward_kernel, dim3(fb_grid()), dim3(kFbWarps * 32), smem, stream, p, L, cam, proj,
                                     (const uint8_t *)(ws + L.off_flags), (const uint32_t *)(ws + L.off_fb_count),
                                     (const uint32_t *)(ws + L.off_fb_list), (const float *)(ws + L.off_fb_pm),
                                     (const float *)(ws + L.off_fb_ey2), cost, keys, tc_threshold, head));
    CUSTMA_LAUNCH_CHECK("fallback_forward_kernel");
    return CUSTMA_OK;
}

size_t fallback_backward_smem(const Problem &p) { return (size_t)kFbWarps * (p.k * p.k + 2 * p.C) * sizeof(float); }

int launch_fallback_patch_grad(const Problem &p, const SlidingLayout &L, const float *grad, const HeadGrad &hg,
                               const float *cam, const float *proj, const char *ws, float *patch_grad,
                               uint32_t tc_threshold, cudaStream_t stream) {
    const size_t smem = fallback_backward_smem(p);
    if (smem > 200 * 1024)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "fallback backward: last axis %d too long for shared memory", p.C);
    CUSTMA_CUDA_CHECK(cudaFuncSetAttribute(fallback_patch_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)std::max<size_t>(smem, 48 * 1024)));
    int rc = launch_fallback_proj_stats(p, L, proj, ws, stream);
    if (rc) return rc;
    CUSTMA_CUDA_CHECK(launch_chained(fallback_patch_grad_kernel, dim3(fb_grid()), dim3(kFbWarps * 32), smem, stream, p, L, grad, hg,
                                     cam, proj, (const uint8_t *)(ws + L.off_flags), (const uint32_t *)(ws + L.off_fb_count),
                                     (const uint32_t *)(ws + L.off_fb_list), (const float *)(ws + L.off_fb_pm),
                                     (const float *)(ws + L.off_fb_ey2), patch_grad, tc_threshold));
    CUSTMA_LAUNCH_CHECK("fallback_patch_grad_kernel");
    return CUSTMA_OK;
}

}  // namespace custma
