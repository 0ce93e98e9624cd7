// Direct two-pass ZNCC kernels: the always-correct path (any kernel size, full or banded), in the reference's own
// arithmetic order, so the forward is bit-exact with the reference extension.  It is also the fallback the
// sliding-window kernels hand ill-conditioned tiles to.
//
// What it restates (reference = custma/src/stereo_matching_kernel.cu):
//   window_stats_kernel      the per-image halves of :39-70 (mean over k*k incl. zero padding, centred 2nd moment),
//                            hoisted out of the per-cell loop: the reference recomputes them in every cell.
//   direct_forward_kernel    :17-72 (exy chain in i-major/j-minor order, cost at :71) + WTA (examples/verify.py:72)
//   direct_patch_grad_kernel :75-152 without atomics: one warp owns one pixel's k*k patch gradient
//   gather_patch_grad_kernel :155-179 as a gather (each image pixel sums the k*k patch entries that target it)
#include "common.cuh"

namespace custma {

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) window_stats_kernel(const float *__restrict__ img, int64_t pixels, int H,
                                                           int W, int k, float *__restrict__ mean,
                                                           float *__restrict__ e2) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= pixels) return;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const float *plane = img + (pix / ((int64_t)H * W)) * (int64_t)H * W;
    const int r = k / 2;
    float m = 0.f;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) m += query_ij(plane, H, W, h + i - r, w + j - r);
    m /= (float)(k * k);
    float s2 = 0.f;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            const float c = query_ij(plane, H, W, h + i - r, w + j - r) - m;
            s2 = fmaf(c, c, s2);
        }
    mean[pix] = m;
    e2[pix] = s2;
}

int launch_window_stats(const float *img, int B, int H, int W, int k, float *mean, float *e2, cudaStream_t stream) {
    const int64_t pixels = (int64_t)B * H * W;
    const int threads = 256;
    const int64_t blocks = (pixels + threads - 1) / threads;
    window_stats_kernel<<<(unsigned)blocks, threads, 0, stream>>>(img, pixels, H, W, k, mean, e2);
    CUSTMA_LAUNCH_CHECK("window_stats_kernel");
    return CUSTMA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// One warp per camera pixel (b,h,w); lanes stride over the last axis.  The centred camera patch is shared by all
// cells of the pixel and staged once per warp in shared memory.
constexpr int kDirectWarps = 8;

__device__ __forceinline__ float cell_exy(const float *__restrict__ camc /* smem [k*k] */,
                                          const float *__restrict__ proj_plane, int H, int W, int k, int r, int h,
                                          int d, float pm) {
    float exy = 0.f;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            const float p = query_ij(proj_plane, H, W, h + i - r, d + j - r) - pm;
            exy = fmaf(camc[i * k + j], p, exy);
        }
    return exy;
}

__global__ void __launch_bounds__(kDirectWarps * 32)
    direct_forward_kernel(Problem p, const float *__restrict__ cam, const float *__restrict__ proj,
                          const float *__restrict__ cmean, const float *__restrict__ cex2,
                          const float *__restrict__ pmean, const float *__restrict__ pey2, float *__restrict__ cost,
                          float *__restrict__ best, int32_t *__restrict__ index) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t pix = (int64_t)blockIdx.x * kDirectWarps + warp;
    if (pix >= p.pixels()) return;
    const int w = (int)(pix % p.W);
    const int h = (int)((pix / p.W) % p.H);
    const int64_t plane_off = (pix / ((int64_t)p.H * p.W)) * (int64_t)p.H * p.W;
    const float *cam_plane = cam + plane_off, *proj_plane = proj + plane_off;
    const float *pm_row = pmean + plane_off + (int64_t)h * p.W, *ey2_row = pey2 + plane_off + (int64_t)h * p.W;
    float *camc = smem + warp * p.k * p.k;
    const float cm = cmean[pix], ex2 = cex2[pix];
    for (int t = lane; t < p.k * p.k; t += 32)
        camc[t] = query_ij(cam_plane, p.H, p.W, h + t / p.k - p.r, w + t % p.k - p.r) - cm;
    __syncwarp();

    float bval = -INFINITY;
    int bd = INT32_MAX;  // winning projector column; ties -> lowest column
    for (int c = lane; c < p.C; c += 32) {
        const int d = p.banded ? w - c : c;
        float v = kInvalid;
        if (d >= 0) {
            const float exy = cell_exy(camc, proj_plane, p.H, p.W, p.k, p.r, h, d, pm_row[d]);
            v = (exy + kEps) / sqrtf(fmaf(ex2, ey2_row[d], kEps));  // reference kernel.cu:71
            if (v > bval || (v == bval && d < bd)) { bval = v; bd = d; }
        }
        if (cost) cost[pix * p.C + c] = v;
    }
    if (best) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bval, o);
            const int od = __shfl_xor_sync(0xffffffffu, bd, o);
            if (ov > bval || (ov == bval && od < bd)) { bval = ov; bd = od; }
        }
        if (lane == 0) {
            best[pix] = bval;
            index[pix] = p.banded ? w - bd : bd;
        }
    }
}

int launch_direct_forward(const Problem &p, const float *cam, const float *proj, const float *cmean,
                          const float *cex2, const float *pmean, const float *pey2, float *cost, float *best,
                          int32_t *index, cudaStream_t stream) {
    const int64_t blocks = (p.pixels() + kDirectWarps - 1) / kDirectWarps;
    const size_t smem = (size_t)kDirectWarps * p.k * p.k * sizeof(float);
    direct_forward_kernel<<<(unsigned)blocks, kDirectWarps * 32, smem, stream>>>(p, cam, proj, cmean, cex2, pmean,
                                                                                pey2, cost, best, index);
    CUSTMA_LAUNCH_CHECK("direct_forward_kernel");
    return CUSTMA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Backward, step 1: patch gradient of one camera pixel, one warp per pixel, no atomics.
//   patch_grad[i,j] = sum_d a_d * (proj[h+i-r, d+j-r] - pm_d) - cam_c[i,j] * sum_d b_d
//   a_d = g_d / den_d,  b_d = g_d * ey2_d * (exy_d + eps) / den_d^3,  den_d = sqrt(ex2 * ey2_d + eps)   (kernel.cu:135-148)
// The projector factor is centred BEFORE it is multiplied (as the reference does at :144-145): summing a_d * proj
// and a_d * pm_d separately cancels catastrophically on low-contrast images.
// Dynamic shared memory per warp: k*k (centred camera patch) + 2*C (a_d and pm_d of this pixel).
constexpr int kBwdWarps = 4;

// SWAP = false: gradient with respect to the camera image (what the reference computes).  SWAP = true: with respect to
// the PROJECTOR image - ZNCC is symmetric in its two patches, so the same kernel runs with the roles exchanged: the
// pixel is a projector pixel (h, d), the loop runs over the camera columns w that meet it (banded: w = d + s), `cam` /
// `cmean` / `cex2` are the projector's arrays and vice versa.  The reference has no such kernel
// (custma/stereo_matching_wrapper.py:33 returns None for the projector): SURVEY.md 8f #2.
template <bool SWAP>
__global__ void __launch_bounds__(kBwdWarps * 32)
    direct_patch_grad_kernel(Problem p, const float *__restrict__ grad, const float *__restrict__ cam,
                             const float *__restrict__ proj, const float *__restrict__ cmean,
                             const float *__restrict__ cex2, const float *__restrict__ pmean,
                             const float *__restrict__ pey2, float *__restrict__ patch_grad) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t pix = (int64_t)blockIdx.x * kBwdWarps + warp;
    if (pix >= p.pixels()) return;
    const int kk = p.k * p.k;
    const int w = (int)(pix % p.W);     // SWAP: the projector column d of this pixel
    const int h = (int)((pix / p.W) % p.H);
    const int64_t bb = pix / ((int64_t)p.H * p.W);
    const int64_t plane_off = bb * (int64_t)p.H * p.W;
    const float *cam_plane = cam + plane_off, *proj_plane = proj + plane_off;
    const float *pm_row = pmean + plane_off + (int64_t)h * p.W, *ey2_row = pey2 + plane_off + (int64_t)h * p.W;
    float *camc = smem + (size_t)warp * (kk + 2 * p.C);
    float *a_s = camc + kk;
    float *pm_s = a_s + p.C;
    const float cm = cmean[pix], ex2 = cex2[pix];
    for (int t = lane; t < kk; t += 32)
        camc[t] = query_ij(cam_plane, p.H, p.W, h + t / p.k - p.r, w + t % p.k - p.r) - cm;
    __syncwarp();
    const bool grow = h >= p.g0 && h < p.g1;
    const int64_t grow_base = (bb * p.grows() + (h - p.g0)) * p.W;

    float bsum = 0.f;
    for (int c = lane; c < p.C; c += 32) {
        // the other image's column this cell pairs the pixel with
        const int d = SWAP ? (p.banded ? w + c : c) : (p.banded ? w - c : c);
        float a = 0.f, pm = 0.f;
        if (d >= 0 && d < p.W) {
            pm = pm_row[d];
            const float ey2 = ey2_row[d];
            const float exy = cell_exy(camc, proj_plane, p.H, p.W, p.k, p.r, h, d, pm);
            const float den = sqrtf(fmaf(ex2, ey2, kEps));
            // cell [h, camera column, last-axis index]: camera column = w (or d when swapped); banded index c, full index = projector column
            const int64_t gi = SWAP ? (grow_base + d) * p.C + (p.banded ? c : w) : (grow_base + w) * p.C + c;
            const float g = grow ? grad[gi] : 0.f;
            a = g / den;
            bsum += g * ey2 * (exy + kEps) / (den * den * den);
        }
        a_s[c] = a;
        pm_s[c] = pm;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
    __syncwarp();
    for (int t = 0; t < kk; ++t) {
        const int y = h + t / p.k - p.r, xo = t % p.k - p.r;
        float acc = 0.f;
        if (y >= 0 && y < p.H) {
            const float *prow = proj_plane + (int64_t)y * p.W;
            for (int c = lane; c < p.C; c += 32) {
                const int d = SWAP ? (p.banded ? w + c : c) : (p.banded ? w - c : c);
                const int x = d + xo;
                if (d >= 0 && d < p.W) acc = fmaf(a_s[c], ((x >= 0 && x < p.W) ? __ldg(prow + x) : 0.f) - pm_s[c], acc);
            }
        } else {
            for (int c = lane; c < p.C; c += 32) acc = fmaf(a_s[c], -pm_s[c], acc);  // zero-padded row: proj = 0
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) patch_grad[pix * kk + t] = acc - bsum * camc[t];
    }
}

// Backward, step 2: camera_grad[y,x] = sum_{i,j} patch_grad[y-i+r, x-j+r, i, j] over in-image sources
// (the transpose of the reference's scatter at kernel.cu:172-178).
__global__ void __launch_bounds__(256)
    gather_patch_grad_kernel(Problem p, const float *__restrict__ patch_grad, float *__restrict__ camera_grad) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p.pixels()) return;
    const int x = (int)(pix % p.W);
    const int y = (int)((pix / p.W) % p.H);
    const int64_t plane = (pix / ((int64_t)p.H * p.W)) * (int64_t)p.H * p.W;
    const int kk = p.k * p.k;
    float acc = 0.f;
    for (int i = 0; i < p.k; ++i) {
        const int h = y - i + p.r;
        if (h < 0 || h >= p.H) continue;
        for (int j = 0; j < p.k; ++j) {
            const int w = x - j + p.r;
            if (w < 0 || w >= p.W) continue;
            acc += patch_grad[(plane + (int64_t)h * p.W + w) * kk + i * p.k + j];
        }
    }
    camera_grad[pix] = acc;
}

static int launch_direct_backward_impl(const Problem &p, bool swap, const float *grad, const float *cam, const float *proj,
                                       const float *cmean, const float *cex2, const float *pmean, const float *pey2,
                                       float *patch_grad, float *image_grad, cudaStream_t stream) {
    const size_t smem = (size_t)kBwdWarps * (p.k * p.k + 2 * p.C) * sizeof(float);
    if (smem > 200 * 1024)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "direct backward: last axis %d too long for shared memory", p.C);
    auto kern = swap ? direct_patch_grad_kernel<true> : direct_patch_grad_kernel<false>;
    if (smem > 48 * 1024)
        CUSTMA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks = (p.pixels() + kBwdWarps - 1) / kBwdWarps;
    kern<<<(unsigned)blocks, kBwdWarps * 32, smem, stream>>>(p, grad, cam, proj, cmean, cex2, pmean, pey2, patch_grad);
    CUSTMA_LAUNCH_CHECK("direct_patch_grad_kernel");
    const int64_t gblocks = (p.pixels() + 255) / 256;
    gather_patch_grad_kernel<<<(unsigned)gblocks, 256, 0, stream>>>(p, patch_grad, image_grad);
    CUSTMA_LAUNCH_CHECK("gather_patch_grad_kernel");
    return CUSTMA_OK;
}

int launch_direct_backward(const Problem &p, const float *grad, const float *cam, const float *proj,
                           const float *cmean, const float *cex2, const float *pmean, const float *pey2,
                           float *patch_grad, float *camera_grad, cudaStream_t stream) {
    return launch_direct_backward_impl(p, false, grad, cam, proj, cmean, cex2, pmean, pey2, patch_grad, camera_grad, stream);
}

// gradient with respect to the projector image: the same two kernels with the two images' roles exchanged
int launch_direct_backward_projector(const Problem &p, const float *grad, const float *cam, const float *proj,
                                     const float *cmean, const float *cex2, const float *pmean, const float *pey2,
                                     float *patch_grad, float *projector_grad, cudaStream_t stream) {
    return launch_direct_backward_impl(p, true, grad, proj, cam, pmean, pey2, cmean, cex2, patch_grad, projector_grad, stream);
}

}  // namespace custma
