/* CPU restatement of the three reference CUDA kernels, cell by cell.  TEST INFRASTRUCTURE ONLY (see oracle/README).
 *
 * Follows /root/reference/custma/src/stereo_matching_kernel.cu:
 *   query_ij                       :6-12    bounds-checked read, 0 outside the image
 *   forward_cost_volume_kernel     :17-72   two passes over the k x k window, cost = (exy+eps)/sqrtf(ex2*ey2+eps)
 *   get_patches_grad_kernel        :75-152  per-cell patch gradient (atomicAdd at :149 becomes a plain ordered sum)
 *   patches_grad_to_image_kernel   :155-179 overlap-add, out-of-image targets dropped (:177)
 * nvcc contracts `acc += a*b` and `a*b + c` into FFMA (default -fmad=true; setup.py:30-38 passes no flags), so the
 * same places use fmaf() here; division and sqrtf are IEEE on both sides.  The summation ORDER of the reference's
 * atomics is undefined (run-to-run differences were measured on B200, tests/golden/manifest.json), here it is fixed:
 * ascending d, then ascending (h,w).
 *
 * The banded entry points apply the definition band[h,w,s] = full[h,w,w-s] with the same
 * per-cell arithmetic; cells with w-s < 0 hold `invalid` and take no part in the gradient.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off; contraction is explicit via fmaf)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EPSILON 1e-8f /* kernel.cu:4 */

static inline float query_ij(const float *img, int H, int W, int i, int j) {
    return (i < 0 || i >= H || j < 0 || j >= W) ? 0.f : img[(size_t)i * W + j];
}

/* the per-cell body shared by forward and backward (kernel.cu:39-70 and :96-128 are the same code) */
static inline void cell_moments(const float *cam, const float *proj, int H, int W, int k, int h, int w, int d,
                                float *cam_mean, float *proj_mean, float *exy, float *ex2, float *ey2) {
    const int r = k / 2;
    float cm = 0.f, pm = 0.f;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            cm += query_ij(cam, H, W, h + i - r, w + j - r);
            pm += query_ij(proj, H, W, h + i - r, d + j - r);
        }
    cm /= (float)(k * k);
    pm /= (float)(k * k);
    float sxy = 0.f, sx2 = 0.f, sy2 = 0.f;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            const float c = query_ij(cam, H, W, h + i - r, w + j - r) - cm;
            const float p = query_ij(proj, H, W, h + i - r, d + j - r) - pm;
            sxy = fmaf(c, p, sxy);
            sx2 = fmaf(c, c, sx2);
            sy2 = fmaf(p, p, sy2);
        }
    *cam_mean = cm; *proj_mean = pm; *exy = sxy; *ex2 = sx2; *ey2 = sy2;
}

static inline float cell_cost(const float *cam, const float *proj, int H, int W, int k, int h, int w, int d) {
    float cm, pm, exy, ex2, ey2;
    cell_moments(cam, proj, H, W, k, h, w, d, &cm, &pm, &exy, &ex2, &ey2);
    return (exy + EPSILON) / sqrtf(fmaf(ex2, ey2, EPSILON)); /* kernel.cu:71 */
}

/* out: [H, W, W] */
void ref_forward_full(const float *cam, const float *proj, int H, int W, int k, float *out) {
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w)
            for (int d = 0; d < W; ++d)
                out[((size_t)h * W + w) * W + d] = cell_cost(cam, proj, H, W, k, h, w, d);
}

/* out: [H, W, D], band[h,w,s] = full[h,w,w-s] */
void ref_forward_banded(const float *cam, const float *proj, int H, int W, int D, int k, float invalid, float *out) {
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w)
            for (int s = 0; s < D; ++s) {
                const int d = w - s;
                out[((size_t)h * W + w) * D + s] = d >= 0 ? cell_cost(cam, proj, H, W, k, h, w, d) : invalid;
            }
}

/* accumulate one cell's contribution into the k*k patch gradient of pixel (h,w): kernel.cu:130-151 */
static inline void cell_patch_grad(const float *cam, const float *proj, int H, int W, int k, int h, int w, int d,
                                   float g, float *pg /* [k*k] */) {
    const int r = k / 2;
    float cm, pm, exy, ex2, ey2;
    cell_moments(cam, proj, H, W, k, h, w, d, &cm, &pm, &exy, &ex2, &ey2);
    const float den = sqrtf(fmaf(ex2, ey2, EPSILON));
    const float deno = 1.f / den, deno3 = 1.f / powf(den, 3.f); /* :135 */
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            const float c = query_ij(cam, H, W, h + i - r, w + j - r) - cm;
            const float p = query_ij(proj, H, W, h + i - r, d + j - r) - pm;
            const float exy_factor = p * deno;                            /* :145 */
            const float ex2_factor = -(ey2 * c * (exy + EPSILON)) * deno3; /* :147 */
            pg[i * k + j] += g * (exy_factor + ex2_factor);               /* :148-149 */
        }
}

static void scatter_patches(const float *patch, int H, int W, int k, float *cam_grad) {
    const int r = k / 2; /* kernel.cu:155-179, serial and ordered */
    memset(cam_grad, 0, sizeof(float) * (size_t)H * W);
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w)
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j) {
                    const int y = h + i - r, x = w + j - r;
                    if (y < 0 || y >= H || x < 0 || x >= W) continue; /* :177 */
                    cam_grad[(size_t)y * W + x] += patch[(((size_t)h * W + w) * k + i) * k + j];
                }
}

/* g: [H, W, W] -> cam_grad [H, W] */
void ref_backward_full(const float *g, const float *cam, const float *proj, int H, int W, int k, float *cam_grad) {
    float *patch = (float *)calloc((size_t)H * W * k * k, sizeof(float));
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w)
            for (int d = 0; d < W; ++d)
                cell_patch_grad(cam, proj, H, W, k, h, w, d, g[((size_t)h * W + w) * W + d],
                                patch + ((size_t)h * W + w) * k * k);
    scatter_patches(patch, H, W, k, cam_grad);
    free(patch);
}

/* g: [H, W, D] on the band -> cam_grad [H, W]; invalid cells contribute nothing */
void ref_backward_banded(const float *g, const float *cam, const float *proj, int H, int W, int D, int k,
                         float *cam_grad) {
    float *patch = (float *)calloc((size_t)H * W * k * k, sizeof(float));
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w)
            for (int s = 0; s < D; ++s) {
                const int d = w - s;
                if (d < 0) continue;
                cell_patch_grad(cam, proj, H, W, k, h, w, d, g[((size_t)h * W + w) * D + s],
                                patch + ((size_t)h * W + w) * k * k);
            }
    scatter_patches(patch, H, W, k, cam_grad);
    free(patch);
}

/* WTA over the last axis of a reference-shaped volume: first maximal index (torch.max, examples/verify.py:72) */
void ref_wta_full(const float *vol, int H, int W, float *best, int32_t *corr) {
    for (size_t p = 0; p < (size_t)H * W; ++p) {
        const float *row = vol + p * W;
        float b = row[0]; int32_t a = 0;
        for (int d = 1; d < W; ++d) if (row[d] > b) { b = row[d]; a = d; }
        best[p] = b; corr[p] = a;
    }
}

/* WTA on the band: lowest projector column wins ties == largest disparity; invalid cells skipped */
void ref_wta_banded(const float *vol, int H, int W, int D, float *best, int32_t *disp) {
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
            const float *row = vol + ((size_t)h * W + w) * D;
            const int smax = w < D - 1 ? w : D - 1;
            float b = row[smax]; int32_t a = smax;
            for (int s = smax - 1; s >= 0; --s) if (row[s] > b) { b = row[s]; a = s; }
            best[(size_t)h * W + w] = b; disp[(size_t)h * W + w] = a;
        }
}
