// C-ABI entry points of libcustma_b200.so (see include/custma_b200.h): argument validation, workspace layout,
// dispatch between the sliding-window kernels and the direct two-pass kernels.
#include <stdarg.h>
#include <algorithm>
#include <atomic>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sliding_common.cuh"

namespace custma {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool pdl_enabled() {
    static const bool on = getenv("CUSTMA_NO_PDL") == nullptr;
    return on;
}

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

// SMs of the current device (B200: 148).  The tilings are sized from it; without a device (host-only layout checks)
// the B200 count is assumed.
int device_sm_count() {
    static thread_local int cached_dev = -2, cached_sms = 148;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }
    if (dev == cached_dev) return cached_sms;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { cudaGetLastError(); return 148; }
    cached_dev = dev; cached_sms = sms;
    return sms;
}

static int check_device() {
    static thread_local int checked_device = -1;
    int dev = 0;
    CUSTMA_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev == checked_device) return CUSTMA_OK;
    int major = 0;
    CUSTMA_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10)
        return set_error(CUSTMA_ERR_UNSUPPORTED,
                         "device %d has compute capability %d.x; this library carries sm_100a code only", dev, major);
    checked_device = dev;
    return CUSTMA_OK;
}

static int make_problem(int32_t B, int32_t H, int32_t W, int32_t D, int32_t k, Problem *p) {
    if (B <= 0 || H <= 0 || W <= 0) return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "B, H, W must be positive (got %d, %d, %d)", B, H, W);
    if (D < 0) return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "D must be >= 0 (0 selects the reference-shaped [H,W,W] volume), got %d", D);
    if (k < 1 || k > CUSTMA_MAX_KERNEL_SIZE)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "kernel_size must be in [1, %d], got %d", CUSTMA_MAX_KERNEL_SIZE, k);
    p->B = B; p->H = H; p->W = W; p->D = D; p->k = k; p->r = k / 2;
    p->banded = D > 0;
    p->C = D > 0 ? D : W;
    p->g0 = 0; p->g1 = H;
    return CUSTMA_OK;
}

// workspace layout: [cam mean][cam e2][proj mean][proj e2][path-specific...]
struct StatsPtrs { float *cmean, *cex2, *pmean, *pey2; char *rest; size_t rest_bytes; };

static size_t stats_bytes(const Problem &p) { return 4 * align256((size_t)p.pixels() * sizeof(float)); }

static int carve_stats(const Problem &p, void *ws, size_t ws_bytes, size_t need, StatsPtrs *s) {
    if (!ws || ws_bytes < need)
        return set_error(CUSTMA_ERR_WORKSPACE, "workspace of %zu bytes required, %zu given", need, ws ? ws_bytes : (size_t)0);
    if (((uintptr_t)ws & 255) != 0) return set_error(CUSTMA_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    const size_t one = align256((size_t)p.pixels() * sizeof(float));
    char *base = (char *)ws;
    s->cmean = (float *)base; s->cex2 = (float *)(base + one);
    s->pmean = (float *)(base + 2 * one); s->pey2 = (float *)(base + 3 * one);
    s->rest = base + 4 * one; s->rest_bytes = ws_bytes - 4 * one;
    return CUSTMA_OK;
}

// The vectorised banded kernels (D % 4 == 0) move the volume with 16-byte stores / cp.async: a misaligned pointer would
// raise a sticky "misaligned address" fault instead of an error code, so it is rejected here.  Volumes whose last axis
// is not a multiple of 4 floats are accessed element by element and need only 4-byte alignment.
static int check_volume_alignment(const Problem &p, const void *ptr, const char *name) {
    const uintptr_t need = (p.banded && (p.D & 3) == 0) ? 15u : 3u;
    if (ptr && ((uintptr_t)ptr & need) != 0)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "%s must be %u-byte aligned for this shape (got %p)", name,
                         (unsigned)need + 1u, ptr);
    return CUSTMA_OK;
}
static int check_image_alignment(const void *ptr, const char *name) {
    if (ptr && ((uintptr_t)ptr & 3u) != 0)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "%s must be 4-byte aligned (got %p)", name, ptr);
    return CUSTMA_OK;
}

static bool use_sliding_fwd(const Problem &p, uint32_t flags) {
    return !(flags & CUSTMA_FLAG_DIRECT) && sliding_forward_supported(p);
}
static bool use_sliding_bwd(const Problem &p, uint32_t flags) {
    return !(flags & CUSTMA_FLAG_DIRECT) && sliding_backward_supported(p);
}

static size_t forward_ws(const Problem &p, uint32_t flags) {
    return stats_bytes(p) + (use_sliding_fwd(p, flags) ? sliding_forward_workspace_bytes(p) : 0);
}
static size_t backward_ws(const Problem &p, uint32_t flags) {
    if (use_sliding_bwd(p, flags)) return stats_bytes(p) + sliding_backward_workspace_bytes(p);
    return stats_bytes(p) + align256((size_t)p.pixels() * p.k * p.k * sizeof(float));
}

}  // namespace custma

using namespace custma;

extern "C" {

int custma_abi_version(void) { return CUSTMA_ABI_VERSION; }
const char *custma_last_error(void) { return g_error; }
uint64_t custma_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t custma_forward_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags) {
    flags &= ~CUSTMA_FLAG_PREPARED;
    Problem p;
    if (make_problem(B, H, W, D, k, &p) != CUSTMA_OK) return 0;
    return forward_ws(p, flags);
}

size_t custma_backward_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags) {
    flags &= ~CUSTMA_FLAG_PREPARED;
    Problem p;
    if (make_problem(B, H, W, D, k, &p) != CUSTMA_OK) return 0;
    return backward_ws(p, flags);
}

int custma_debug_verdict_info(int32_t B, int32_t H, int32_t W, int32_t D, int32_t k, size_t *count_offset,
                              uint32_t *capacity, int32_t *tensor_core_available) {
    Problem p;
    int rc = make_problem(B, H, W, D, k, &p);
    if (rc) return rc;
    SlidingConfig cfg;
    if (!sliding_pick_config(p, false, &cfg))
        return set_error(CUSTMA_ERR_UNSUPPORTED, "no sliding-window kernel (hence no verdict) for k=%d", k);
    SlidingLayout L;
    make_sliding_layout(p, cfg, false, &L);
    if (count_offset) *count_offset = stats_bytes(p) + L.off_fb_count;
    if (capacity) *capacity = (uint32_t)((int64_t)p.B * L.NB * L.n_wtiles * L.fb_groups);
    if (tensor_core_available) *tensor_core_available = tc_forward_supported(p) ? 1 : 0;
    return CUSTMA_OK;
}

int custma_debug_validate_layout(int32_t B, int32_t H, int32_t W, int32_t D, int32_t k) {
    Problem p;
    int rc = make_problem(B, H, W, D, k, &p);
    if (rc) return rc;
    if ((rc = validate_sliding_layout(p, false))) return rc;
    return validate_sliding_layout(p, true);
}

static int forward_impl(const float *camera, const float *projector, float *cost_volume, float *best, int32_t *index,
                        const WtaExtras &extras, int32_t B, int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags,
                        void *workspace, size_t workspace_bytes, void *stream_, bool prepare_only = false) {
    Problem p;
    int rc = make_problem(B, H, W, D, k, &p);
    if (rc) return rc;
    const CallPhase phase = prepare_only ? kCallPrepareOnly : (flags & CUSTMA_FLAG_PREPARED) ? kCallPrepared : kCallAll;
    flags &= ~CUSTMA_FLAG_PREPARED;
    if (!camera || !projector) return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "camera and projector must not be NULL");
    if (!prepare_only && !cost_volume && !best)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "no output requested (cost_volume and best are both NULL)");
    if ((best == nullptr) != (index == nullptr)) return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "best and index must be given together");
    if ((extras.mask || extras.masked_disparity) && !best)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "mask / masked_disparity need best and index");
    if ((rc = check_volume_alignment(p, cost_volume, "cost_volume"))) return rc;
    if ((rc = check_image_alignment(camera, "camera")) || (rc = check_image_alignment(projector, "projector")) ||
        (rc = check_image_alignment(best, "best")) || (rc = check_image_alignment(index, "index")) ||
        (rc = check_image_alignment(extras.mask, "mask")) || (rc = check_image_alignment(extras.masked_disparity, "masked_disparity")))
        return rc;
    if ((rc = check_device())) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    StatsPtrs s;
    if ((rc = carve_stats(p, workspace, workspace_bytes, forward_ws(p, flags), &s))) return rc;
    if ((flags & CUSTMA_FLAG_TENSOR) && ((flags & CUSTMA_FLAG_DIRECT) || !tc_forward_supported(p) || !sliding_forward_supported(p)))
        return set_error(CUSTMA_ERR_UNSUPPORTED, "CUSTMA_FLAG_TENSOR needs a banded volume with D %% 4 == 0, D <= 572, k = 3 or 5, and no CUSTMA_FLAG_DIRECT");
    if (use_sliding_fwd(p, flags))
        return launch_sliding_forward(p, camera, projector, cost_volume, best, index, extras, s.rest, s.rest_bytes,
                                      (flags & CUSTMA_FLAG_TENSOR) != 0, stream, phase);
    if (phase != kCallPrepared) {   // the direct kernels' preparation: window means and second moments of both images
        if ((rc = launch_window_stats(camera, B, H, W, k, s.cmean, s.cex2, stream))) return rc;
        if ((rc = launch_window_stats(projector, B, H, W, k, s.pmean, s.pey2, stream))) return rc;
    }
    if (phase == kCallPrepareOnly) return CUSTMA_OK;
    if ((rc = launch_direct_forward(p, camera, projector, s.cmean, s.cex2, s.pmean, s.pey2, cost_volume, best, index, stream)))
        return rc;
    return best ? launch_wta_extras(p, best, index, extras, stream) : CUSTMA_OK;
}

int custma_forward(const float *camera, const float *projector, float *cost_volume, float *best, int32_t *index,
                   int32_t B, int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags, void *workspace,
                   size_t workspace_bytes, void *stream_) {
    return forward_impl(camera, projector, cost_volume, best, index, WtaExtras(), B, H, W, D, k, flags, workspace,
                        workspace_bytes, stream_);
}

int custma_forward_prepare(const float *camera, const float *projector, int32_t B, int32_t H, int32_t W, int32_t D,
                           int32_t k, uint32_t flags, void *workspace, size_t workspace_bytes, void *stream_) {
    return forward_impl(camera, projector, nullptr, nullptr, nullptr, WtaExtras(), B, H, W, D, k, flags & ~CUSTMA_FLAG_PREPARED,
                        workspace, workspace_bytes, stream_, true);
}

int custma_forward_wta(const float *camera, const float *projector, float *cost_volume, float *best, int32_t *index,
                       float *mask, float *masked_disparity, float mask_threshold, int32_t B, int32_t H, int32_t W,
                       int32_t D, int32_t k, uint32_t flags, void *workspace, size_t workspace_bytes, void *stream_) {
    WtaExtras ex;
    ex.mask = mask; ex.masked_disparity = masked_disparity; ex.threshold = mask_threshold;
    return forward_impl(camera, projector, cost_volume, best, index, ex, B, H, W, D, k, flags, workspace, workspace_bytes,
                        stream_);
}

static int backward_impl(const float *cost_volume_grad, const float *camera, const float *projector, float *camera_grad,
                         int32_t B, int32_t H, int32_t W, int32_t D, int32_t k, int32_t row_begin, int32_t row_end,
                         uint32_t flags, void *workspace, size_t workspace_bytes, void *stream_, bool prepare_only = false) {
    Problem p;
    int rc = make_problem(B, H, W, D, k, &p);
    if (rc) return rc;
    const CallPhase phase = prepare_only ? kCallPrepareOnly : (flags & CUSTMA_FLAG_PREPARED) ? kCallPrepared : kCallAll;
    flags &= ~CUSTMA_FLAG_PREPARED;
    if (prepare_only) {
        if (!camera || !projector) return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "camera and projector must not be NULL");
    } else if (!cost_volume_grad || !camera || !projector || !camera_grad)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "cost_volume_grad, camera, projector and camera_grad must not be NULL");
    if (row_begin < 0 || row_end > H || row_begin >= row_end)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "gradient rows [%d, %d) must be a non-empty range inside [0, %d)", row_begin, row_end, H);
    p.g0 = row_begin; p.g1 = row_end;
    if (!prepare_only && (rc = check_volume_alignment(p, cost_volume_grad, "cost_volume_grad"))) return rc;
    if ((rc = check_image_alignment(camera, "camera")) || (rc = check_image_alignment(projector, "projector")) ||
        (!prepare_only && (rc = check_image_alignment(camera_grad, "camera_grad"))))
        return rc;
    if ((rc = check_device())) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    StatsPtrs s;
    if ((rc = carve_stats(p, workspace, workspace_bytes, backward_ws(p, flags), &s))) return rc;
    if ((flags & CUSTMA_FLAG_TENSOR) && ((flags & CUSTMA_FLAG_DIRECT) || !tc_backward_supported(p) || !sliding_backward_supported(p)))
        return set_error(CUSTMA_ERR_UNSUPPORTED, "CUSTMA_FLAG_TENSOR needs a banded volume with D %% 4 == 0, D <= 540, k = 3 or 5, and no CUSTMA_FLAG_DIRECT");
    if (use_sliding_bwd(p, flags))
        return launch_sliding_backward(p, cost_volume_grad, camera, projector, camera_grad, s.rest, s.rest_bytes,
                                       (flags & CUSTMA_FLAG_TENSOR) != 0, stream, phase);
    if (phase != kCallPrepared) {   // the direct kernels' preparation: window means and second moments of both images
        if ((rc = launch_window_stats(camera, B, H, W, k, s.cmean, s.cex2, stream))) return rc;
        if ((rc = launch_window_stats(projector, B, H, W, k, s.pmean, s.pey2, stream))) return rc;
    }
    if (phase == kCallPrepareOnly) return CUSTMA_OK;
    return launch_direct_backward(p, cost_volume_grad, camera, projector, s.cmean, s.cex2, s.pmean, s.pey2,
                                  (float *)s.rest, camera_grad, stream);
}

int custma_backward(const float *cost_volume_grad, const float *camera, const float *projector, float *camera_grad,
                    int32_t B, int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags, void *workspace,
                    size_t workspace_bytes, void *stream_) {
    return backward_impl(cost_volume_grad, camera, projector, camera_grad, B, H, W, D, k, 0, H, flags, workspace,
                         workspace_bytes, stream_);
}

int custma_backward_prepare(const float *camera, const float *projector, int32_t B, int32_t H, int32_t W, int32_t D,
                            int32_t k, uint32_t flags, void *workspace, size_t workspace_bytes, void *stream_) {
    return backward_impl(nullptr, camera, projector, nullptr, B, H, W, D, k, 0, H, flags & ~CUSTMA_FLAG_PREPARED, workspace,
                         workspace_bytes, stream_, true);
}

int custma_backward_rows(const float *cost_volume_grad, const float *camera, const float *projector, float *camera_grad,
                         int32_t B, int32_t H, int32_t W, int32_t D, int32_t k, int32_t row_begin, int32_t row_end,
                         uint32_t flags, void *workspace, size_t workspace_bytes, void *stream_) {
    return backward_impl(cost_volume_grad, camera, projector, camera_grad, B, H, W, D, k, row_begin, row_end, flags,
                         workspace, workspace_bytes, stream_);
}

// ---- gradient with respect to the projector image (the reference returns None for it) -----------
// ZNCC is symmetric in its two patches.  Two routes:
//   direct   direct_patch_grad_kernel<SWAP> (direct.cu): a warp per projector pixel, O(k^2) per cell, every k
//   fast     (banded volume, odd k with a sliding-window backward, i.e. k = 3 or 5) the camera-gradient kernels on the
//            mirrored, role-exchanged problem: camera' = flip_x(projector), projector' = flip_x(camera) - mirroring
//            turns "projector column = camera column + s" into "- s", which is what the banded kernels index - with the
//            upstream gradient sheared accordingly, g'[h, c', s] = g[h, W-1-c'+s, s] (a bijection between the valid
//            cells of the two volumes), and projector_grad = flip_x(camera_grad').  Costs one extra pass over the
//            gradient (read once, written once into the workspace) next to the O(1)-per-cell backward.
__global__ void __launch_bounds__(256)
    flip_images_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ a_out,
                       float *__restrict__ b_out, int64_t rows, int W) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * W) return;
    const int64_t r = i / W;
    const int x = (int)(i - r * W);
    a_out[i] = a[r * W + (W - 1 - x)];
    if (b) b_out[i] = b[r * W + (W - 1 - x)];
}

constexpr int kShearPitch = 34;   // (pitch + 1) odd: the diagonal reads of the tile hit 32 different banks
__global__ void __launch_bounds__(256)
    shear_gradient_kernel(const float *__restrict__ g, float *__restrict__ gs, int W, int D) {
    __shared__ float tile[63][kShearPitch];
    const int s0 = blockIdx.x * 32, c0 = blockIdx.y * 32, tx = threadIdx.x, ty = threadIdx.y;
    const int64_t plane = (int64_t)blockIdx.z * W * D;
    const int w_lo = W - 1 - (c0 + 31) + s0, s = s0 + tx;
    for (int r = ty; r < 63; r += 8) {
        const int w = w_lo + r;
        tile[r][tx] = (w >= 0 && w < W && s < D) ? __ldcs(g + plane + (int64_t)w * D + s) : 0.f;
    }
    __syncthreads();
    for (int cy = ty; cy < 32; cy += 8) {
        const int c = c0 + cy;
        if (c < W && s < D) {
            const int w = W - 1 - c + s;                       // row of the source cell; beyond the image: no cell
            __stcs(gs + plane + (int64_t)c * D + s, w < W ? tile[31 - cy + tx][tx] : 0.f);
        }
    }
}

static bool projector_fast_path(const Problem &p, uint32_t flags) {
    return !(flags & CUSTMA_FLAG_DIRECT) && p.banded && (p.k & 1) && sliding_backward_supported(p);
}

size_t custma_backward_projector_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags) {
    Problem p;
    if (make_problem(B, H, W, D, k, &p) != CUSTMA_OK) return 0;
    if (projector_fast_path(p, flags))
        return backward_ws(p, 0) + align256((size_t)p.cells() * sizeof(float)) + 3 * align256((size_t)p.pixels() * sizeof(float));
    return stats_bytes(p) + align256((size_t)p.pixels() * p.k * p.k * sizeof(float));
}

int custma_backward_projector(const float *cost_volume_grad, const float *camera, const float *projector,
                              float *projector_grad, int32_t B, int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags,
                              void *workspace, size_t workspace_bytes, void *stream_) {
    Problem p;
    int rc = make_problem(B, H, W, D, k, &p);
    if (rc) return rc;
    if (!cost_volume_grad || !camera || !projector || !projector_grad)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "cost_volume_grad, camera, projector and projector_grad must not be NULL");
    if (flags & CUSTMA_FLAG_TENSOR) return set_error(CUSTMA_ERR_UNSUPPORTED, "no tensor-core kernel for the projector gradient");
    if ((rc = check_image_alignment(cost_volume_grad, "cost_volume_grad")) || (rc = check_image_alignment(camera, "camera")) ||
        (rc = check_image_alignment(projector, "projector")) || (rc = check_image_alignment(projector_grad, "projector_grad")))
        return rc;
    if ((rc = check_device())) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t need = custma_backward_projector_workspace_bytes(B, H, W, D, k, flags);
    if (projector_fast_path(p, flags)) {
        if (!workspace || workspace_bytes < need)
            return set_error(CUSTMA_ERR_WORKSPACE, "workspace of %zu bytes required, %zu given", need, workspace ? workspace_bytes : (size_t)0);
        if (((uintptr_t)workspace & 255) != 0) return set_error(CUSTMA_ERR_WORKSPACE, "workspace must be 256-byte aligned");
        char *base = (char *)workspace;
        const size_t bws = backward_ws(p, 0), img = align256((size_t)p.pixels() * sizeof(float));
        float *gs = (float *)(base + bws);
        float *camF = (float *)(base + bws + align256((size_t)p.cells() * sizeof(float)));
        float *projF = (float *)((char *)camF + img), *gradF = (float *)((char *)projF + img);
        const int64_t rows = (int64_t)B * H;
        const unsigned fb = (unsigned)((rows * W + 255) / 256);
        // camera' = mirrored projector, projector' = mirrored camera
        flip_images_kernel<<<fb, 256, 0, stream>>>(projector, camera, camF, projF, rows, W);
        CUSTMA_LAUNCH_CHECK("flip_images_kernel");
        if (rows > 65535 * 1ll) {   // gridDim.z limit: one launch per slab of image rows
            for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
                const int64_t nr = std::min<int64_t>(65535, rows - r0);
                shear_gradient_kernel<<<dim3((D + 31) / 32, (W + 31) / 32, (unsigned)nr), dim3(32, 8), 0, stream>>>(
                    cost_volume_grad + r0 * W * D, gs + r0 * W * D, W, D);
                CUSTMA_LAUNCH_CHECK("shear_gradient_kernel");
            }
        } else {
            shear_gradient_kernel<<<dim3((D + 31) / 32, (W + 31) / 32, (unsigned)rows), dim3(32, 8), 0, stream>>>(cost_volume_grad, gs, W, D);
            CUSTMA_LAUNCH_CHECK("shear_gradient_kernel");
        }
        StatsPtrs s;
        if ((rc = carve_stats(p, workspace, bws, bws, &s))) return rc;
        if ((rc = launch_sliding_backward(p, gs, camF, projF, gradF, s.rest, s.rest_bytes, false, stream))) return rc;
        flip_images_kernel<<<fb, 256, 0, stream>>>(gradF, nullptr, projector_grad, nullptr, rows, W);
        CUSTMA_LAUNCH_CHECK("flip_images_kernel");
        return CUSTMA_OK;
    }
    StatsPtrs s;
    if ((rc = carve_stats(p, workspace, workspace_bytes, need, &s))) return rc;
    if ((rc = launch_window_stats(camera, B, H, W, k, s.cmean, s.cex2, stream))) return rc;
    if ((rc = launch_window_stats(projector, B, H, W, k, s.pmean, s.pey2, stream))) return rc;
    return launch_direct_backward_projector(p, cost_volume_grad, camera, projector, s.cmean, s.cex2, s.pmean, s.pey2,
                                            (float *)s.rest, projector_grad, stream);
}

// ---- fused differentiable disparity head -------------------------------------------------------
size_t custma_head_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags) {
    Problem p;
    if (make_problem(B, H, W, D, k, &p) != CUSTMA_OK || (flags & (CUSTMA_FLAG_DIRECT | CUSTMA_FLAG_TENSOR))) return 0;
    const size_t f = sliding_head_forward_workspace_bytes(p), b = sliding_head_backward_workspace_bytes(p);
    if (f == 0 || b == 0) return 0;
    return stats_bytes(p) + (f > b ? f : b);
}

int custma_forward_head(const float *camera, const float *projector, float *soft_disparity, float *best, int32_t *index,
                        float *mask, float *head_state, float beta, float mask_threshold, int32_t B, int32_t H, int32_t W,
                        int32_t D, int32_t k, uint32_t flags, void *workspace, size_t workspace_bytes, void *stream_) {
    Problem p;
    int rc = make_problem(B, H, W, D, k, &p);
    if (rc) return rc;
    if (!camera || !projector || !soft_disparity)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "camera, projector and soft_disparity must not be NULL");
    if ((best == nullptr) != (index == nullptr)) return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "best and index must be given together");
    if (!(beta > 0.f) || !(beta <= 1e4f)) return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "beta must be in (0, 1e4], got %g", (double)beta);
    if (flags & (CUSTMA_FLAG_DIRECT | CUSTMA_FLAG_TENSOR))
        return set_error(CUSTMA_ERR_UNSUPPORTED, "the fused head runs on the sliding-window kernels only (no CUSTMA_FLAG_DIRECT / CUSTMA_FLAG_TENSOR)");
    if (head_state && ((uintptr_t)head_state & 15u)) return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "head_state must be 16-byte aligned");
    if ((rc = check_image_alignment(camera, "camera")) || (rc = check_image_alignment(projector, "projector")) ||
        (rc = check_image_alignment(soft_disparity, "soft_disparity")) || (rc = check_image_alignment(best, "best")) ||
        (rc = check_image_alignment(index, "index")) || (rc = check_image_alignment(mask, "mask")))
        return rc;
    if ((rc = check_device())) return rc;
    const size_t need = custma_head_workspace_bytes(B, H, W, D, k, flags);
    if (need == 0) return set_error(CUSTMA_ERR_UNSUPPORTED, "the fused head needs kernel_size 3 or 5 (forward and backward fast paths), got %d", k);
    StatsPtrs s;
    if ((rc = carve_stats(p, workspace, workspace_bytes, need, &s))) return rc;
    return launch_sliding_forward_head(p, camera, projector, soft_disparity, best, index, mask, (float4 *)head_state, beta,
                                       mask_threshold, s.rest, s.rest_bytes, (cudaStream_t)stream_);
}

int custma_backward_head(const float *soft_disparity_grad, const float *camera, const float *projector,
                         const float *head_state, float beta, float *camera_grad, int32_t B, int32_t H, int32_t W, int32_t D,
                         int32_t k, uint32_t flags, void *workspace, size_t workspace_bytes, void *stream_) {
    Problem p;
    int rc = make_problem(B, H, W, D, k, &p);
    if (rc) return rc;
    if (!soft_disparity_grad || !camera || !projector || !head_state || !camera_grad)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "soft_disparity_grad, camera, projector, head_state and camera_grad must not be NULL");
    if (!(beta > 0.f) || !(beta <= 1e4f)) return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "beta must be in (0, 1e4], got %g", (double)beta);
    if (flags & (CUSTMA_FLAG_DIRECT | CUSTMA_FLAG_TENSOR))
        return set_error(CUSTMA_ERR_UNSUPPORTED, "the fused head runs on the sliding-window kernels only (no CUSTMA_FLAG_DIRECT / CUSTMA_FLAG_TENSOR)");
    if ((uintptr_t)head_state & 15u) return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "head_state must be 16-byte aligned");
    if ((rc = check_image_alignment(camera, "camera")) || (rc = check_image_alignment(projector, "projector")) ||
        (rc = check_image_alignment(soft_disparity_grad, "soft_disparity_grad")) || (rc = check_image_alignment(camera_grad, "camera_grad")))
        return rc;
    if ((rc = check_device())) return rc;
    const size_t need = custma_head_workspace_bytes(B, H, W, D, k, flags);
    if (need == 0) return set_error(CUSTMA_ERR_UNSUPPORTED, "the fused head needs kernel_size 3 or 5, got %d", k);
    StatsPtrs s;
    if ((rc = carve_stats(p, workspace, workspace_bytes, need, &s))) return rc;
    return launch_sliding_backward_head(p, soft_disparity_grad, camera, projector, (const float4 *)head_state, beta,
                                        camera_grad, s.rest, s.rest_bytes, (cudaStream_t)stream_);
}

// uint8 ingestion (examples/verify.py:138-142,149: cv2.imread(...) / 255, channel 0 of the RGB image): one pass from the
// interleaved 8-bit image to the fp32 plane the kernels read
__global__ void __launch_bounds__(256)
    ingest_u8_kernel(const uint8_t *__restrict__ src, float *__restrict__ dst, int64_t pixels, int channels, int channel,
                     float scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pixels) dst[i] = (float)src[i * channels + channel] * scale;
}

int custma_ingest_u8(const uint8_t *src, float *dst, int32_t B, int32_t H, int32_t W, int32_t channels, int32_t channel,
                     float scale, void *stream_) {
    if (!src || !dst) return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "src and dst must not be NULL");
    if (B <= 0 || H <= 0 || W <= 0 || channels < 1 || channel < 0 || channel >= channels)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "bad shape / channel: B=%d H=%d W=%d channels=%d channel=%d", B, H, W, channels, channel);
    int rc = check_image_alignment(dst, "dst");
    if (rc) return rc;
    if ((rc = check_device())) return rc;
    const int64_t pixels = (int64_t)B * H * W;
    ingest_u8_kernel<<<(unsigned)((pixels + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(src, dst, pixels, channels, channel, scale);
    CUSTMA_LAUNCH_CHECK("ingest_u8_kernel");
    return CUSTMA_OK;
}

}  // extern "C"
