// Shared definitions for the custma_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#include "custma_b200.h"

namespace custma {

constexpr float kEps = CUSTMA_EPSILON;          // reference: custma/src/stereo_matching_kernel.cu:4
constexpr float kInvalid = CUSTMA_INVALID_COST;

// One problem instance.  C is the length of the volume's last axis: W in full (reference-shaped) mode, D when banded.
struct Problem {
    int32_t B, H, W, D, C, k, r;  // r = k / 2 (window offsets i - r, reference kernel.cu:44-46)
    int32_t banded;               // 0: last axis = projector column d ; 1: last axis = disparity s, d = w - s
    int32_t g0, g1;               // backward only: volume rows [g0, g1) carry an upstream gradient and the gradient buffer
                                  // is [B, g1 - g0, W, C]; the other rows count as zero (custma_backward_rows)
    __host__ __device__ int32_t grows() const { return g1 - g0; }
    __host__ __device__ int64_t pixels() const { return (int64_t)B * H * W; }
    __host__ __device__ int64_t cells() const { return (int64_t)B * H * W * C; }
};

// Bounds-checked read, 0 outside the image: restates query_ij (reference kernel.cu:6-12) for one [H,W] plane.
__device__ __forceinline__ float query_ij(const float *__restrict__ img, int H, int W, int i, int j) {
    return (i < 0 || i >= H || j < 0 || j >= W) ? 0.f : __ldg(img + (int64_t)i * W + j);
}

// example-level outputs fused into the WTA decode (examples/verify.py:72-74, examples/test.py:78-86); all optional
struct WtaExtras {
    float *mask = nullptr;               // 1 where best > threshold else 0
    float *masked_disparity = nullptr;   // (column - correspondence) * mask
    float threshold = 0.6f;              // cost_volume_threshold, examples/verify.py:13
};
int launch_wta_extras(const Problem &p, const float *best, const int32_t *index, const WtaExtras &ex, cudaStream_t stream);

// Fused differentiable disparity head (examples/verify.py:31-39 soft_argmax with beta = 50, :72-74 mask; the disparity of
// examples/test.py:79-86): softmax over the last axis of beta * cost, soft disparity = sum_s w_s * s.  The forward kernels
// leave one partial (m, z, n) per pixel and slot - slot = (disparity chunk, 64-disparity unit) - with m = max cost of the
// slot's valid cells, z = sum exp(beta (c - m)), n = sum exp(beta (c - m)) * s; head_decode_kernel merges the slots.
struct HeadOut {
    float4 *part = nullptr;      // [slots][B*H*W]
    int32_t slots = 0;
    float beta_log2e = 0.f;      // beta * log2(e): exp(beta x) = exp2(beta_log2e * x)
};
// what the backward needs per pixel to rebuild the upstream gradient of every cell without reading a volume:
//   g[s] = exp2(beta_log2e * c[s] - x) * (s * y - z),  x = beta_log2e * max,  y = beta * gd * mask / Z,  z = y * soft
struct HeadGrad {
    const float4 *state = nullptr;   // [B*H*W] (x, y, z, -); nullptr: the gradient comes from a tensor as usual
    float beta_log2e = 0.f;
};
__device__ __forceinline__ float exp2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// thread-local error message storage (custma_api.cu)
int set_error(int code, const char *fmt, ...);
// multiprocessors of the current device (148 when there is none: host-only layout checks)
int device_sm_count();
// process-wide count of kernels this library has launched (custma_launch_count)
void note_launch();

#define CUSTMA_CUDA_CHECK(expr)                                                                        \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return custma::set_error(CUSTMA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                     __FILE__, __LINE__);                                              \
    } while (0)

#define CUSTMA_LAUNCH_CHECK(name)                                                                      \
    do {                                                                                               \
        cudaError_t _e = cudaGetLastError();                                                           \
        if (_e != cudaSuccess)                                                                         \
            return custma::set_error(CUSTMA_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
        custma::note_launch();                                                                         \
    } while (0)

inline size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

// Programmatic dependent launch (sm_90+): the kernels of one call form a chain of small dependent launches (prep,
// main, fallback, decode / finalize).  A chained kernel starts with pdl_wait() - it returns once the preceding kernel of
// the stream has completed and its writes are visible - and then pdl_release(), which lets the NEXT chained launch be
// scheduled while this one is still running (its blocks wait in their own pdl_wait()).  Only the launch latency between
// dependent kernels overlaps; ordering and visibility are those of plain stream order (every kernel waits before it
// reads anything).  Both instructions are no-ops for a kernel launched the ordinary way.  CUSTMA_NO_PDL=1 switches the
// launch attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                  Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---- launchers implemented in the .cu files -------------------------------------------------------------------
// per-pixel window statistics of one [B,H,W] image in the reference's own arithmetic order
int launch_window_stats(const float *img, int B, int H, int W, int k, float *mean, float *e2, cudaStream_t stream);

// direct (two-pass, reference arithmetic order) kernels: any k <= CUSTMA_MAX_KERNEL_SIZE, full or banded
int launch_direct_forward(const Problem &p, const float *cam, const float *proj, const float *cmean,
                          const float *cex2, const float *pmean, const float *pey2, float *cost, float *best,
                          int32_t *index, cudaStream_t stream);
int launch_direct_backward(const Problem &p, const float *grad, const float *cam, const float *proj,
                           const float *cmean, const float *cex2, const float *pmean, const float *pey2,
                           float *patch_grad /* [B,H,W,k*k] workspace */, float *camera_grad, cudaStream_t stream);

int launch_direct_backward_projector(const Problem &p, const float *grad, const float *cam, const float *proj,
                                     const float *cmean, const float *cex2, const float *pmean, const float *pey2,
                                     float *patch_grad /* [B,H,W,k*k] workspace */, float *projector_grad,
                                     cudaStream_t stream);

// sliding-window kernels (sliding_forward.cu / sliding_backward.cu); return CUSTMA_ERR_UNSUPPORTED if the
// (k, mode) combination has no instantiation, in which case the caller falls back to the direct kernels.
bool sliding_forward_supported(const Problem &p);
bool sliding_backward_supported(const Problem &p);
size_t sliding_forward_workspace_bytes(const Problem &p);
size_t sliding_backward_workspace_bytes(const Problem &p);
// phase: the whole call | only its image-dependent preparation (outputs and gradients unused) | everything after a
// preparation the workspace still holds (custma_forward_prepare, custma_backward_prepare / CUSTMA_FLAG_PREPARED)
enum CallPhase { kCallAll = 0, kCallPrepareOnly = 1, kCallPrepared = 2 };
int launch_sliding_forward(const Problem &p, const float *cam, const float *proj, float *cost, float *best,
                           int32_t *index, const WtaExtras &extras, void *workspace, size_t workspace_bytes,
                           bool force_tensor, cudaStream_t stream, CallPhase phase = kCallAll);
// fused head (sliding_forward.cu): soft disparity * mask, best, index, mask, and the per-pixel state of the backward
size_t sliding_head_forward_workspace_bytes(const Problem &p);
int launch_sliding_forward_head(const Problem &p, const float *cam, const float *proj, float *soft_disparity, float *best,
                                int32_t *index, float *mask, float4 *head_state, float beta, float threshold,
                                void *workspace, size_t workspace_bytes, cudaStream_t stream);
size_t sliding_head_backward_workspace_bytes(const Problem &p);
int launch_sliding_backward_head(const Problem &p, const float *soft_grad, const float *cam, const float *proj,
                                 const float4 *head_state, float beta, float *camera_grad, void *workspace,
                                 size_t workspace_bytes, cudaStream_t stream);
// tensor-core forward (tc_forward.cu): runs when fb_count is NULL or *fb_count > threshold, writes packed WTA keys
bool tc_forward_supported(const Problem &p);
int launch_tc_forward(const Problem &p, const float *cam, const float *proj, float *cost, unsigned long long *keys,
                      const uint32_t *fb_count, uint32_t threshold, cudaStream_t stream);
int launch_sliding_backward(const Problem &p, const float *grad, const float *cam, const float *proj,
                            float *camera_grad, void *workspace, size_t workspace_bytes, bool force_tensor,
                            cudaStream_t stream, CallPhase phase = kCallAll);
// tensor-core backward (tc_backward.cu): runs when fb_count is NULL or *fb_count > threshold, then writes camera_grad
bool tc_backward_supported(const Problem &p);
size_t tc_backward_scratch_bytes(const Problem &p);
int launch_tc_backward(const Problem &p, const float *grad, const float *cam, const float *proj, float *camera_grad,
                       float *scratch, const uint32_t *fb_count, uint32_t threshold, cudaStream_t stream);

}  // namespace custma
