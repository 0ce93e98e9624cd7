/* custma_b200.h - C ABI of the B200-native ZNCC cost-volume hot path (libcustma_b200.so).
 *
 * Drop-in boundary for the one data-parallel path of lzhnb/CuStereoMatching.  Every entry point takes plain
 * pointers and sizes (no torch types); device pointers are CUDA device memory of the CURRENT device, `stream` is a
 * cudaStream_t passed as void*.  All functions return CUSTMA_OK (0) or a CUSTMA_ERR_* code; custma_last_error()
 * returns a thread-local message for the last failure.  Kernels are enqueued asynchronously on `stream`; nothing
 * here synchronises unless stated.
 *
 * Reference interfaces replaced (paths relative to the reference repository root):
 *   custma_forward        <- stereo::stereo_matching_forward          custma/src/stereo_matching.cpp:16-42
 *                            stereo::stereo_matching_forward_wrapper  custma/src/stereo_matching_kernel.cu:182-216
 *                            forward_cost_volume_kernel               custma/src/stereo_matching_kernel.cu:17-72
 *                            + the WTA the examples run in torch      examples/verify.py:72-74, examples/test.py:78-86
 *   custma_backward       <- stereo::stereo_matching_backward         custma/src/stereo_matching.cpp:45-73
 *                            stereo::stereo_matching_backward_wrapper custma/src/stereo_matching_kernel.cu:218-261
 *                            get_patches_grad_kernel                  custma/src/stereo_matching_kernel.cu:75-152
 *                            patches_grad_to_image_kernel             custma/src/stereo_matching_kernel.cu:155-179
 *   declarations          <- custma/include/stereo_matching.hpp:9-16
 *
 * Layouts (row-major, fp32 unless noted):
 *   camera, projector   [B, H, W]
 *   cost volume / grad  [B, H, W, C]   C = W  when D == 0  ("full": reference-shaped, last axis = projector column d)
 *                                      C = D  when D  > 0  ("banded": last axis = disparity s, d = w - s;
 *                                                          cells with w - s < 0 hold CUSTMA_INVALID_COST, are
 *                                                          excluded from WTA and carry zero gradient)
 *   best                [B, H, W]      max over the last axis
 *   index               [B, H, W]      int32; full: first maximal projector column (torch.max semantics);
 *                                      banded: disparity s of the lowest maximal projector column (= largest s on ties)
 *   camera_grad         [B, H, W]
 * The reference has no batch axis (B == 1) and ignores D (custma/src/stereo_matching_kernel.cu:14): its behaviour is
 * B = 1, D = 0 here.
 */
#ifndef CUSTMA_B200_H_
#define CUSTMA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUSTMA_ABI_VERSION 1

#define CUSTMA_OK 0
#define CUSTMA_ERR_INVALID_ARGUMENT 1 /* null pointer, non-positive size, kernel_size out of range, ... */
#define CUSTMA_ERR_WORKSPACE 2        /* workspace missing or smaller than custma_*_workspace_bytes() */
#define CUSTMA_ERR_CUDA 3             /* a CUDA runtime call or launch failed; message holds cudaGetErrorString */
#define CUSTMA_ERR_UNSUPPORTED 4      /* device is not sm_100 (this library carries sm_100a code only) */

#define CUSTMA_INVALID_COST (-2.0f)
#define CUSTMA_EPSILON 1e-8f          /* custma/src/stereo_matching_kernel.cu:4 */
#define CUSTMA_MAX_KERNEL_SIZE 31

/* flags */
#define CUSTMA_FLAG_DIRECT 1u /* force the direct two-pass kernels (reference arithmetic order; forward is bit-exact
                                 with the reference extension).  Default: sliding-window kernels where available. */
#define CUSTMA_FLAG_TENSOR 2u /* force the tensor-core (tcgen05, 3xTF32) kernels, which the default path otherwise picks
                                 by itself for ill-conditioned (low-texture) inputs.  Needs a banded volume with
                                 D % 4 == 0, D <= 572 (forward) / 540 (backward) and k = 3 or 5;
                                 CUSTMA_ERR_UNSUPPORTED otherwise. */

int custma_abi_version(void);
const char *custma_last_error(void);
/* Number of CUDA kernels this library has launched in this process so far (all threads, all entry points);
 * bench.py reports the per-step difference as "gpu_launches". */
uint64_t custma_launch_count(void);

/* Host-only self-check of the tiling the fast kernels would use for this problem (every shared-memory row segment inside
 * its workspace row and 16-byte aligned, chunks covering every disparity, ...).  Needs no GPU.  0 = consistent. */
int custma_debug_validate_layout(int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size);

/* Workspace (device memory, 256-byte aligned) the caller must provide; depends only on the arguments shown. */
size_t custma_forward_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size, uint32_t flags);
size_t custma_backward_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size, uint32_t flags);

/* Forward: cost volume and/or winner-take-all.  Any of cost_volume / best / index may be NULL (at least one output
 * must be given; best and index come together or not at all).  Every cell of cost_volume is written exactly once
 * (the buffer need not be zeroed - the reference's torch::zeros at stereo_matching_kernel.cu:200 is not needed). */
int custma_forward(const float *camera, const float *projector, float *cost_volume, float *best, int32_t *index,
                   int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size, uint32_t flags, void *workspace,
                   size_t workspace_bytes, void *stream);

/* Backward: gradient of sum(cost_volume * cost_volume_grad) with respect to the camera image only
 * (custma/stereo_matching_wrapper.py:33).  Deterministic: no global atomics anywhere.  camera_grad is overwritten. */
int custma_backward(const float *cost_volume_grad, const float *camera, const float *projector, float *camera_grad,
                    int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size, uint32_t flags, void *workspace,
                    size_t workspace_bytes, void *stream);

/* Host-buffer step (what a non-torch caller binds): images in host memory (pinned for full copy speed), results
 * back in host memory.  Runs forward + WTA and, when cost_volume_grad_dev is non-NULL, backward, pair by pair on
 * internal streams so that copies overlap the kernels; the cost volume lives in an internal device buffer (or in
 * cost_volume_dev if given) and never crosses PCIe.  Synchronous: returns after the results are in host memory.
 *   cost_volume_grad_dev  device [B,H,W,C] upstream gradient (produced on the device by the caller's loss) or NULL
 *   h_camera_grad         host [B,H,W] or NULL (required iff cost_volume_grad_dev != NULL) */
int custma_host_step(const float *h_camera, const float *h_projector, float *h_best, int32_t *h_index,
                     float *h_camera_grad, float *cost_volume_dev, const float *cost_volume_grad_dev, int32_t B,
                     int32_t H, int32_t W, int32_t D, int32_t kernel_size, uint32_t flags);
/* The same step without the final wait, for callers that stream batches: returns once everything is enqueued and
 * stores a ticket; the host buffers of a ticket may be read (results) or rewritten (images) after
 * custma_host_wait(ticket) - 0 waits for everything submitted so far.  Consecutive submits overlap: the copies of
 * one step run under the kernels of the other, so keep two sets of host result buffers. */
int custma_host_submit(const float *h_camera, const float *h_projector, float *h_best, int32_t *h_index,
                       float *h_camera_grad, float *cost_volume_dev, const float *cost_volume_grad_dev, int32_t B,
                       int32_t H, int32_t W, int32_t D, int32_t kernel_size, uint32_t flags, uint64_t *ticket);
int custma_host_wait(uint64_t ticket);
/* Releases the device/stream resources custma_host_step caches between calls. */
int custma_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* CUSTMA_B200_H_ */
