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
 *
 * Alignment: the workspace must be 256-byte aligned; images and [B,H,W] results 4-byte aligned; cost_volume and
 * cost_volume_grad 16-byte aligned when the volume is banded with D % 4 == 0 (the kernels move it with 128-bit
 * accesses), 4-byte aligned otherwise.  A pointer that violates this is refused with CUSTMA_ERR_INVALID_ARGUMENT
 * (a view at an odd storage offset must be copied by the caller; custereomatching_b200.functional does).
 */
#ifndef CUSTMA_B200_H_
#define CUSTMA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUSTMA_ABI_VERSION 3

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

#define CUSTMA_FLAG_PREPARED 4u /* custma_forward / custma_forward_wta / custma_backward / custma_backward_rows:
                                   `workspace` still holds what custma_forward_prepare (custma_backward_prepare) left
                                   there for these very images (same shape, kernel size and other flags; nothing else
                                   has used the workspace since), so the image-dependent preparation is skipped */

int custma_abi_version(void);
const char *custma_last_error(void);
/* Number of CUDA kernels this library has launched in this process so far (all threads, all entry points);
 * bench.py reports the per-step difference as "gpu_launches". */
uint64_t custma_launch_count(void);

/* Host-only self-check of the tiling the fast kernels would use for this problem (every shared-memory row segment inside
 * its workspace row and 16-byte aligned, chunks covering every disparity, ...).  Needs no GPU.  0 = consistent. */
int custma_debug_validate_layout(int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size);

/* Where a custma_forward call leaves its conditioning verdict (bench.py reports it): after the call, the uint32 at byte
 * count_offset of the workspace holds the number of work items (flagged tile x 4-row groups) handed to the per-cell
 * fallback; capacity is the number of such items in the whole call.  Above 4 % of the capacity the tensor-core kernels
 * compute the call when tensor_core_available.  CUSTMA_ERR_UNSUPPORTED when the shape has no sliding-window kernel. */
int custma_debug_verdict_info(int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size, size_t *count_offset,
                              uint32_t *capacity, int32_t *tensor_core_available);

/* Workspace (device memory, 256-byte aligned) the caller must provide; depends only on the arguments shown. */
size_t custma_forward_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size, uint32_t flags);
size_t custma_backward_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size, uint32_t flags);

/* Forward: cost volume and/or winner-take-all.  Any of cost_volume / best / index may be NULL (at least one output
 * must be given; best and index come together or not at all).  Every cell of cost_volume is written exactly once
 * (the buffer need not be zeroed - the reference's torch::zeros at stereo_matching_kernel.cu:200 is not needed). */
int custma_forward(const float *camera, const float *projector, float *cost_volume, float *best, int32_t *index,
                   int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size, uint32_t flags, void *workspace,
                   size_t workspace_bytes, void *stream);

/* custma_forward plus the example-level outputs that depend only on the winner (examples/verify.py:72-74,
 * examples/test.py:78-86), written by the same kernel that decodes the WTA:
 *   mask              [B,H,W] fp32: 1 where best > mask_threshold, else 0 (cost_volume_threshold = 0.6 in verify.py:13)
 *   masked_disparity  [B,H,W] fp32: (column - correspondence) * mask, i.e. the winning disparity, zeroed where the
 *                     match is not confident
 * Either may be NULL; both need best and index. */
int custma_forward_wta(const float *camera, const float *projector, float *cost_volume, float *best, int32_t *index,
                       float *mask, float *masked_disparity, float mask_threshold, int32_t B, int32_t H, int32_t W,
                       int32_t D, int32_t kernel_size, uint32_t flags, void *workspace, size_t workspace_bytes,
                       void *stream);

/* Backward: gradient of sum(cost_volume * cost_volume_grad) with respect to the camera image only
 * (custma/stereo_matching_wrapper.py:33).  Deterministic: no global atomics anywhere.  camera_grad is overwritten. */
int custma_backward(const float *cost_volume_grad, const float *camera, const float *projector, float *camera_grad,
                    int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size, uint32_t flags, void *workspace,
                    size_t workspace_bytes, void *stream);

/* The part of custma_forward / custma_forward_wta that depends on the images only (see custma_backward_prepare below):
 * a pipeline that streams batches runs it for batch i+1 beside the kernels of batch i. */
int custma_forward_prepare(const float *camera, const float *projector, int32_t B, int32_t H, int32_t W, int32_t D,
                           int32_t kernel_size, uint32_t flags, void *workspace, size_t workspace_bytes, void *stream);

/* The part of custma_backward that depends on the images only (pivots, band copies, window statistics, conditioning
 * verdict; the window statistics of the direct kernels), run ahead of time into the backward's workspace: in a training
 * step the upstream gradient does not exist until the loss has run, but the images do - call this on a second stream
 * while the forward and the loss run (the forward keeps its own workspace), then custma_backward with
 * CUSTMA_FLAG_PREPARED on the same workspace once that stream's work is done (event).  Results are bit-identical to
 * the one-call form.  A no-op where nothing is prepared (CUSTMA_FLAG_TENSOR). */
int custma_backward_prepare(const float *camera, const float *projector, int32_t B, int32_t H, int32_t W, int32_t D,
                            int32_t kernel_size, uint32_t flags, void *workspace, size_t workspace_bytes, void *stream);

/* Backward for an upstream gradient that exists only on volume rows [row_begin, row_end): cost_volume_grad is
 * [B, row_end - row_begin, W, C]; all other rows count as zero.  camera_grad is still [B,H,W] (rows the window cannot
 * reach come out as 0).  This is what a row-band shard calls (one rank owns rows [row_begin, row_end) of the volume
 * computed on its haloed crop): no volume-sized zero fill or copy on the caller's side. */
int custma_backward_rows(const float *cost_volume_grad, const float *camera, const float *projector, float *camera_grad,
                         int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size, int32_t row_begin,
                         int32_t row_end, uint32_t flags, void *workspace, size_t workspace_bytes, void *stream);

/* Gradient of sum(cost_volume * cost_volume_grad) with respect to the PROJECTOR image.  The reference has none
 * (custma/stereo_matching_wrapper.py:33 returns None); ZNCC is symmetric in its two patches, so this is the camera
 * gradient's arithmetic (custma/src/stereo_matching_kernel.cu:75-179) with the two images' roles exchanged.  Default:
 * for a banded volume and kernel_size 3 or 5 the sliding-window backward runs on the mirrored, role-exchanged problem
 * (the workspace then also holds a sheared copy of the gradient, i.e. it is volume-sized); CUSTMA_FLAG_DIRECT, or any
 * other shape, uses the direct two-pass kernels (every kernel_size, image-sized workspace).  Deterministic, no atomics. */
size_t custma_backward_projector_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size,
                                                 uint32_t flags);
int custma_backward_projector(const float *cost_volume_grad, const float *camera, const float *projector,
                              float *projector_grad, int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size,
                              uint32_t flags, void *workspace, size_t workspace_bytes, void *stream);

/* Fused differentiable disparity head: the caller on the output side of the path (examples/verify.py:31-39 soft_argmax
 * with softargmax_beta = 50, :72-74 confidence mask; examples/test.py:79-86 disparity = column - correspondence, times the
 * mask) computed WITHOUT materialising the cost volume:
 *   w[s]           = softmax over the last axis of beta * cost (cells that do not exist take no part)
 *   soft_disparity = mask * sum_s w[s] * s        [B,H,W] fp32 (banded: s is the disparity; full: column - projector column)
 *   best, index    as custma_forward (may be NULL together);  mask [B,H,W] fp32 = best > mask_threshold (may be NULL;
 *                  pass a threshold below -1 for an all-ones mask)
 *   head_state     [B,H,W,4] fp32, 16-byte aligned: what custma_backward_head needs (may be NULL for inference)
 * custma_backward_head turns d loss / d soft_disparity into the camera gradient: every cell's upstream gradient
 * beta * w[s] * (s - soft) * mask * gd is rebuilt inside the backward kernel from head_state, so neither the volume nor a
 * volume-sized gradient is ever written or read (0 bytes per cell of HBM traffic instead of 8).
 * Sliding-window kernels only: kernel_size 3 or 5, no CUSTMA_FLAG_DIRECT / CUSTMA_FLAG_TENSOR; one workspace size
 * (custma_head_workspace_bytes) serves both calls. */
size_t custma_head_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size, uint32_t flags);
int custma_forward_head(const float *camera, const float *projector, float *soft_disparity, float *best, int32_t *index,
                        float *mask, float *head_state, float beta, float mask_threshold, int32_t B, int32_t H, int32_t W,
                        int32_t D, int32_t kernel_size, uint32_t flags, void *workspace, size_t workspace_bytes,
                        void *stream);
int custma_backward_head(const float *soft_disparity_grad, const float *camera, const float *projector,
                         const float *head_state, float beta, float *camera_grad, int32_t B, int32_t H, int32_t W, int32_t D,
                         int32_t kernel_size, uint32_t flags, void *workspace, size_t workspace_bytes, void *stream);

/* 8-bit ingestion (examples/verify.py:138-142,149 load PNGs with cv2, divide by 255 and take channel 0):
 * dst[b,h,w] = src[b,h,w,channel] * scale for an interleaved uint8 image [B,H,W,channels] in device memory. */
int custma_ingest_u8(const uint8_t *src, float *dst, int32_t B, int32_t H, int32_t W, int32_t channels, int32_t channel,
                     float scale, void *stream);

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
 * one step run under the kernels of the other, so keep two sets of host result buffers.
 * Device buffers: the library reads cost_volume_grad_dev and writes cost_volume_dev on its own non-blocking streams,
 * which are NOT ordered after the stream the caller produced the gradient on: the gradient must be complete (event or
 * stream synchronised) before the submit, and two tickets that may be in flight together must not share a
 * cost_volume_dev (pass two volumes alternately, or NULL to use the library's per-slot volumes). */
int custma_host_submit(const float *h_camera, const float *h_projector, float *h_best, int32_t *h_index,
                       float *h_camera_grad, float *cost_volume_dev, const float *cost_volume_grad_dev, int32_t B,
                       int32_t H, int32_t W, int32_t D, int32_t kernel_size, uint32_t flags, uint64_t *ticket);
/* custma_host_submit for 8-bit host images (examples/verify.py:138-142,149: cv2.imread gives uint8 [H,W,channels]; the
 * reference divides by 255 and takes channel 0 on the host).  The interleaved uint8 images [B,H,W,channels] cross PCIe as
 * they are - a quarter of the fp32 bytes per channel - and custma_ingest_u8 turns the selected channel into the fp32
 * plane (value * scale) on the device, inside the same stream as the kernels.  camera_grad is the gradient with respect
 * to that fp32 plane. */
int custma_host_submit_u8(const uint8_t *h_camera_u8, int32_t camera_channels, int32_t camera_channel,
                          const uint8_t *h_projector_u8, int32_t projector_channels, int32_t projector_channel, float scale,
                          float *h_best, int32_t *h_index, float *h_camera_grad, float *cost_volume_dev,
                          const float *cost_volume_grad_dev, int32_t B, int32_t H, int32_t W, int32_t D, int32_t kernel_size,
                          uint32_t flags, uint64_t *ticket);
int custma_host_wait(uint64_t ticket);
/* Releases the device/stream resources the host entry points cache between calls (streams, events, image / result /
 * volume buffers and workspaces for the two most recently used (device, shape, flags) combinations). */
int custma_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* CUSTMA_B200_H_ */
