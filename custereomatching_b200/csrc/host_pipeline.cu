// custma_host_step: the host-buffer entry point (include/custma_b200.h).  Images come from host memory, results go
// back to host memory; the batch is cut into chunks of pairs that are pipelined over two streams so that the
// host->device copy of chunk i+1 and the device->host copy of chunk i-1 overlap the kernels of chunk i.
// The cost volume stays in HBM.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace custma {

constexpr int kSlots = 2;
constexpr size_t kChunkVolumeBytes = (size_t)3 << 30;  // per-slot cost-volume chunk kept in HBM

struct Slot {
    cudaStream_t stream = nullptr, d2h = nullptr;   // kernels + host->device copies | device->host copies
    cudaEvent_t fwd_done = nullptr, bwd_done = nullptr, d2h_done = nullptr;
    float *cam = nullptr, *proj = nullptr, *best = nullptr, *grad = nullptr, *vol = nullptr;
    int32_t *index = nullptr;
    void *ws = nullptr;
    size_t ws_bytes = 0;
};

struct HostCtx {
    bool live = false;
    int device = -1;
    int32_t H = 0, W = 0, D = 0, k = 0, chunk = 0;
    uint32_t flags = 0;
    bool with_volume = false;
    Slot slot[kSlots];
};

static HostCtx g_ctx;
static std::mutex g_mutex;

static void release_locked() {
    if (!g_ctx.live) return;
    for (Slot &s : g_ctx.slot) {
        if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
        if (s.d2h) { cudaStreamSynchronize(s.d2h); cudaStreamDestroy(s.d2h); }
        if (s.fwd_done) cudaEventDestroy(s.fwd_done);
        if (s.bwd_done) cudaEventDestroy(s.bwd_done);
        if (s.d2h_done) cudaEventDestroy(s.d2h_done);
        cudaFree(s.cam); cudaFree(s.proj); cudaFree(s.best); cudaFree(s.grad); cudaFree(s.vol); cudaFree(s.index);
        cudaFree(s.ws);
        s = Slot();
    }
    g_ctx = HostCtx();
}

static int ensure_ctx(int32_t B, int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags, bool need_volume) {
    int dev = 0;
    CUSTMA_CUDA_CHECK(cudaGetDevice(&dev));
    const int32_t C = D > 0 ? D : W;
    const size_t pair_vol = (size_t)H * W * C * sizeof(float);
    int32_t chunk = (int32_t)std::max<size_t>(1, std::min<size_t>((size_t)B, kChunkVolumeBytes / std::max<size_t>(pair_vol, 1)));
    // two chunks in flight: the copies of one overlap the kernels of the other; larger chunks keep the kernels' grids
    // full (measured on 8 KITTI pairs: chunks of 1 / 2 / 4 / 8 pairs -> 3.54 / 3.24 / 3.03 / 3.38 ms per step)
    if (B >= kSlots) chunk = std::min(chunk, (B + kSlots - 1) / kSlots);
    if (const char *e = getenv("CUSTMA_HOST_CHUNK")) {   // tuning knob: pairs per pipelined chunk
        const int v = atoi(e);
        if (v > 0) chunk = std::min<int32_t>(chunk > 0 ? std::max(chunk, v) : v, B), chunk = std::min<int32_t>(v, B);
    }
    if (g_ctx.live && g_ctx.device == dev && g_ctx.H == H && g_ctx.W == W && g_ctx.D == D && g_ctx.k == k &&
        g_ctx.flags == flags && g_ctx.chunk >= chunk && (g_ctx.with_volume || !need_volume))
        return CUSTMA_OK;
    release_locked();
    g_ctx.live = true; g_ctx.device = dev; g_ctx.H = H; g_ctx.W = W; g_ctx.D = D; g_ctx.k = k; g_ctx.flags = flags;
    g_ctx.chunk = chunk; g_ctx.with_volume = need_volume;
    const size_t img = (size_t)chunk * H * W * sizeof(float);
    for (Slot &s : g_ctx.slot) {
        CUSTMA_CUDA_CHECK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CUSTMA_CUDA_CHECK(cudaStreamCreateWithFlags(&s.d2h, cudaStreamNonBlocking));
        CUSTMA_CUDA_CHECK(cudaEventCreateWithFlags(&s.fwd_done, cudaEventDisableTiming));
        CUSTMA_CUDA_CHECK(cudaEventCreateWithFlags(&s.bwd_done, cudaEventDisableTiming));
        CUSTMA_CUDA_CHECK(cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming));
        CUSTMA_CUDA_CHECK(cudaMalloc(&s.cam, img));
        CUSTMA_CUDA_CHECK(cudaMalloc(&s.proj, img));
        CUSTMA_CUDA_CHECK(cudaMalloc(&s.best, img));
        CUSTMA_CUDA_CHECK(cudaMalloc(&s.grad, img));
        CUSTMA_CUDA_CHECK(cudaMalloc(&s.index, (size_t)chunk * H * W * sizeof(int32_t)));
        if (need_volume) CUSTMA_CUDA_CHECK(cudaMalloc(&s.vol, (size_t)chunk * pair_vol));
        s.ws_bytes = std::max(custma_forward_workspace_bytes(chunk, H, W, D, k, flags),
                              custma_backward_workspace_bytes(chunk, H, W, D, k, flags));
        CUSTMA_CUDA_CHECK(cudaMalloc(&s.ws, s.ws_bytes));
    }
    return CUSTMA_OK;
}

}  // namespace custma

using namespace custma;

extern "C" {

int custma_host_release(void) {
    std::lock_guard<std::mutex> lock(g_mutex);
    release_locked();
    return CUSTMA_OK;
}

int custma_host_step(const float *h_camera, const float *h_projector, float *h_best, int32_t *h_index,
                     float *h_camera_grad, float *cost_volume_dev, const float *cost_volume_grad_dev, int32_t B,
                     int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags) {
    if (!h_camera || !h_projector || !h_best || !h_index)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "h_camera, h_projector, h_best and h_index must not be NULL");
    if ((cost_volume_grad_dev != nullptr) != (h_camera_grad != nullptr))
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "h_camera_grad is required iff cost_volume_grad_dev is given");
    if (B <= 0 || H <= 0 || W <= 0 || D < 0 || k < 1 || k > CUSTMA_MAX_KERNEL_SIZE)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "bad shape B=%d H=%d W=%d D=%d k=%d", B, H, W, D, k);
    std::lock_guard<std::mutex> lock(g_mutex);
    int rc = ensure_ctx(B, H, W, D, k, flags, cost_volume_dev == nullptr);
    if (rc) return rc;
    const int32_t C = D > 0 ? D : W;
    const size_t pix = (size_t)H * W;
    const int32_t chunk = g_ctx.chunk;
    int slot_i = 0;
    for (int32_t b0 = 0; b0 < B; b0 += chunk, slot_i = (slot_i + 1) % kSlots) {
        Slot &s = g_ctx.slot[slot_i];
        const int32_t nb = std::min(chunk, B - b0);
        const size_t img_bytes = (size_t)nb * pix * sizeof(float);
        // the slot's result buffers are free once the device->host copies of its previous chunk are done
        CUSTMA_CUDA_CHECK(cudaStreamWaitEvent(s.stream, s.d2h_done, 0));
        CUSTMA_CUDA_CHECK(cudaMemcpyAsync(s.cam, h_camera + (size_t)b0 * pix, img_bytes, cudaMemcpyHostToDevice, s.stream));
        CUSTMA_CUDA_CHECK(cudaMemcpyAsync(s.proj, h_projector + (size_t)b0 * pix, img_bytes, cudaMemcpyHostToDevice, s.stream));
        float *vol = cost_volume_dev ? cost_volume_dev + (size_t)b0 * pix * C : s.vol;
        rc = custma_forward(s.cam, s.proj, vol, s.best, s.index, nb, H, W, D, k, flags, s.ws, s.ws_bytes, s.stream);
        if (rc) return rc;
        // results leave on the slot's copy stream, so the backward kernels do not queue behind the copies
        CUSTMA_CUDA_CHECK(cudaEventRecord(s.fwd_done, s.stream));
        CUSTMA_CUDA_CHECK(cudaStreamWaitEvent(s.d2h, s.fwd_done, 0));
        CUSTMA_CUDA_CHECK(cudaMemcpyAsync(h_best + (size_t)b0 * pix, s.best, img_bytes, cudaMemcpyDeviceToHost, s.d2h));
        CUSTMA_CUDA_CHECK(cudaMemcpyAsync(h_index + (size_t)b0 * pix, s.index, (size_t)nb * pix * sizeof(int32_t),
                                          cudaMemcpyDeviceToHost, s.d2h));
        if (cost_volume_grad_dev) {
            rc = custma_backward(cost_volume_grad_dev + (size_t)b0 * pix * C, s.cam, s.proj, s.grad, nb, H, W, D, k,
                                 flags, s.ws, s.ws_bytes, s.stream);
            if (rc) return rc;
            CUSTMA_CUDA_CHECK(cudaEventRecord(s.bwd_done, s.stream));
            CUSTMA_CUDA_CHECK(cudaStreamWaitEvent(s.d2h, s.bwd_done, 0));
            CUSTMA_CUDA_CHECK(cudaMemcpyAsync(h_camera_grad + (size_t)b0 * pix, s.grad, img_bytes, cudaMemcpyDeviceToHost, s.d2h));
        }
        CUSTMA_CUDA_CHECK(cudaEventRecord(s.d2h_done, s.d2h));
    }
    for (Slot &s : g_ctx.slot) {
        CUSTMA_CUDA_CHECK(cudaStreamSynchronize(s.stream));
        CUSTMA_CUDA_CHECK(cudaStreamSynchronize(s.d2h));
    }
    return CUSTMA_OK;
}

}  // extern "C"
