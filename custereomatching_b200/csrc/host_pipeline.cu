// custma_host_step: the host-buffer entry point (include/custma_b200.h).  Images come from host memory, results go
// back to host memory; the batch is cut into chunks of pairs that are pipelined so that the host->device copy of
// chunk i+1 and the device->host copy of chunk i-1 overlap the kernels of chunk i: one stream for the copies in, one
// for all kernels, one for the copies out, two sets of device buffers ("slots").  The cost volume stays in HBM.
// (Round 1 gave every slot its own kernel stream; the kernels of two steps then shared the SMs whenever both were
// ready, and a streamed step of 8 KITTI pairs took 2.74 ms instead of 2.47 ms - the device-resident step is 2.45 ms.
// CUSTMA_HOST_PIPELINE=slots still selects that shape.)
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace custma {

constexpr int kSlots = 2;
constexpr int kTickets = 8;   // completion events kept for custma_host_wait
constexpr size_t kChunkVolumeBytes = (size_t)3 << 30;  // per-slot cost-volume chunk kept in HBM
constexpr size_t kSingleStreamCells = (size_t)256 << 20;   // chunks of at least this many cells get the compute stream to themselves

struct Slot {
    cudaStream_t stream = nullptr, d2h = nullptr;   // kernels (+ host->device copies in the per-slot pipeline) | device->host copies
    cudaEvent_t fwd_done = nullptr, bwd_done = nullptr, d2h_done = nullptr;
    cudaEvent_t h2d_done = nullptr, kernels_done = nullptr;   // single-compute-stream pipeline: inputs on the device | inputs free again
    cudaEvent_t fprep_done = nullptr, prep_done = nullptr;   // the image-dependent preparations (side stream) are in ws | in ws_bwd
    void *ws_bwd = nullptr;            // single-compute-stream pipeline: the backward's own workspace
    size_t ws_bwd_bytes = 0;
    float *cam = nullptr, *proj = nullptr, *best = nullptr, *grad = nullptr, *vol = nullptr;
    int32_t *index = nullptr;
    void *ws = nullptr;
    size_t ws_bytes = 0;
    uint8_t *u8 = nullptr;          // staging of 8-bit host images (custma_host_submit_u8), grown on demand
    size_t u8_bytes = 0;
};

struct HostCtx {
    bool live = false;
    int device = -1;
    int32_t H = 0, W = 0, D = 0, k = 0, chunk = 0;
    uint32_t flags = 0;
    bool with_volume = false;
    Slot slot[kSlots];
    // pipeline shape: one compute stream for every chunk (kernels of different chunks never share the SMs) fed by a
    // host->device stream and drained by a device->host stream, or (CUSTMA_HOST_PIPELINE=slots) kernels and
    // host->device copies of a chunk on its slot's own stream
    bool single = true;
    cudaStream_t compute = nullptr, h2d = nullptr, d2h = nullptr;
    cudaStream_t prep = nullptr;   // lowest priority: the next backward's preparation fills what the forward leaves idle
    int next_slot = 0;                                 // slots keep rotating across calls, so consecutive steps overlap
    uint64_t next_ticket = 1;
    cudaEvent_t done[kTickets][kSlots] = {};           // done[t % kTickets][s]: slot s has delivered everything of ticket t
};

// Two contexts are kept (least recently used one is rebuilt): a caller that alternates between two shapes - or two
// devices - does not tear its streams and buffers down on every call.  A ticket carries its context in the top byte.
constexpr int kContexts = 2;
static HostCtx g_ctxs[kContexts];
static uint64_t g_use_stamp[kContexts] = {};
static uint64_t g_stamp = 0;
static std::mutex g_mutex;

static void release_locked(HostCtx &g_ctx) {
    if (!g_ctx.live) return;
    int restore = -1;
    cudaGetDevice(&restore);
    cudaSetDevice(g_ctx.device);
    for (Slot &s : g_ctx.slot) {
        if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
        if (s.d2h) { cudaStreamSynchronize(s.d2h); cudaStreamDestroy(s.d2h); }
        if (s.fwd_done) cudaEventDestroy(s.fwd_done);
        if (s.bwd_done) cudaEventDestroy(s.bwd_done);
        if (s.d2h_done) cudaEventDestroy(s.d2h_done);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
        if (s.kernels_done) cudaEventDestroy(s.kernels_done);
        if (s.prep_done) cudaEventDestroy(s.prep_done);
        if (s.fprep_done) cudaEventDestroy(s.fprep_done);
        cudaFree(s.ws_bwd);
        cudaFree(s.cam); cudaFree(s.proj); cudaFree(s.best); cudaFree(s.grad); cudaFree(s.vol); cudaFree(s.index);
        cudaFree(s.ws); cudaFree(s.u8);
        s = Slot();
    }
    for (cudaStream_t st : {g_ctx.compute, g_ctx.h2d, g_ctx.d2h, g_ctx.prep})
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (auto &row : g_ctx.done)
        for (cudaEvent_t &e : row)
            if (e) cudaEventDestroy(e);
    g_ctx = HostCtx();
    if (restore >= 0) cudaSetDevice(restore);
}

// Pairs per pipelined chunk.  A synchronous step cuts the batch in two so that the copies of one half overlap the
// kernels of the other (measured on 8 KITTI pairs: chunks of 1 / 2 / 4 / 8 pairs -> 3.54 / 3.24 / 3.03 / 3.38 ms per
// step); a streamed step (custma_host_submit) overlaps with its neighbours instead and keeps the batch whole, so the
// kernels' grids stay full.
static int32_t pick_chunk(int32_t B, int32_t H, int32_t W, int32_t D, bool streamed) {
    const int32_t C = D > 0 ? D : W;
    const size_t pair_vol = (size_t)H * W * C * sizeof(float);
    int32_t chunk = (int32_t)std::max<size_t>(1, std::min<size_t>((size_t)B, kChunkVolumeBytes / std::max<size_t>(pair_vol, 1)));
    if (!streamed && B >= kSlots) chunk = std::min(chunk, (B + kSlots - 1) / kSlots);
    if (const char *e = getenv("CUSTMA_HOST_CHUNK")) {   // tuning knob
        const int v = atoi(e);
        if (v > 0) chunk = std::min<int32_t>(v, B);
    }
    return chunk;
}

// Builds every stream, event and buffer of the context.  g_ctx.live is set first so that release_locked() can take a
// half-built context apart; ensure_ctx() does exactly that when anything here fails, so a later call never finds a
// context that passes the cache test with null buffers in it.
static int build_ctx(HostCtx &g_ctx, int32_t chunk, int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags,
                     bool need_volume, int dev) {
    const int32_t C = D > 0 ? D : W;
    const size_t pair_vol = (size_t)H * W * C * sizeof(float);
    g_ctx.live = true; g_ctx.device = dev; g_ctx.H = H; g_ctx.W = W; g_ctx.D = D; g_ctx.k = k; g_ctx.flags = flags;
    g_ctx.chunk = chunk; g_ctx.with_volume = need_volume;
    const size_t img = (size_t)chunk * H * W * sizeof(float);
    const char *shape = getenv("CUSTMA_HOST_PIPELINE");
    // a chunk that fills the device runs alone; small chunks (one KITTI pair is 89 Mcell) gain more from sharing the SMs
    // with the next step's kernels than they lose (measured: 0.42 against 0.49 ms per step)
    g_ctx.single = (size_t)chunk * H * W * C >= kSingleStreamCells;
    if (shape && strcmp(shape, "slots") == 0) g_ctx.single = false;
    if (shape && strcmp(shape, "single") == 0) g_ctx.single = true;
    if (g_ctx.single) {
        int prio_low = 0, prio_high = 0;
        CUSTMA_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
        CUSTMA_CUDA_CHECK(cudaStreamCreateWithPriority(&g_ctx.compute, cudaStreamNonBlocking, prio_high));
        CUSTMA_CUDA_CHECK(cudaStreamCreateWithFlags(&g_ctx.h2d, cudaStreamNonBlocking));
        CUSTMA_CUDA_CHECK(cudaStreamCreateWithFlags(&g_ctx.d2h, cudaStreamNonBlocking));
        if (!getenv("CUSTMA_HOST_NO_PREPARE"))
            CUSTMA_CUDA_CHECK(cudaStreamCreateWithPriority(&g_ctx.prep, cudaStreamNonBlocking, prio_low));
    }
    for (Slot &s : g_ctx.slot) {
        if (g_ctx.single) {
            s.stream = nullptr; s.d2h = nullptr;
            CUSTMA_CUDA_CHECK(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
            CUSTMA_CUDA_CHECK(cudaEventCreateWithFlags(&s.kernels_done, cudaEventDisableTiming));
            if (g_ctx.prep) {
                CUSTMA_CUDA_CHECK(cudaEventCreateWithFlags(&s.prep_done, cudaEventDisableTiming));
                CUSTMA_CUDA_CHECK(cudaEventCreateWithFlags(&s.fprep_done, cudaEventDisableTiming));
                s.ws_bwd_bytes = custma_backward_workspace_bytes(chunk, H, W, D, k, flags);
                CUSTMA_CUDA_CHECK(cudaMalloc(&s.ws_bwd, std::max<size_t>(s.ws_bwd_bytes, 256)));
            }
        } else {
            CUSTMA_CUDA_CHECK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
            CUSTMA_CUDA_CHECK(cudaStreamCreateWithFlags(&s.d2h, cudaStreamNonBlocking));
        }
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
    for (auto &row : g_ctx.done)
        for (cudaEvent_t &e : row) CUSTMA_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return CUSTMA_OK;
}

static bool ctx_matches(const HostCtx &c, int dev, int32_t chunk, int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags,
                        bool need_volume) {
    return c.live && c.device == dev && c.H == H && c.W == W && c.D == D && c.k == k && c.flags == flags &&
           c.chunk >= chunk && (c.with_volume || !need_volume);
}

// index of a context built for these arguments (reused, or built in the free / least recently used entry), < 0 on error
static int ensure_ctx(int32_t chunk, int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags, bool need_volume, int *index) {
    int dev = 0;
    CUSTMA_CUDA_CHECK(cudaGetDevice(&dev));
    int pick = -1;
    for (int i = 0; i < kContexts; ++i)
        if (ctx_matches(g_ctxs[i], dev, chunk, H, W, D, k, flags, need_volume)) pick = i;
    if (pick < 0) {
        // same device and shape but too small (chunk / volume): rebuild that one; else a free entry; else the oldest
        for (int i = 0; i < kContexts && pick < 0; ++i) {
            const HostCtx &c = g_ctxs[i];
            if (c.live && c.device == dev && c.H == H && c.W == W && c.D == D && c.k == k && c.flags == flags) pick = i;
        }
        for (int i = 0; i < kContexts && pick < 0; ++i)
            if (!g_ctxs[i].live) pick = i;
        if (pick < 0) {
            pick = 0;
            for (int i = 1; i < kContexts; ++i)
                if (g_use_stamp[i] < g_use_stamp[pick]) pick = i;
        }
        release_locked(g_ctxs[pick]);
        const int rc = build_ctx(g_ctxs[pick], chunk, H, W, D, k, flags, need_volume, dev);
        if (rc != CUSTMA_OK) {
            release_locked(g_ctxs[pick]);   // e.g. out of memory on the volume buffer: leave no half-built context behind
            cudaGetLastError();             // and no sticky allocation error for the caller's next CUDA call
            return rc;
        }
    }
    g_use_stamp[pick] = ++g_stamp;
    *index = pick;
    return CUSTMA_OK;
}

}  // namespace custma

using namespace custma;

extern "C" {

int custma_host_release(void) {
    std::lock_guard<std::mutex> lock(g_mutex);
    for (HostCtx &c : g_ctxs) release_locked(c);
    return CUSTMA_OK;
}

// how the host images of a submit are encoded: fp32 planes, or interleaved 8-bit images converted on the device
struct HostImages {
    const void *cam = nullptr, *proj = nullptr;
    bool u8 = false;
    int32_t cam_channels = 1, cam_channel = 0, proj_channels = 1, proj_channel = 0;
    float scale = 1.f;
};

static int submit(const HostImages &img, float *h_best, int32_t *h_index,
                  float *h_camera_grad, float *cost_volume_dev, const float *cost_volume_grad_dev, int32_t B,
                  int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags, bool streamed, uint64_t *ticket) {
    if (!img.cam || !img.proj || !h_best || !h_index)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "h_camera, h_projector, h_best and h_index must not be NULL");
    if ((cost_volume_grad_dev != nullptr) != (h_camera_grad != nullptr))
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "h_camera_grad is required iff cost_volume_grad_dev is given");
    if (B <= 0 || H <= 0 || W <= 0 || D < 0 || k < 1 || k > CUSTMA_MAX_KERNEL_SIZE)
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "bad shape B=%d H=%d W=%d D=%d k=%d", B, H, W, D, k);
    if (img.u8 && (img.cam_channels < 1 || img.proj_channels < 1 || img.cam_channel < 0 || img.cam_channel >= img.cam_channels ||
                   img.proj_channel < 0 || img.proj_channel >= img.proj_channels))
        return set_error(CUSTMA_ERR_INVALID_ARGUMENT, "bad channel selection: camera %d of %d, projector %d of %d", img.cam_channel,
                         img.cam_channels, img.proj_channel, img.proj_channels);
    std::lock_guard<std::mutex> lock(g_mutex);
    const int32_t chunk = pick_chunk(B, H, W, D, streamed);
    int ci = -1;
    int rc = ensure_ctx(chunk, H, W, D, k, flags, cost_volume_dev == nullptr, &ci);
    if (rc) return rc;
    HostCtx &g_ctx = g_ctxs[ci];
    const int32_t C = D > 0 ? D : W;
    const size_t pix = (size_t)H * W;
    for (int32_t b0 = 0; b0 < B; b0 += chunk, g_ctx.next_slot = (g_ctx.next_slot + 1) % kSlots) {
        Slot &s = g_ctx.slot[g_ctx.next_slot];
        const int32_t nb = std::min(chunk, B - b0);
        const size_t img_bytes = (size_t)nb * pix * sizeof(float);
        // streams of this chunk: copies in | kernels | copies out
        cudaStream_t sin = g_ctx.single ? g_ctx.h2d : s.stream, sk = g_ctx.single ? g_ctx.compute : s.stream,
                     sout = g_ctx.single ? g_ctx.d2h : s.d2h;
        if (g_ctx.single) {
            // the slot's image buffers are free once the kernels of its previous chunk are done
            CUSTMA_CUDA_CHECK(cudaStreamWaitEvent(sin, s.kernels_done, 0));
        } else {
            // the slot's result buffers are free once the device->host copies of its previous chunk are done
            CUSTMA_CUDA_CHECK(cudaStreamWaitEvent(sk, s.d2h_done, 0));
        }
        uint8_t *dc = nullptr, *dp = nullptr;
        if (!img.u8) {
            CUSTMA_CUDA_CHECK(cudaMemcpyAsync(s.cam, (const float *)img.cam + (size_t)b0 * pix, img_bytes, cudaMemcpyHostToDevice, sin));
            CUSTMA_CUDA_CHECK(cudaMemcpyAsync(s.proj, (const float *)img.proj + (size_t)b0 * pix, img_bytes, cudaMemcpyHostToDevice, sin));
        } else {
            // 8-bit images cross PCIe as they are (a quarter of the bytes per channel) and become fp32 planes on the device
            const size_t cb = (size_t)nb * pix * img.cam_channels, pb = (size_t)nb * pix * img.proj_channels;
            const size_t need = align256(cb) + align256(pb);
            if (s.u8_bytes < need) {
                CUSTMA_CUDA_CHECK(cudaStreamSynchronize(sin));
                CUSTMA_CUDA_CHECK(cudaStreamSynchronize(sk));
                cudaFree(s.u8);
                s.u8 = nullptr; s.u8_bytes = 0;
                CUSTMA_CUDA_CHECK(cudaMalloc(&s.u8, need));
                s.u8_bytes = need;
            }
            dc = s.u8; dp = s.u8 + align256(cb);
            CUSTMA_CUDA_CHECK(cudaMemcpyAsync(dc, (const uint8_t *)img.cam + (size_t)b0 * pix * img.cam_channels, cb, cudaMemcpyHostToDevice, sin));
            CUSTMA_CUDA_CHECK(cudaMemcpyAsync(dp, (const uint8_t *)img.proj + (size_t)b0 * pix * img.proj_channels, pb, cudaMemcpyHostToDevice, sin));
        }
        if (g_ctx.single) {
            CUSTMA_CUDA_CHECK(cudaEventRecord(s.h2d_done, sin));
            CUSTMA_CUDA_CHECK(cudaStreamWaitEvent(sk, s.h2d_done, 0));
            CUSTMA_CUDA_CHECK(cudaStreamWaitEvent(sk, s.d2h_done, 0));   // result buffers of the slot's previous chunk have left
        }
        if (img.u8) {
            if ((rc = custma_ingest_u8(dc, s.cam, nb, H, W, img.cam_channels, img.cam_channel, img.scale, sk))) return rc;
            if ((rc = custma_ingest_u8(dp, s.proj, nb, H, W, img.proj_channels, img.proj_channel, img.scale, sk))) return rc;
        }
        // The preparations of the forward and of the backward need the images only: they run on the lowest-priority
        // stream as soon as the chunk's images are on the device - beside the previous chunk's kernels, whose grids
        // leave their last waves mostly idle - each into its own workspace.
        const bool prepare = g_ctx.single && g_ctx.prep && !img.u8;
        if (prepare) {
            CUSTMA_CUDA_CHECK(cudaStreamWaitEvent(g_ctx.prep, s.h2d_done, 0));
            CUSTMA_CUDA_CHECK(cudaStreamWaitEvent(g_ctx.prep, s.kernels_done, 0));   // the workspaces' previous user
            if ((rc = custma_forward_prepare(s.cam, s.proj, nb, H, W, D, k, flags, s.ws, s.ws_bytes, g_ctx.prep))) return rc;
            CUSTMA_CUDA_CHECK(cudaEventRecord(s.fprep_done, g_ctx.prep));
            CUSTMA_CUDA_CHECK(cudaStreamWaitEvent(sk, s.fprep_done, 0));
            if (cost_volume_grad_dev) {
                if ((rc = custma_backward_prepare(s.cam, s.proj, nb, H, W, D, k, flags, s.ws_bwd, s.ws_bwd_bytes, g_ctx.prep))) return rc;
                CUSTMA_CUDA_CHECK(cudaEventRecord(s.prep_done, g_ctx.prep));
            }
        }
        float *vol = cost_volume_dev ? cost_volume_dev + (size_t)b0 * pix * C : s.vol;
        rc = custma_forward(s.cam, s.proj, vol, s.best, s.index, nb, H, W, D, k, flags | (prepare ? CUSTMA_FLAG_PREPARED : 0u),
                            s.ws, s.ws_bytes, sk);
        if (rc) return rc;
        // results leave on a copy stream, so the backward kernels do not queue behind the copies
        CUSTMA_CUDA_CHECK(cudaEventRecord(s.fwd_done, sk));
        CUSTMA_CUDA_CHECK(cudaStreamWaitEvent(sout, s.fwd_done, 0));
        CUSTMA_CUDA_CHECK(cudaMemcpyAsync(h_best + (size_t)b0 * pix, s.best, img_bytes, cudaMemcpyDeviceToHost, sout));
        CUSTMA_CUDA_CHECK(cudaMemcpyAsync(h_index + (size_t)b0 * pix, s.index, (size_t)nb * pix * sizeof(int32_t),
                                          cudaMemcpyDeviceToHost, sout));
        if (cost_volume_grad_dev) {
            if (prepare) {
                CUSTMA_CUDA_CHECK(cudaStreamWaitEvent(sk, s.prep_done, 0));
                rc = custma_backward(cost_volume_grad_dev + (size_t)b0 * pix * C, s.cam, s.proj, s.grad, nb, H, W, D, k,
                                     flags | CUSTMA_FLAG_PREPARED, s.ws_bwd, s.ws_bwd_bytes, sk);
            } else {
                rc = custma_backward(cost_volume_grad_dev + (size_t)b0 * pix * C, s.cam, s.proj, s.grad, nb, H, W, D, k,
                                     flags, s.ws, s.ws_bytes, sk);
            }
            if (rc) return rc;
            CUSTMA_CUDA_CHECK(cudaEventRecord(s.bwd_done, sk));
            CUSTMA_CUDA_CHECK(cudaStreamWaitEvent(sout, s.bwd_done, 0));
            CUSTMA_CUDA_CHECK(cudaMemcpyAsync(h_camera_grad + (size_t)b0 * pix, s.grad, img_bytes, cudaMemcpyDeviceToHost, sout));
        }
        if (g_ctx.single) CUSTMA_CUDA_CHECK(cudaEventRecord(s.kernels_done, sk));
        CUSTMA_CUDA_CHECK(cudaEventRecord(s.d2h_done, sout));
    }
    const uint64_t t = g_ctx.next_ticket++;
    for (int i = 0; i < kSlots; ++i)
        CUSTMA_CUDA_CHECK(cudaEventRecord(g_ctx.done[t % kTickets][i], g_ctx.single ? g_ctx.d2h : g_ctx.slot[i].d2h));
    if (ticket) *ticket = ((uint64_t)(ci + 1) << 56) | t;
    return CUSTMA_OK;
}

static HostImages fp32_images(const float *cam, const float *proj) {
    HostImages img;
    img.cam = cam; img.proj = proj;
    return img;
}

int custma_host_submit_u8(const uint8_t *h_camera_u8, int32_t camera_channels, int32_t camera_channel,
                          const uint8_t *h_projector_u8, int32_t projector_channels, int32_t projector_channel, float scale,
                          float *h_best, int32_t *h_index, float *h_camera_grad, float *cost_volume_dev,
                          const float *cost_volume_grad_dev, int32_t B, int32_t H, int32_t W, int32_t D, int32_t k,
                          uint32_t flags, uint64_t *ticket) {
    HostImages img;
    img.cam = h_camera_u8; img.proj = h_projector_u8; img.u8 = true;
    img.cam_channels = camera_channels; img.cam_channel = camera_channel;
    img.proj_channels = projector_channels; img.proj_channel = projector_channel;
    img.scale = scale;
    return submit(img, h_best, h_index, h_camera_grad, cost_volume_dev, cost_volume_grad_dev, B, H, W, D, k, flags, true, ticket);
}

int custma_host_submit(const float *h_camera, const float *h_projector, float *h_best, int32_t *h_index,
                       float *h_camera_grad, float *cost_volume_dev, const float *cost_volume_grad_dev, int32_t B,
                       int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags, uint64_t *ticket) {
    return submit(fp32_images(h_camera, h_projector), h_best, h_index, h_camera_grad, cost_volume_dev, cost_volume_grad_dev,
                  B, H, W, D, k, flags, true, ticket);
}

int custma_host_wait(uint64_t ticket) {
    std::lock_guard<std::mutex> lock(g_mutex);
    const int ci = (int)(ticket >> 56) - 1;
    const uint64_t seq = ticket & (((uint64_t)1 << 56) - 1);
    if (ticket != 0 && ci >= 0 && ci < kContexts) {
        HostCtx &c = g_ctxs[ci];
        if (c.live && seq != 0 && seq < c.next_ticket && seq + kTickets > c.next_ticket) {
            for (int i = 0; i < kSlots; ++i) CUSTMA_CUDA_CHECK(cudaEventSynchronize(c.done[seq % kTickets][i]));
            return CUSTMA_OK;
        }
    }
    // 0, an unknown ticket, or one whose events have been reused (later work completes later): everything submitted so far
    for (HostCtx &c : g_ctxs) {
        if (!c.live) continue;
        for (Slot &s : c.slot) {
            if (s.stream) CUSTMA_CUDA_CHECK(cudaStreamSynchronize(s.stream));
            if (s.d2h) CUSTMA_CUDA_CHECK(cudaStreamSynchronize(s.d2h));
        }
        for (cudaStream_t st : {c.h2d, c.prep, c.compute, c.d2h})
            if (st) CUSTMA_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    return CUSTMA_OK;
}

int custma_host_step(const float *h_camera, const float *h_projector, float *h_best, int32_t *h_index,
                     float *h_camera_grad, float *cost_volume_dev, const float *cost_volume_grad_dev, int32_t B,
                     int32_t H, int32_t W, int32_t D, int32_t k, uint32_t flags) {
    uint64_t t = 0;
    const int rc = submit(fp32_images(h_camera, h_projector), h_best, h_index, h_camera_grad, cost_volume_dev,
                          cost_volume_grad_dev, B, H, W, D, k, flags, false, &t);
    return rc ? rc : custma_host_wait(0);
}

}  // extern "C"
