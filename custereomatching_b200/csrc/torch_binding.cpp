// custma.src - the native torch module of the drop-in package: the reference's two host entry points
//   stereo_matching_forward(camera, projector, D, kernel_size) -> Tensor[H, W, W]
//   stereo_matching_backward(cost_volume_grad, camera, projector, kernel_size) -> Tensor[H, W]
// (reference: custma/src/bindings.cpp:4-7, custma/src/stereo_matching.cpp:16-73, custma/include/stereo_matching.hpp)
// implemented as thin calls into the torch-free C ABI of libcustma_b200.so (include/custma_b200.h).  This is the only
// translation unit that sees torch headers; it owns allocation, the input checks and the stream / device plumbing.
//
// Behaviour kept from the reference: positional signatures; D accepted and ignored (the volume is [H, W, W]); the
// CHECK_INPUT messages ("camera must be a CUDA tensor", "camera must be contiguous"); callee allocates.
// Deliberate differences: kernels run on torch's current stream of the inputs' device instead of
// the legacy default stream; shape / dtype mismatches raise instead of being undefined behaviour; outputs are not
// zero-filled first; every C-ABI failure becomes a c10::Error carrying custma_last_error().
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "custma_b200.h"

namespace {

#define CUSTMA_CHECK_CUDA(x) TORCH_CHECK((x).is_cuda(), #x " must be a CUDA tensor")
#define CUSTMA_CHECK_CONTIGUOUS(x) TORCH_CHECK((x).is_contiguous(), #x " must be contiguous")
#define CUSTMA_CHECK_INPUT(x) \
    CUSTMA_CHECK_CUDA(x);     \
    CUSTMA_CHECK_CONTIGUOUS(x)
#define CUSTMA_CHECK_FLOAT(x) \
    TORCH_CHECK((x).scalar_type() == at::kFloat, #x " must be a float32 tensor, got ", (x).scalar_type())

void check_pair(const at::Tensor &camera, const at::Tensor &projector) {
    TORCH_CHECK(camera.dim() == 2, "camera must be a 2-D [H, W] tensor, got ", camera.sizes());
    TORCH_CHECK(camera.sizes() == projector.sizes(), "camera ", camera.sizes(), " and projector ", projector.sizes(),
                " must have the same shape");
    TORCH_CHECK(camera.device() == projector.device(), "camera and projector must be on the same device");
    TORCH_CHECK(camera.numel() > 0, "empty input ", camera.sizes());
}

at::Tensor workspace(size_t bytes, const at::Tensor &like) {
    TORCH_CHECK(bytes > 0, "custma workspace query failed: ", custma_last_error());
    return at::empty({(int64_t)bytes}, like.options().dtype(at::kByte));
}

at::Tensor stereo_matching_forward(const at::Tensor &camera, const at::Tensor &projector, const int32_t D,
                                   const int32_t kernel_size) {
    (void)D;  // accepted and ignored, as in the reference (stereo_matching_kernel.cu:14)
    CUSTMA_CHECK_INPUT(camera);
    CUSTMA_CHECK_INPUT(projector);
    CUSTMA_CHECK_FLOAT(camera);
    CUSTMA_CHECK_FLOAT(projector);
    check_pair(camera, projector);
    const c10::cuda::CUDAGuard guard(camera.device());
    const int32_t H = (int32_t)camera.size(0), W = (int32_t)camera.size(1);
    at::Tensor cost_volume = at::empty({H, W, W}, camera.options());
    const size_t ws_bytes = custma_forward_workspace_bytes(1, H, W, 0, kernel_size, 0);
    at::Tensor ws = workspace(ws_bytes, camera);
    const int rc = custma_forward(camera.data_ptr<float>(), projector.data_ptr<float>(), cost_volume.data_ptr<float>(),
                                  nullptr, nullptr, 1, H, W, 0, kernel_size, 0, ws.data_ptr(), ws_bytes,
                                  at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == CUSTMA_OK, "custma_forward failed (code ", rc, "): ", custma_last_error());
    return cost_volume;
}

at::Tensor stereo_matching_backward(const at::Tensor &cost_volume_grad, const at::Tensor &camera,
                                    const at::Tensor &projector, const int32_t kernel_size) {
    CUSTMA_CHECK_INPUT(cost_volume_grad);  // the reference checks only the gradient (stereo_matching.cpp:52)
    CUSTMA_CHECK_INPUT(camera);
    CUSTMA_CHECK_INPUT(projector);
    CUSTMA_CHECK_FLOAT(cost_volume_grad);
    CUSTMA_CHECK_FLOAT(camera);
    CUSTMA_CHECK_FLOAT(projector);
    check_pair(camera, projector);
    const int32_t H = (int32_t)camera.size(0), W = (int32_t)camera.size(1);
    TORCH_CHECK(cost_volume_grad.dim() == 3, "cost_volume_grad must be a 3-D [H, W, W] tensor, got ", cost_volume_grad.sizes());
    TORCH_CHECK(cost_volume_grad.size(0) == H && cost_volume_grad.size(1) == W && cost_volume_grad.size(2) == W,
                "cost_volume_grad must have shape [", H, ", ", W, ", ", W, "], got ", cost_volume_grad.sizes());
    TORCH_CHECK(cost_volume_grad.device() == camera.device(), "cost_volume_grad must be on the images' device");
    const c10::cuda::CUDAGuard guard(camera.device());
    at::Tensor camera_grad = at::empty({H, W}, camera.options());
    const size_t ws_bytes = custma_backward_workspace_bytes(1, H, W, 0, kernel_size, 0);
    at::Tensor ws = workspace(ws_bytes, camera);
    const int rc = custma_backward(cost_volume_grad.data_ptr<float>(), camera.data_ptr<float>(), projector.data_ptr<float>(),
                                   camera_grad.data_ptr<float>(), 1, H, W, 0, kernel_size, 0, ws.data_ptr(), ws_bytes,
                                   at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == CUSTMA_OK, "custma_backward failed (code ", rc, "): ", custma_last_error());
    return camera_grad;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "custma.src: ZNCC cost volume forward / backward on libcustma_b200.so (B200, sm_100a)";
    m.def("stereo_matching_forward", &stereo_matching_forward);
    m.def("stereo_matching_backward", &stereo_matching_backward);
}
