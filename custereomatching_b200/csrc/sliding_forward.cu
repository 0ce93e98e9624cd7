// placeholder until the sliding-window kernels land
#include "common.cuh"
namespace custma {
bool sliding_supported(const Problem &) { return false; }
size_t sliding_forward_workspace_bytes(const Problem &) { return 0; }
size_t sliding_backward_workspace_bytes(const Problem &) { return 0; }
int launch_sliding_forward(const Problem &, const float *, const float *, float *, float *, int32_t *, void *, size_t, cudaStream_t) {
    return set_error(CUSTMA_ERR_UNSUPPORTED, "sliding forward not built");
}
int launch_sliding_backward(const Problem &, const float *, const float *, const float *, float *, void *, size_t, cudaStream_t) {
    return set_error(CUSTMA_ERR_UNSUPPORTED, "sliding backward not built");
}
}
