"""Host-side sweep of the fast kernels' tiling (custma_debug_validate_layout): for many shapes, every row segment the
kernels copy out of the workspace lies inside its row, is 16-byte aligned, and the disparity chunks cover what every
column needs.  Runs without a GPU (compute-sanitizer is not available on the test pool)."""
import itertools

import pytest

from custereomatching_b200 import binding


def test_baseline_configs():
    for B, H, W, D, k in [(1, 240, 320, 64, 5), (1, 375, 1242, 192, 5), (1, 1988, 2880, 256, 5), (64, 375, 1242, 192, 5),
                          (1, 540, 7680, 512, 5), (1, 4320, 7680, 512, 5), (1, 375, 1242, 0, 5), (1, 330, 422, 0, 5)]:
        binding.validate_layout(B, H, W, D, k)
        assert binding.forward_workspace_bytes(B, H, W, D, k) > 0
        assert binding.backward_workspace_bytes(B, H, W, D, k) > 0


@pytest.mark.parametrize("k", [3, 5, 7])
def test_shape_sweep(k):
    Hs = [1, 2, 5, 16, 17, 63, 64, 65, 129, 375, 1000]
    Ws = [1, 3, 4, 15, 16, 17, 47, 48, 49, 64, 65, 100, 191, 192, 193, 255, 256, 257, 422, 1242, 2049]
    Ds = [0, 1, 2, 3, 4, 7, 63, 64, 65, 100, 127, 128, 129, 191, 192, 193, 255, 256, 257, 300, 512, 700]
    for H, W, D in itertools.product(Hs, Ws, Ds):
        binding.validate_layout(1, H, W, D, k)


def test_invalid_arguments_are_reported():
    with pytest.raises(RuntimeError, match="kernel_size"):
        binding.validate_layout(1, 8, 8, 4, 0)
    with pytest.raises(RuntimeError, match="positive"):
        binding.validate_layout(1, 0, 8, 4, 5)
