"""The C-ABI library loads on a CPU-only box and exports every symbol include/custma_b200.h declares; argument
validation works without a GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from custereomatching_b200 import binding, build

HEADER = os.path.join(ROOT, "include", "custma_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(custma_[a-z_0-9]+)\s*\(", text)))


def test_library_is_built_and_loads():
    build.build()
    lib = binding.load()
    assert lib.custma_abi_version() == binding.ABI_VERSION


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert names == sorted(binding.SYMBOLS)
    raw = ctypes.CDLL(binding.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/custma_b200.h but not exported"


def test_header_cites_reference():
    text = open(HEADER).read()
    for cite in ("custma/src/stereo_matching.cpp:16-42", "custma/src/stereo_matching.cpp:45-73",
                 "custma/src/stereo_matching_kernel.cu:17-72", "custma/src/stereo_matching_kernel.cu:75-152"):
        assert cite in text


def test_header_constants_match_binding():
    text = open(HEADER).read()
    assert f"#define CUSTMA_ABI_VERSION {binding.ABI_VERSION}" in text
    assert "#define CUSTMA_INVALID_COST (-2.0f)" in text and binding.INVALID_COST == -2.0
    assert "#define CUSTMA_FLAG_DIRECT 1u" in text and binding.FLAG_DIRECT == 1
    assert "#define CUSTMA_FLAG_TENSOR 2u" in text and binding.FLAG_TENSOR == 2


@pytest.mark.parametrize("args", [(0, 4, 4, 0, 5), (1, 0, 4, 0, 5), (1, 4, 4, -1, 5), (1, 4, 4, 0, 0), (1, 4, 4, 0, 32)])
def test_workspace_query_rejects_bad_arguments(args):
    assert binding.forward_workspace_bytes(*args) == 0
    assert binding.last_error() != ""
    assert binding.backward_workspace_bytes(*args) == 0


def test_workspace_query_accepts_good_arguments():
    assert binding.forward_workspace_bytes(1, 240, 320, 64, 5) > 0
    assert binding.backward_workspace_bytes(2, 375, 1242, 192, 5) > 0
    assert binding.forward_workspace_bytes(1, 240, 320, 0, 15, binding.FLAG_DIRECT) > 0
    # the backward workspace holds the tensor-core kernel's halo tiles wherever that kernel can take the call
    assert binding.backward_workspace_bytes(1, 64, 256, 64, 5) > binding.backward_workspace_bytes(1, 64, 256, 62, 5) - 64 * 256 * 8
    assert binding.forward_workspace_bytes(1, 240, 320, 64, 5, binding.FLAG_TENSOR) > 0


def test_null_pointers_are_rejected_before_any_cuda_call():
    lib = binding.load()
    rc = lib.custma_forward(None, None, None, None, None, 1, 4, 4, 0, 5, 0, None, 0, None)
    assert rc == binding.ERR_INVALID_ARGUMENT
    rc = lib.custma_backward(None, None, None, None, 1, 4, 4, 0, 5, 0, None, 0, None)
    assert rc == binding.ERR_INVALID_ARGUMENT
    rc = lib.custma_host_step(None, None, None, None, None, None, None, 1, 4, 4, 0, 5, 0)
    assert rc == binding.ERR_INVALID_ARGUMENT
    with pytest.raises(RuntimeError, match="custma_forward failed"):
        binding.forward(0, 0, 0, 0, 0, 1, 4, 4, 0, 5, 0, 0, 0, 0)
