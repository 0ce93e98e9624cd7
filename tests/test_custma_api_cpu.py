"""Host-side behaviour of the drop-in `custma` package that needs no GPU: exports, Timer, input checks."""
import re
import time

import pytest
import torch

import custma
from custma.utils import Timer, TimerError


def test_exports_match_reference():
    # custma/__init__.py:2-6 of the reference
    assert custma.__version__ == "0.0.1"
    for name in ("stereo_matching", "Timer", "version", "src", "stereo_matching_wrapper", "utils"):
        assert name in custma.__all__
    assert callable(custma.stereo_matching)
    assert callable(custma.src.stereo_matching_forward) and callable(custma.src.stereo_matching_backward)


def test_cpu_tensors_raise_like_check_input():
    cam = torch.rand(8, 8)
    with pytest.raises(RuntimeError, match="camera must be a CUDA tensor"):
        custma.stereo_matching(cam, cam, 4, 5)
    with pytest.raises(RuntimeError, match="camera must be a CUDA tensor"):
        custma.src.stereo_matching_forward(cam, cam, 4, 5)
    with pytest.raises(RuntimeError, match="cost_volume_grad must be a CUDA tensor"):
        custma.src.stereo_matching_backward(torch.rand(8, 8, 8), cam, cam, 5)


def test_timer_template_rules(capsys):
    with Timer("it takes {:.1f} seconds"):
        time.sleep(0.01)
    out = capsys.readouterr().out
    assert re.fullmatch(r"it takes \d\.\d seconds\n", out)
    with Timer("cuda forward time"):       # no float field -> " {:.3f}" appended (custma/utils.py:38-42)
        pass
    assert re.fullmatch(r"cuda forward time \d+\.\d{3}\n", capsys.readouterr().out)
    with Timer():
        pass
    assert re.fullmatch(r"\d+\.\d{3}\n", capsys.readouterr().out)


def test_timer_checks():
    t = Timer()
    assert t.is_running
    time.sleep(0.02)
    a = t.since_start()
    time.sleep(0.02)
    b = t.since_last_check()
    c = t.since_start()
    assert a >= 0.02 and b >= 0.02 and c >= a + b - 1e-3
    idle = Timer(start=False)
    assert not idle.is_running
    with pytest.raises(TimerError) as e:
        idle.since_start()
    assert e.value.message == "timer is not running"
    with pytest.raises(TimerError):
        idle.since_last_check()
    idle.start()
    assert idle.since_last_check() >= 0
