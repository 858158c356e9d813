import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        from asr_rescoring_b200 import _lib
        return _lib.load().pllb_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gold_dir():
    return GOLD


@pytest.fixture(scope="session")
def require_gpu():
    # -m gpu tests must FAIL (not skip) without a device: the product has no CPU fallback
    assert _has_gpu(), "no B200 visible: -m gpu tests need the GPU box"
