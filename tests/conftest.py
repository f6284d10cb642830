import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: full-size CPU oracle runs (opt in with KW_SLOW=1)")


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "slow" in item.keywords and os.environ.get("KW_SLOW") != "1":
            item.add_marker(pytest.mark.skip(reason="set KW_SLOW=1 to run full-size CPU oracle checks"))


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return {n: np.load(os.path.join(d, f"{n}.npz")) for n in ("logmel", "tiny", "kotoba", "teacher")}


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build libkwb200.so if it is missing (nvcc cross-compiles without a GPU)."""
    from kotoba_whisper_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from kotoba_whisper_b200.build import build
        build()
