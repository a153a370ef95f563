import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_checkpoint(name="checkpoint_3class.npz"):
    ck = np.load(GOLDEN / name)
    return {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck["__order__"]}


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def checkpoint():
    return load_checkpoint()


@pytest.fixture(scope="session")
def windows():
    w = np.load(GOLDEN / "eeg_windows.npz")
    return {"X": w["X"], "names": list(w["names"]), "prefix": list(w["prefix"])}


@pytest.fixture(scope="session")
def built_lib():
    """Compile (or reuse) the in-tree shared object; needs nvcc but no GPU."""
    from neural_speech_decoding_b200.build import build_library
    return build_library()


_CPU_KERNELS_REGISTERED = False


@pytest.fixture
def cpu_backend(monkeypatch):
    """HOST-LOGIC tests only: route the C ABI to tests/fake_lib.FakeLib (numpy on host pointers) and
    let the neuroalpha::* ops accept CPU tensors.  Never used by -m gpu tests."""
    global _CPU_KERNELS_REGISTERED
    from tests.fake_lib import FakeLib
    from neural_speech_decoding_b200 import _lib, ops
    fake = FakeLib()
    monkeypatch.setattr(_lib, "load", lambda path=None: fake)
    monkeypatch.setattr(ops, "_require_cuda", lambda *t: None)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(ops, "compute_device", lambda d: torch.device("cpu"))
    monkeypatch.setattr(ops, "EXACT_TC", False)      # the fake implements the granular fp32 entry points (host-logic path)
    monkeypatch.setattr(ops, "EXACT_TC_TRAIN", False)
    if not _CPU_KERNELS_REGISTERED:
        for op in ops.all_custom_ops():
            op.register_kernel("cpu")(op._init_fn)
        _CPU_KERNELS_REGISTERED = True
    return fake
