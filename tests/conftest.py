"""Shared test plumbing.

* ``-m "not gpu"``: oracle vs golden vectors, host logic, C-ABI symbol checks (no compute calls).
* ``-m gpu``: parity tests proper — the CUDA path (through the C ABI) vs ``oracle/`` and vs the
  unmodified reference extensions in ``oracle/_ref`` on a B200.
Nothing here reads /root/reference at run time (it does not exist on the GPU box).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "segment-anything-nerf_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _load_npz(name):
    path = os.path.join(GOLDEN_DIR, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated yet (oracle/make_golden.py)")
    return np.load(path)


@pytest.fixture(scope="session")
def ref_cpu():
    """Vectors recorded from the reference's pure-torch modules (oracle/make_golden.py --cpu)."""
    return _load_npz("ref_cpu.npz")


@pytest.fixture(scope="session")
def ref_gpu():
    """Vectors recorded from the reference CUDA extensions on a B200 (oracle/make_golden.py --gpu)."""
    return _load_npz("ref_gpu.npz")


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def ref_ext():
    """The unmodified reference extensions rebuilt for sm_100 (oracle/_ref); loader only."""
    from oracle import build_ref

    def get(name):
        try:
            return build_ref.load(name)
        except FileNotFoundError as e:
            pytest.skip(str(e))

    return get
