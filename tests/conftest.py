import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)


def float64_equal(a, b):
    """dsputils/compare.go:94-96 Float64Equal: |a-b| <= 1e-8 or |1-a/b| <= 1e-8."""
    a, b = float(a), float(b)
    if abs(a - b) <= 1e-8:
        return True
    return b != 0 and abs(1 - a / b) <= 1e-8


def pretty_close(a, b):
    """dsputils/compare.go:27-38 PrettyClose / :41-52 PrettyCloseC (per real/imag part)."""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    a, b = a.astype(np.complex128).ravel(), b.astype(np.complex128).ravel()
    return all(float64_equal(x.real, y.real) and float64_equal(x.imag, y.imag) for x, y in zip(a, b))


def cplx(pairs):
    a = np.asarray(pairs, dtype=np.float64)
    return a[..., 0] + 1j * a[..., 1]


def rel_l2(a, b):
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d else float(np.linalg.norm(a - b))
