import os
import sys
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests skip (not fail) on a machine without a CUDA device, so a plain `pytest` run is green there."""
    try:
        import vkhrt_b200
        n_dev = vkhrt_b200.device_count()
    except Exception:
        n_dev = 0
    if n_dev > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (vkhrt_b200 has no CPU path)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle if they are missing (the GPU box gets them prebuilt)."""
    from oracle import oracle as O
    O.build()
    lib = os.path.join(ROOT, "vkhrt_b200", "_lib", "libvkhrt_b200.so")
    exe = os.path.join(ROOT, "vkhrt_b200", "_lib", "vkhrt_headless")
    if not os.path.exists(lib) or not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "vkhrt_b200", "csrc")])
    yield


def has_gpu():
    import vkhrt_b200 as V
    return V.device_count() > 0


@pytest.fixture(scope="session")
def V():
    import vkhrt_b200
    return vkhrt_b200


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    return oracle


def default_camera(V, width, height):
    return V.camera_matrices(aspect=float(np.float32(width) / np.float32(height)))


def psnr(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    mse = np.mean((a - b) ** 2)
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)
