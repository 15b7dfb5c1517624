import os, sys, ctypes, subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    """The C restatements (oracle/_build/liboracle.so); built on demand with gcc."""
    import oracle_py
    return oracle_py.load_oracle()


@pytest.fixture(scope="session")
def ref_dir():
    """oracle/_ref (the reference's own code compiled here) or None where it was never built."""
    d = os.path.join(ROOT, "oracle", "_ref")
    return d if os.path.isdir(d) else None


@pytest.fixture(scope="session")
def ctx():
    import ucoslam_b200
    c = ucoslam_b200.Context(0)
    yield c
    c.close()
