import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    """True on a GPU box.  A fresh box occasionally fails the very first CUDA initialisation of a process (seen once:
    "CUDA driver initialization failed" in pytest, while the next process on the same box ran fine), and a failed
    initialisation is cached by torch for the life of the process -- so when nvidia-smi is present, probe in short-lived
    subprocesses first and only then initialise CUDA here."""
    import shutil
    import subprocess
    import time

    if shutil.which("nvidia-smi") is not None:
        for _ in range(6):
            try:
                r = subprocess.run([sys.executable, "-c", "import torch,sys; sys.exit(0 if torch.cuda.is_available() and torch.zeros(1, device='cuda').item() == 0 else 1)"],
                                   capture_output=True, timeout=180)
                if r.returncode == 0:
                    break
            except Exception:
                pass
            time.sleep(5)
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g

    return g.load_package()


@pytest.fixture(scope="session")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def dev():
    import torch

    return torch.device("cuda:0")
