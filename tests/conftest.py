import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """Make sure libtt_b200.so matches the sources (incremental: a no-op when it is up to date).  The product never
    builds itself at import time — a missing library is a loud error there — but a test run should not depend on
    somebody having called __graft_entry__.build() first."""
    import importlib.util
    import shutil

    path = os.path.join(ROOT, "two-towers-overlords_b200", "csrc", "build.py")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not (os.path.exists(nvcc) or shutil.which("nvcc")):
        return
    spec = importlib.util.spec_from_file_location("tt_b200_build", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        mod.build()
    except Exception as e:  # noqa: BLE001
        # a library that is missing or OLDER than any source must never be tested silently
        srcs = [os.path.join(mod.HERE, f) for f in os.listdir(mod.HERE) if f.endswith((".cu", ".cuh"))]
        srcs.append(os.path.join(ROOT, "include", "tt_b200.h"))
        if not os.path.exists(mod.LIB) or any(os.path.getmtime(f) > os.path.getmtime(mod.LIB) for f in srcs):
            raise RuntimeError(f"{mod.LIB} is missing or older than its sources and could not be rebuilt: {e}") from e
        print(f"warning: could not re-run the build, but {mod.LIB} is newer than every source: {e}")


def pytest_collection_modifyitems(config, items):
    """GPU-marked tests are skipped automatically where no CUDA device exists."""
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
