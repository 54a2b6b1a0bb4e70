"""pytest configuration: `gpu` marker (tests that need a B200) and shared fixtures."""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def lib_path():
    """Path of the in-tree shared library, building it first if it is missing (nvcc cross-compiles on CPU)."""
    import reachability_jl_b200 as rb
    from reachability_jl_b200 import _build
    if not rb.LIB_PATH.exists():
        _build.build_library()
    return rb.LIB_PATH


@pytest.fixture(scope="session")
def ctx(lib_path):
    """One `reach_ctx` on cuda:0 for the whole GPU session.  Fails loudly without a GPU (no CPU fallback)."""
    import reachability_jl_b200 as rb
    c = rb.Context(int(os.environ.get("REACH_TEST_DEVICE", "0")))
    yield c
    c.close()


@pytest.fixture()
def ctx_on_torch_stream(lib_path):
    """A `reach_ctx` that borrows a torch CUDA stream (reach_ctx_create_on_stream): what the stream-ordered multi-GPU driver
    uses, so that torch / NCCL work and the library's kernels are ordered on the device."""
    import torch
    import reachability_jl_b200 as rb
    dev = int(os.environ.get("REACH_TEST_DEVICE", "0"))
    stream = torch.cuda.Stream(device=dev)
    c = rb.Context(dev, stream=stream.cuda_stream)
    yield c, stream
    c.close()
