import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "needs_grad: test runs with autograd enabled (training path, row f2)")


@pytest.fixture(autouse=True)
def _no_grad_by_default(request):
    """The evaluation path is what these tests are about: utils/eval.py and the notebooks call the model under
    torch.no_grad(), and that is where `Aline.forward` runs the CUDA kernels through the C ABI.  With autograd enabled
    and trainable parameters `Aline.forward` switches to the differentiable torch-op composition (row f2) -- a test that
    forgot no_grad() would silently check that instead of the kernels.  Training tests opt in with `needs_grad`."""
    import torch
    if request.node.get_closest_marker("needs_grad"):
        yield
        return
    with torch.no_grad():
        yield


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
