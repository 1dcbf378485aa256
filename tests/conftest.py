import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 GPUs")
    config.addinivalue_line("markers", "extended: longer stress / redundancy tests, run with KUNGFU_TEST_EXTENDED=1")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        have = torch.cuda.is_available()
        ngpu = torch.cuda.device_count() if have else 0
    except Exception:
        have, ngpu = False, 0
    extended = os.environ.get("KUNGFU_TEST_EXTENDED") == "1"
    for item in items:
        if "extended" in item.keywords and not extended:
            item.add_marker(pytest.mark.skip(reason="extended test: set KUNGFU_TEST_EXTENDED=1"))
        if "gpu" in item.keywords and not have:
            item.add_marker(pytest.mark.skip(reason="no GPU in this container"))
        if "multigpu" in item.keywords and ngpu < 2:
            item.add_marker(pytest.mark.skip(reason="needs >= 2 GPUs"))


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Native pieces are built in-tree once per session."""
    from kungfu_b200 import _build

    _build.build_all()
